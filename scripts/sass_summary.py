"""Dev tool (build container, no GPU needed): per-kernel counts of the Blackwell-native SASS mnemonics in
edtr_b200/libedtr_b200.so (cuobjdump -sass) -> profiles/sass_summary.txt.  UTC*MMA = tcgen05.mma, LDTM / STTM =
tcgen05.ld / st, UTMALDG / UTMASTG = TMA loads / stores (B200_PROFILING.md)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "edtr_b200", "libedtr_b200.so")
PAT = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UBLKCP", "HMMA", "SYNCS", "UCGABAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        kernels[cur]["_total"] += 1
        for p in PAT:
            if p == "UTCHMMA.2CTA":
                kernels[cur][p] += int(op.startswith("UTCHMMA") and ".2CTA" in op)
            elif op == p or op.startswith(p + "."):
                kernels[cur][p] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    lines = ["# cuobjdump -sass edtr_b200/libedtr_b200.so (sm_100a): instruction counts per kernel",
             f"# {'kernel':58s} {'instrs':>7s} " + " ".join(f"{p:>12s}" for p in PAT)]
    for (k, c), name in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", name)[:58]
        lines.append(f"  {short:58s} {c['_total']:7d} " + " ".join(f"{c[p]:12d}" for p in PAT))
    path = os.path.join(ROOT, "profiles", "sass_summary.txt")
    open(path, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    sys.exit(main())
