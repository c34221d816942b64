"""Dev tool: top stall-sampled SASS instructions of an ncu report (source page), with their dominant stall reason."""
import csv
import subprocess
import sys


def main(rep, top=40):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ci = hdr.index("Source")
    cs = hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for k, r in enumerate(rows[2:]):
        try:
            n = int(r[cs])
        except (ValueError, IndexError):
            continue
        st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
        data.append((n, k, r[ci].strip(), st))
    tot = sum(d[0] for d in data)
    print(f"total samples {tot}")
    for n, k, src, st in sorted(data, reverse=True)[:top]:
        print(f"{n:7d} {100 * n / tot:5.1f}%  line {k:5d}  {src[:70]:70s} {st}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)


def totals(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = {hdr[i]: 0 for i in stall_cols}
    for r in rows[2:]:
        for i in stall_cols:
            try:
                tot[hdr[i]] += int(r[i] or 0)
            except (ValueError, IndexError):
                pass
    s = sum(tot.values())
    print("stall totals:", ", ".join(f"{k[6:]} {100 * v / s:.1f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v))
