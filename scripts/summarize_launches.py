"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by (kernel, grid): share of the step per
kernel family.  Usage: python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.txt"""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    fam = collections.defaultdict(lambda: [0, 0.0])
    dram = collections.defaultdict(float)     # kernel family -> DRAM bytes (read + write), when the list has them
    for row in csv.DictReader(lines):
        if row.get("Metric Name") in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v = float(row["Metric Value"].replace(",", ""))
            unit = row["Metric Unit"].lower()
            mult = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
            dram[row["Kernel Name"].split("(")[0].replace("void ", "")[-60:]] += v * mult
            continue
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"]
        short = name.split("(")[0].replace("void ", "")[-60:]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        us = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        agg[(short, row["Grid Size"], row["Block Size"])][0] += 1
        agg[(short, row["Grid Size"], row["Block Size"])][1] += us
        fam[short][0] += 1
        fam[short][1] += us
    tot = sum(v[1] for v in fam.values())
    n = sum(v[0] for v in fam.values())
    print(f"# {path}: {n} launches, {tot / 1e3:.3f} ms summed kernel time (ncu: serialised, cold cache; compare shares)")
    print("## by kernel")
    for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / 1e3:9.3f} ms {100 * v[1] / tot:5.1f}%  x{v[0]:<5d} avg {v[1] / v[0]:9.1f} us  {k}")
    if dram:
        print("## DRAM traffic by kernel (dram__bytes_read.sum + dram__bytes_write.sum, whole list and per launch)")
        for k, v in sorted(dram.items(), key=lambda kv: -kv[1]):
            n_k = fam[k][0] if k in fam else 0
            print(f"{v / 1e9:9.3f} GB  x{n_k:<5d} avg {v / max(n_k, 1) / 1e6:9.3f} MB  {k}")
        if len(sys.argv) > 2:    # machine-readable copy for bench.py's roofline.traffic
            import json
            k = max((k for k in dram if "gemm2_kernel" in k), key=lambda k: dram[k], default=None)
            if k is not None:
                json.dump({"kernel": k, "launches": fam[k][0], "dram_bytes_total": dram[k],
                           "dram_bytes_per_launch": dram[k] / max(fam[k][0], 1), "source": path},
                          open(sys.argv[2], "w"))
    print("## by kernel and grid (top 60)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
        print(f"{v[1] / 1e3:9.3f} ms {100 * v[1] / tot:5.1f}%  x{v[0]:<5d} avg {v[1] / v[0]:9.1f} us  {k[0]} grid {k[1]} block {k[2]}")


if __name__ == "__main__":
    main(sys.argv[1])
