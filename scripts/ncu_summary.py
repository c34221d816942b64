"""Dev tool (runs here, no GPU): summarise an .ncu-rep (ncu --set full) — one block of the metrics DESIGN.md quotes per
captured launch.  python scripts/ncu_summary.py <file.ncu-rep> [title] > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    "launch__grid_size", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
]


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else rep
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, body = rows[0], rows[1], rows[2:]
    print(f"# {title}")
    for r in body:
        d = dict(zip(head, r))
        u = dict(zip(head, units))
        print(f"--- {d['Kernel Name'][:48]} id {d['ID']}  grid {d['Grid Size']} block {d['Block Size']}")
        for w in WANT:
            for k in head:
                if k == w or k.endswith("." + w):
                    if d.get(k, "") != "":
                        print(f"   {w:<70s} {d[k]} {u[k]}")
                    break


if __name__ == "__main__":
    main()
