"""Summarise ncu outputs into small text files under profiles/.
  python tools/ncu_summarize.py launches <launches.csv> <out.txt> "<command>"
  python tools/ncu_summarize.py full <report.ncu-rep> <out.txt> "<command>"
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]


def launches(path, out, cmd):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0.0, 0])
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        n += 1
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        k = row["Kernel Name"].split("(")[0][:70]
        agg[k][0] += v
        agg[k][1] += 1
    tot = sum(v[0] for v in agg.values())
    with open(out, "w") as f:
        f.write(cmd + "\n")
        f.write(f"({n} launches captured; per-launch times are cold-cache and serialised -> compare SHARES)\n\n")
        f.write(f"{'total ms':>10} {'count':>6} {'share':>6}  kernel\n")
        for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write(f"{t / 1e3:10.2f} {c:6d} {100 * t / tot:5.1f}%  {k}\n")


def full(rep, out, cmd):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(cmd + "\n\n")
        for r in rows[2:]:
            f.write(r[idx["Kernel Name"]] + "\n")
            for k in KEYS:
                if k in idx:
                    f.write(f"    {k:75s} {r[idx[k]]:>16s} {units[idx[k]]}\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4])
