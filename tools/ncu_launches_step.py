"""Aggregate one bench step out of an ncu launch list (gpu__time_duration + dram bytes per launch).
  python tools/ncu_launches_step.py <launches.csv[.gz]> <out.txt> <out.json> "<command>"
A step = the launches from one conv_in kernel (first kernel of a UNet evaluation after the time
embedding) to the next; the LAST complete step in the capture is used."""
import collections
import csv
import gzip
import json
import sys


def main(path, out_txt, out_json, cmd):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        lines = [l for l in f if not l.startswith("==")]
    launches = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = launches.setdefault(r["ID"], {"name": r["Kernel Name"].split("(")[0]})
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)  # us
        else:
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        d[r["Metric Name"]] = v
    L = list(launches.values())
    starts = [i for i, d in enumerate(L) if "conv_in" in d["name"]]
    assert len(starts) >= 2, "need two UNet evaluations in the capture"
    a, b = starts[-2], starts[-1]
    step = L[a:b]
    agg = collections.defaultdict(lambda: [0.0, 0, 0.0, 0.0])
    for d in step:
        k = d["name"].replace("void ", "").replace("pf::", "")[:64]
        e = agg[k]
        e[0] += d.get("gpu__time_duration.sum", 0.0)
        e[1] += 1
        e[2] += d.get("dram__bytes_read.sum", 0.0)
        e[3] += d.get("dram__bytes_write.sum", 0.0)
    tot = sum(e[0] for e in agg.values())
    gemm = [e for k, e in agg.items() if k.startswith("gemm_tc")]
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import kernel_source_hash
    summary = {
        "command": cmd,
        "kernel_source_hash": kernel_source_hash(),  # bench.py drops the traffic figure when the sources changed
        "launches_per_step": len(step),
        "step_us_sum_of_kernels_under_ncu": tot,
        "gemm_launches": sum(e[1] for e in gemm),
        "gemm_us": sum(e[0] for e in gemm),
        "gemm_share": sum(e[0] for e in gemm) / tot,
        "gemm_dram_bytes_read": sum(e[2] for e in gemm),
        "gemm_dram_bytes_write": sum(e[3] for e in gemm),
        "step_dram_bytes": sum(e[2] + e[3] for e in agg.values()),
    }
    with open(out_txt, "w") as f:
        f.write(cmd + "\n")
        f.write(f"(one bench step = launches [{a}, {b}) of the capture, {len(step)} launches; ncu times are "
                "cold-cache and serialised -> compare SHARES with the live CUDA-event breakdown)\n\n")
        f.write(f"{'us':>10} {'count':>6} {'share':>6} {'dram rd MB':>11} {'dram wr MB':>11}  kernel\n")
        for k, (t, c, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write(f"{t:10.1f} {c:6d} {100 * t / tot:5.1f}% {rd / 1e6:11.1f} {wr / 1e6:11.1f}  {k}\n")
        f.write(f"\nstep total {tot / 1e3:.2f} ms; tcgen05 GEMM launches: {summary['gemm_launches']}, "
                f"{summary['gemm_us'] / 1e3:.2f} ms = {100 * summary['gemm_share']:.1f} % of the step, DRAM traffic "
                f"{(summary['gemm_dram_bytes_read'] + summary['gemm_dram_bytes_write']) / 1e9:.2f} GB per step\n")
    with open(out_json, "w") as f:
        json.dump(summary, f, indent=1)
    print(open(out_txt).read())


if __name__ == "__main__":
    main(*sys.argv[1:5])
