#!/usr/bin/env python3
"""Per-kernel totals / shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv
import sys


def main():
    path = sys.argv[1]
    cmd = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = {}
    for r in rows[1:]:
        agg.setdefault(r[ki].split("(")[0], []).append(float(r[vi]))
    tot = sum(sum(v) for k, v in agg.items() if "bench_kernel" not in k)
    print(f"ncu --metrics gpu__time_duration.sum --clock-control none : {cmd}")
    print("(per-launch times are cold-cache and serialised: compare SHARES; *_bench_kernel are the roofline micro-benchmarks, outside the timed steps)")
    for k, v in agg.items():
        share = f"{sum(v) / tot * 100:6.2f}%" if "bench_kernel" not in k else "   n/a"
        print(f"{k:60s} launches={len(v):3d} total_ms={sum(v) / 1e6:10.3f} avg_ms={sum(v) / len(v) / 1e6:9.4f} share_of_step={share}")


if __name__ == "__main__":
    main()
