#!/usr/bin/env python3
"""Warp-stall sampling of one kernel from an .ncu-rep (source page, SASS view): totals per stall reason, with and without the
step-barrier wait, the instruction mix, and the hottest instructions.

    python profiles/summarize_stalls.py gpurun_out/prof.ncu-rep [units_per_launch] > profiles/<name>.txt
"""
import collections
import csv
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    print(rows[0][0], rows[0][1])
    hdr = rows[1]
    ia, ii, isamp, isrc = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = [r for r in rows[2:] if len(r) > isamp and r[ia].startswith("0x")]
    tot_i = sum(int(r[ii] or 0) for r in data)
    tot_s = sum(int(r[isamp] or 0) for r in data)
    print(f"warp instructions executed: {tot_i}   stall samples: {tot_s}" + (f"   instructions per unit: {tot_i / units:.1f}" if units else ""))
    by = collections.Counter()
    for r in data:
        for i in stall:
            by[hdr[i]] += int(r[i] or 0)
    bar = by["stall_barrier"]
    print("\nstall reason            samples   share   share without the step barrier")
    for k, v in by.most_common():
        if v:
            print(f"  {k:20s} {v:9d}  {100 * v / tot_s:5.1f}%   " + ("" if k == "stall_barrier" else f"{100 * v / max(1, tot_s - bar):5.1f}%"))
    mix = collections.Counter()
    for r in data:
        op = re.sub(r"^@!?U?P\d+\s+", "", r[isrc].strip()).split()[0]
        mix[op] += int(r[ii] or 0)
    print("\ninstruction mix (executed warp instructions)" + (" per unit" if units else ""))
    for op, n in mix.most_common(16):
        print(f"  {op:24s} {100 * n / tot_i:5.1f}%" + (f"  {n / units:9.1f}" if units else ""))
    print("\nhottest instructions by samples")
    for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:12]:
        print(f"  {int(r[isamp] or 0):8d}  {r[isrc].strip()[:70]}")


if __name__ == "__main__":
    main()
