#!/usr/bin/env python3
"""Instruction census of the sm_100a cubins inside acvm_b200/libacvm_b200.so (cuobjdump -sass; runs without a GPU).

    python profiles/sass_census.py > profiles/r2_sass_census.txt

Per kernel: TMA bulk copies (UBLKCP), mbarrier ops (SYNCS), wide integer multiply-adds with / without carry
(IMAD.WIDE.U32.X / IMAD.WIDE.U32), other IMADs, tensor-core ops (none expected: there is no dense contraction on this path).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "acvm_b200", "libacvm_b200.so")
KEYS = ["UBLKCP", "SYNCS", "IMAD.WIDE.U32.X", "IMAD.WIDE.U32", "IMAD.HI.U32", "IMAD", "IADD3.X", "IADD3", "LDG", "STG", "LDS", "STS", "BAR",
        "HMMA", "IMMA", "UTCMMA", "UTCHMMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    arch = set(re.findall(r"arch = (sm_\w+)", out))
    print(f"library: {os.path.relpath(LIB, ROOT)}   cubin architectures: {sorted(arch)}")
    fn = None
    counts = collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            counts[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            op = m.group(1)
            counts[fn]["total"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    counts[fn][k] += 1
                    break
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print(f"{'kernel':78s} {'total':>7s} " + " ".join(f"{k:>16s}" for k in KEYS[:9]) + "  tensor-core ops")
    for (fn, c), name in zip(counts.items(), demangle):
        name = name.replace("acvmb::", "").replace("(acvmb::VmArgs)", "")[:78]
        tc = sum(c[k] for k in ("HMMA", "IMMA", "UTCMMA", "UTCHMMA"))
        print(f"{name:78s} {c['total']:7d} " + " ".join(f"{c[k]:16d}" for k in KEYS[:9]) + f"  {tc}")
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print("\nwhole library: " + ", ".join(f"{k} {tot[k]}" for k in KEYS if tot[k]) +
          f"; tensor-core ops {sum(tot[k] for k in ('HMMA', 'IMMA', 'UTCMMA', 'UTCHMMA'))}")
    print("\nPer Montgomery product (fr.cuh mont_dot_fn<1>, measured on the frmul_bench_kernel<0> loop body: 2 products per iteration):")
    for (fn, c), name in zip(counts.items(), demangle):
        if "frmul_bench_kernel<0>" in name:
            print(f"  {name}: IMAD.WIDE.U32.X {c['IMAD.WIDE.U32.X']}, IMAD.WIDE.U32 {c['IMAD.WIDE.U32']}, other IMAD {c['IMAD'] + c['IMAD.HI.U32']} "
                  f"in the whole kernel (loop body = 2 products + loop control)")


if __name__ == "__main__":
    sys.exit(main())
