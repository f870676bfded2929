#!/usr/bin/env python3
"""Bare device->host ceiling of the box: concurrent cudaMemcpyAsync D2H of 1 GiB chunks from 1/2/4/8 GPUs into pinned host
memory, one host thread per GPU, each thread bound to the CPUs NVML reports as local to its GPU before it allocates (so the
pinned pages are first-touched next to the GPU's PCIe root).  Prints one JSON line per GPU count: per-GPU and aggregate GB/s.
The end-to-end path of bench.py (full dense witness map back to the host) cannot exceed these numbers.

    python tools/d2h_ceiling.py [--gib 1] [--reps 8] [--no-bind]
"""
import argparse
import json
import os
import threading
import time

import torch


def bind(gpu):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)   # per-thread on Linux
            return len(cpus)
    except Exception:
        pass
    return 0


def worker(gpu, nbytes, reps, do_bind, barrier, out):
    ncpu = bind(gpu) if do_bind else 0
    torch.cuda.set_device(gpu)
    src = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{gpu}")
    dst = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(gpu)
    barrier.wait()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    e1.record()
    e1.synchronize()
    out[gpu] = (nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9, ncpu)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=8)
    ap.add_argument("--no-bind", action="store_true")
    args = ap.parse_args()
    n_gpu = torch.cuda.device_count()
    nbytes = int(args.gib * 2**30)
    for n in (1, 2, 4, 8):
        if n > n_gpu:
            break
        out = {}
        barrier = threading.Barrier(n)
        th = [threading.Thread(target=worker, args=(g, nbytes, args.reps, not args.no_bind, barrier, out)) for g in range(n)]
        t0 = time.time()
        for t in th:
            t.start()
        for t in th:
            t.join()
        per = [round(out[g][0], 2) for g in range(n)]
        print(json.dumps({"gpus": n, "bound_to_local_cpus": not args.no_bind, "local_cpus": [out[g][1] for g in range(n)],
                          "per_gpu_GBps": per, "aggregate_GBps": round(sum(per), 2), "chunk_GiB": args.gib, "reps": args.reps,
                          "wall_s": round(time.time() - t0, 1)}), flush=True)


if __name__ == "__main__":
    main()
