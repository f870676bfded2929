#!/usr/bin/env python3
"""Secondary measurements: BASELINE.json configs 2-4 (and the alternative operand / coefficient modes of config 1) with the
inputs resident in HBM.  These are NOT the bench.py headline line; each JSON line states its exact (reduced) size.

    python bench_configs.py [--which pedersen,hash,mixed,modes] [--reps 3]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import acvm_b200  # noqa: E402
from acvm_b200 import acir_builder as ab  # noqa: E402


def rand_inputs(batch, n_inputs, byte_valued=False, seed=1):
    rng = np.random.default_rng(seed)
    if byte_valued:
        out = np.zeros((batch, n_inputs, 32), dtype=np.uint8)
        out[:, :, 31] = rng.integers(0, 256, size=(batch, n_inputs), dtype=np.uint8)
    else:
        out = rng.integers(0, 256, size=(batch, n_inputs, 32), dtype=np.uint8)
        out[:, :, 0] &= 0x1F   # < 2^253 < p
    return out.tobytes()


def measure(ctx, name, data, inputs, batch, inp, reps, extra):
    t0 = time.time()
    circ = acvm_b200.CompiledCircuit(ctx, data, inputs)
    tc = time.time() - t0
    b = acvm_b200.DeviceBatch(circ, batch)
    b.stage_inputs(0, inp)
    b.run_staged(0)
    ms = [b.run_staged(0)[0] for _ in range(reps)]
    st = b.status()
    solved = sum(s.status == "Solved" for s in st)
    best = min(ms)
    info = circ.info
    line = dict(config=name, batch=batch, ms=best, witnesses_per_s=batch / (best * 1e-3), solved=solved, plan_compile_s=tc,
                n_opcodes=info["n_opcodes"], n_steps=info["n_steps"], S=info["S"], n_micro_ops=info["n_micro_ops"],
                alg_GBps=info["alg_bytes"] * batch / (best * 1e-3) / 1e9, **extra)
    print(json.dumps(line), flush=True)
    b.close()
    circ.close()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="modes,hash,pedersen,mixed")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--pedersen-calls", type=int, default=256)
    ap.add_argument("--hash-calls", type=int, default=1 << 12)
    ap.add_argument("--mixed-ops", type=int, default=1 << 16)
    ap.add_argument("--S", type=int, default=0)
    ap.add_argument("--mixed-batch", type=int, default=8192)
    ap.add_argument("--opt", action="append", default=[], help="context option key=value (e.g. split_curve=0, temp_pool=65536)")
    args = ap.parse_args()
    ctx = acvm_b200.Context(0)
    ctx.set_option("pedersen_unpinned", 1)   # configs 2 and 4 measure the Pedersen kernel (structure-identical, parity unpinned)
    if args.S:
        ctx.set_option("S", args.S)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    which = args.which.split(",")
    if "modes" in which:   # config 1 shape at 2^16 gates: operand locality x coefficient distribution
        for mode in ("local", "global"):
            for coeffs in ("dense", "noir-like"):
                data, inputs, _ = ab.synthetic_arith_circuit(1 << 16, mode=mode, coeffs=coeffs)
                batch = 8192
                line = measure(ctx, f"arith_2^16_{mode}_{coeffs}", data, inputs, batch, ab.synthetic_inputs(64) * (batch // 64), args.reps,
                               dict(gates=1 << 16))
    if "hash" in which:    # config 3
        data, inputs, nw = ab.hash_chain_circuit(args.hash_calls)
        batch = 4096
        line = measure(ctx, "hash_chain_sha256_keccak256", data, inputs, batch, rand_inputs(batch, len(inputs), byte_valued=True), args.reps,
                       dict(hash_calls=args.hash_calls, bytes_per_call=64))
        print(json.dumps(dict(config="hash_chain derived", hashes_per_s=args.hash_calls * batch / (line["ms"] * 1e-3))), flush=True)
    if "pedersen" in which:  # config 2
        data, inputs, nw = ab.pedersen_chain_circuit(args.pedersen_calls)
        batch = 4096
        line = measure(ctx, "pedersen_chain", data, inputs, batch, rand_inputs(batch, len(inputs)), args.reps,
                       dict(pedersen_calls=args.pedersen_calls, inputs_per_call=2))
        print(json.dumps(dict(config="pedersen_chain derived", pedersen_per_s=args.pedersen_calls * batch / (line["ms"] * 1e-3))), flush=True)
    if "mixed" in which:   # config 4
        data, inputs, nw, counts = ab.mixed_circuit(args.mixed_ops)
        batch = args.mixed_batch
        measure(ctx, "mixed_93_4_2_1", data, inputs, batch, ab.synthetic_inputs(64, seed_id=4) * (batch // 64), args.reps,
                dict(ops=args.mixed_ops, counts=counts, n_witnesses=nw))


if __name__ == "__main__":
    main()
