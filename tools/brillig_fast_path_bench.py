#!/usr/bin/env python3
"""stdlib-shaped circuit: N `bytecode: vec![Stop]` Brillig constant loads (stdlib/src/blackbox_fallbacks/uint.rs:51-63,85),
each feeding an arithmetic gate, every fourth result decomposed with ToLeRadix and divided with Quotient -- solved with the
plan-time lowering of straight-line Brillig on (default) and off (every Brillig opcode a host segment).

    python tools/brillig_fast_path_bench.py [N] [batch]   > profiles/r2_brillig_fast_path.txt
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acvm_b200
from acvm_b200 import acir_builder as ab

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
BATCH = int(sys.argv[2]) if len(sys.argv) > 2 else 1024


def circuit(n):
    b = ab.CircuitBuilder()
    nxt = 3
    acc = 1
    for i in range(n):
        c = nxt
        b.brillig([("Single", ([], [], (1 << 32) + i))], [("Simple", c)], [dict(op="Stop")])          # load_constant
        s = nxt + 1
        b.arithmetic([], [(1, acc), (1, c), (ab.P - 1, s)], 0)                                       # s = acc + const
        nxt += 2
        if i % 4 == 3:
            bits = list(range(nxt, nxt + 8))
            b.directive_to_le_radix(ab.wexpr(2), bits, 2)                                                     # 8 low bits of input w2
            q, r = nxt + 8, nxt + 9
            b.directive_quotient(ab.wexpr(s), ab.wexpr(c), q, r)
            nxt += 10
        acc = s
    return b.to_bytes(), [1, 2]


def main():
    data, inputs = circuit(N)
    rnd = ab.synthetic_inputs(BATCH, n_inputs=1, seed_id=7)   # w1: a field element; w2: a byte (it is decomposed into 8 bits)
    inp = b"".join(rnd[32 * i:32 * i + 32] + ((37 * i + 11) % 256).to_bytes(32, "big") for i in range(BATCH))
    ref = None
    for dev in (1, 0):
        ctx = acvm_b200.Context(0)
        ctx.set_option("device_brillig", dev)
        t0 = time.time()
        circ = acvm_b200.CompiledCircuit(ctx, data, inputs)
        t_plan = time.time() - t0
        info = circ.info
        tail = list(range(circ.num_witnesses - 16, circ.num_witnesses))
        circ.solve_batch(inp, BATCH, out_ids=tail)
        t0 = time.perf_counter()
        out, st = circ.solve_batch(inp, BATCH, out_ids=tail)
        dt = time.perf_counter() - t0
        assert all(s.status == "Solved" for s in st)
        ref = ref or out
        assert out == ref, "both lowerings must give the same witnesses"
        print(json.dumps(dict(device_brillig=dev, brillig_opcodes=info["n_brillig"], lowered_to_gates=info["n_brillig_device"],
                              segments=info["n_segments"], host_segments=info["n_host_segments"], n_steps=info["n_steps"],
                              plan_compile_s=round(t_plan, 2), batch=BATCH, solve_ms=round(1e3 * dt, 1),
                              kernel_launches=circ.run_info()["kernel_launches"])), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
