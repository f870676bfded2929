"""Ad-hoc GPU tuning sweep (not a test, not a bench line): tile shape / staging depth on a 2^16-gate circuit."""
import sys, time, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acvm_b200
from acvm_b200 import acir_builder as ab
ctx = acvm_b200.Context(0)
print("device", ctx.device_name())
data, inputs, nw = ab.synthetic_arith_circuit(1 << 16)
combos = [(16, 8, 4, 2), (16, 8, 2, 2), (16, 8, 8, 2), (16, 8, 4, 1), (16, 8, 4, 4), (16, 8, 2, 4), (16, 4, 4, 2), (16, 16, 4, 2), (32, 4, 4, 2)]
if len(sys.argv) > 1:
    combos = [tuple(int(v) for v in c.split(',')) for c in sys.argv[1:]]
for S, T, nst, chunk in combos:
    ctx.set_option("S", S); ctx.set_option("T", T); ctx.set_option("n_stage", nst); ctx.set_option("chunk_steps", chunk)
    try:
        circ = acvm_b200.CompiledCircuit(ctx, data, inputs)
        for batch in (4736,):
            b = acvm_b200.DeviceBatch(circ, batch)
            b.stage_inputs(0, ab.synthetic_inputs(16) * (batch // 16))
            ms = [b.run_staged(0)[1] for _ in range(3)]
            ok = all(s.status == "Solved" for s in b.status())
            gi = (1 << 16) * batch
            print(f"S={S} T={T} n_stage={nst} chunk={chunk} batch={batch} kernel_ms={min(ms):.3f} ok={ok} gate-inst/s={gi/(min(ms)*1e-3):.3e}", flush=True)
            b.close()
        circ.close()
    except Exception as e:
        print(f"S={S} T={T} n_stage={nst} chunk={chunk}: {e}", flush=True)
