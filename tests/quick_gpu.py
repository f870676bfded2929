"""Ad-hoc GPU tuning sweep (not a test, not a bench line): multiplier representation x tile shape on a 2^16-gate circuit."""
import sys, time, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acvm_b200
from acvm_b200 import acir_builder as ab
ctx = acvm_b200.Context(0)
print("device", ctx.device_name())
print("imad", json.dumps(ctx.imad_microbench()))
for coeffs in ("dense", "noir-like"):
    data, inputs, nw = ab.synthetic_arith_circuit(1 << 16, coeffs=coeffs)
    for S, T, a29 in [(16, 8, 0), (16, 8, 1), (16, 4, 1), (32, 4, 1), (8, 16, 1)]:
        ctx.set_option("S", S); ctx.set_option("T", T); ctx.set_option("arith29", a29)
        circ = acvm_b200.CompiledCircuit(ctx, data, inputs)
        for batch in (4736, 8192):
            b = acvm_b200.DeviceBatch(circ, batch)
            b.stage_inputs(0, ab.synthetic_inputs(16) * (batch // 16))
            ms = [b.run_staged(0)[1] for _ in range(3)]
            ok = all(s.status == "Solved" for s in b.status())
            gi = (1 << 16) * batch
            print(f"{coeffs} S={S} T={T} arith29={a29} batch={batch} kernel_ms={min(ms):.3f} ok={ok} gate-inst/s={gi/(min(ms)*1e-3):.3e}", flush=True)
            b.close()
        circ.close()
