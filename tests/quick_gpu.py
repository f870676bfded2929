import sys, time, json
sys.path.insert(0, '.')
import acvm_b200
from acvm_b200 import acir_builder as ab
ctx = acvm_b200.Context(0)
print("device", ctx.device_name())
print("imad", json.dumps(ctx.imad_microbench()))
t = time.time(); data, inputs, nw = ab.synthetic_arith_circuit(1 << 16); print("gen", time.time() - t)
for S, T in [(16, 8), (16, 4), (8, 16), (32, 4), (4, 32)]:
    ctx.set_option("S", S); ctx.set_option("T", T)
    t = time.time(); circ = acvm_b200.CompiledCircuit(ctx, data, inputs); tc = time.time() - t
    for batch in (1024, 8192):
        b = acvm_b200.DeviceBatch(circ, batch)
        inp = ab.synthetic_inputs(16) * (batch // 16)
        b.upload(inp)
        ms = [b.run() for _ in range(3)]
        ok = all(s.status == "Solved" for s in b.status())
        gi = (1 << 16) * batch
        print(f"S={S} T={T} batch={batch} compile={tc:.2f}s steps={circ.info['n_steps']} kernel_ms={ms} ok={ok} gate-inst/s={gi/(min(ms)*1e-3):.3e} imad/s={circ.info['dev_imad']*batch/(min(ms)*1e-3):.3e}")
        b.close()
    circ.close()
