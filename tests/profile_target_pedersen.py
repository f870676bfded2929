"""Short single-GPU profiling target (never a bench number): one Pedersen-chain launch of the FULL step-VM variant."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import acvm_b200
from acvm_b200 import acir_builder as ab
ctx = acvm_b200.Context(0)
ctx.set_option("pedersen_unpinned", 1)   # structure/cost measurement only: values are not barretenberg's
for kv in filter(None, os.environ.get("ACVMB_OPTS", "").split(",")):
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
rng = np.random.default_rng(1)
data, inputs, nw = ab.pedersen_chain_circuit(int(sys.argv[1]) if len(sys.argv) > 1 else 32)
batch = 4096
arr = rng.integers(0, 256, size=(batch, len(inputs), 32), dtype=np.uint8)
arr[:, :, 0] &= 0x1F
circ = acvm_b200.CompiledCircuit(ctx, data, inputs)
print(circ.info)
b = acvm_b200.DeviceBatch(circ, batch)
b.stage_inputs(0, arr.tobytes())
print("pedersen", b.run_staged(0), sum(s.status == "Solved" for s in b.status()))
