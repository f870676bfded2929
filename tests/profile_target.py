"""Short, single-GPU profiling target for ncu (never a bench number): 2^16 gates, one full-occupancy sub-batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acvm_b200
from acvm_b200 import acir_builder as ab
gates = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4736
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = acvm_b200.Context(0)
for kv in filter(None, os.environ.get("ACVMB_OPTS", "").split(",")):
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
data, inputs, _ = ab.synthetic_arith_circuit(gates)
circ = acvm_b200.CompiledCircuit(ctx, data, inputs)
b = acvm_b200.DeviceBatch(circ, batch)
b.stage_inputs(0, ab.synthetic_inputs(16) * (batch // 16))
for _ in range(reps):
    print(b.run_staged(0))
print(circ.info)
