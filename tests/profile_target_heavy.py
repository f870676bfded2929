"""Short single-GPU profiling target for the FULL step-VM variant (never a bench number): hash chain, then Pedersen chain."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import acvm_b200
from acvm_b200 import acir_builder as ab
ctx = acvm_b200.Context(0)
ctx.set_option("pedersen_unpinned", 1)   # structure/cost measurement only: values are not barretenberg's
rng = np.random.default_rng(1)
for name, (data, inputs, nw), byte_valued in (("hash", ab.hash_chain_circuit(256), True), ("pedersen", ab.pedersen_chain_circuit(16), False)):
    batch = 4096
    arr = np.zeros((batch, len(inputs), 32), dtype=np.uint8)
    if byte_valued:
        arr[:, :, 31] = rng.integers(0, 256, size=(batch, len(inputs)), dtype=np.uint8)
    else:
        arr[:] = rng.integers(0, 256, size=arr.shape, dtype=np.uint8)
        arr[:, :, 0] &= 0x1F
    circ = acvm_b200.CompiledCircuit(ctx, data, inputs)
    b = acvm_b200.DeviceBatch(circ, batch)
    b.stage_inputs(0, arr.tobytes())
    print(name, b.run_staged(0), sum(s.status == "Solved" for s in b.status()))
    b.close(); circ.close()
