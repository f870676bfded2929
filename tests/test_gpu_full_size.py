"""Full-size parity (BASELINE.json configs[1]: 2^20 arithmetic gates, batch 8192): every instance must solve (1/16 of the
opcodes are all-known CHECK gates), and sampled instances are compared bit-for-bit with the C++ CPU restatement of the
reference algorithm (oracle/ref_solver.cpp) on the LAST witnesses of the dependency chain -- any wrong gate anywhere in an
instance's 2^20-long chain changes them.  ACVMB_FULL_GATES / ACVMB_FULL_BATCH shrink it for quick runs."""
import os

import numpy as np
import pytest

import acvm_b200
from acvm_b200 import acir_builder as ab
from oracle import acir, cref

pytestmark = pytest.mark.gpu


def test_config1_full_size_sampled_parity(ctx):
    gates = int(os.environ.get("ACVMB_FULL_GATES", 1 << 20))
    batch = int(os.environ.get("ACVMB_FULL_BATCH", 8192))
    cache = f"/tmp/acvmb_circuit_{gates}_local_dense.bin"
    if os.path.exists(cache):
        data, inputs = open(cache, "rb").read(), list(range(ab.N_INPUTS))
    else:
        data, inputs, _ = ab.synthetic_arith_circuit(gates)
        try:
            open(cache, "wb").write(data)
        except OSError:
            pass
    circ = acvm_b200.CompiledCircuit(ctx, data, inputs)
    nw = circ.num_witnesses
    tail = list(range(nw - 32, nw))
    inp = ab.synthetic_inputs(batch)
    out, st = circ.solve_batch(inp, batch, out_ids=tail)           # sub-batches internally; only the tail comes back
    assert all(s.status == "Solved" for s in st)
    ri = circ.run_info()
    assert ri["kernel_launches"] >= 3 * ri["n_subbatches"]
    # sampled instances (first / last of every sub-batch region and a few in between) against the CPU restatement
    sample = sorted({0, 1, batch // 3, batch // 2, batch // 2 + 1, batch - 2, batch - 1, 7 * batch // 8})
    oc = acir.decode_circuit(data)
    sinp = b"".join(inp[i * 256:(i + 1) * 256] for i in sample)
    cref.build()
    res, ow, op = cref.solve_batch(oc, inputs, sinp, len(sample), nw, threads=min(len(sample), os.cpu_count() or 1), want_witness=True)
    assert (res[:, 0] == 0).all()
    for k, i in enumerate(sample):
        got = [int.from_bytes(out[(i * 32 + j) * 32:(i * 32 + j + 1) * 32], "big") for j in range(32)]
        exp = [int.from_bytes(ow[k, w].tobytes(), "little") for w in tail]
        assert got == exp, f"instance {i}"
    circ.close()


def test_config1_replicated_instances_checksum(ctx):
    """Size-independent property at a full sub-batch: the SAME instance placed in every lane of every tile must produce the
    same witness map everywhere (device-side per-instance checksums, no D2H of the 150 GB of columns), and that checksum
    must equal the one of the CPU restatement's witness map."""
    gates = int(os.environ.get("ACVMB_FULL_GATES", 1 << 20))
    n = int(os.environ.get("ACVMB_FULL_SUBBATCH", 4736))
    cache = f"/tmp/acvmb_circuit_{gates}_local_dense.bin"
    if os.path.exists(cache):
        data, inputs = open(cache, "rb").read(), list(range(ab.N_INPUTS))
    else:
        data, inputs, _ = ab.synthetic_arith_circuit(gates)
    circ = acvm_b200.CompiledCircuit(ctx, data, inputs)
    one = ab.synthetic_inputs(1, first_instance=12345)
    b = acvm_b200.DeviceBatch(circ, n)
    b.stage_inputs(0, one * n)
    b.run_staged(0)
    assert all(s.status == "Solved" for s in b.status())
    sums = b.checksums()
    assert len(set(sums)) == 1, "some lane computed a different witness map for identical inputs"
    oc = acir.decode_circuit(data)
    res, ow, op = cref.solve_batch(oc, inputs, one, 1, circ.num_witnesses, threads=1, want_witness=True)
    assert res[0, 0] == 0
    # vectorised host definition of the checksum (acvm_b200.witness_checksum) over ~10^6 witnesses
    present = np.nonzero(op[0])[0].astype(np.uint64)
    limbs32 = ow[0].view("<u4").reshape(-1, 8)[present.astype(np.int64)].astype(np.uint64)
    h = (present + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
    for k in range(8):
        h = (h ^ limbs32[:, k]) * np.uint64(0x100000001B3)
    assert int(h.sum(dtype=np.uint64)) == sums[0]
    b.close()
    circ.close()
