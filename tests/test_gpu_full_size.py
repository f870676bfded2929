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


# ---- BASELINE.json configs 2-4 at their stated sizes (ACVMB_FULL_SCALE=k shrinks every length by k for quick runs) ----
def _sampled_tail_parity(ctx, data, inputs, nw, batch, inp, n_tail, sample):
    """Solve `batch` instances through the C ABI (sub-batches inside the library), bring back only the last `n_tail`
    witnesses -- the end of every instance's dependency chain -- and compare sampled instances bit-for-bit with the C++
    CPU restatement of the reference algorithm."""
    circ = acvm_b200.CompiledCircuit(ctx, data, inputs)
    assert circ.num_witnesses >= nw
    tail = list(range(nw - n_tail, nw))       # the last witnesses the circuit defines
    nw = circ.num_witnesses
    out, st = circ.solve_batch(inp, batch, out_ids=tail)
    assert all(s.status == "Solved" for s in st)
    n_in = len(inputs)
    sinp = b"".join(inp[i * n_in * 32:(i + 1) * n_in * 32] for i in sample)
    cref.build()
    oc = acir.decode_circuit(data)
    res, ow, op = cref.solve_batch(oc, inputs, sinp, len(sample), nw, threads=min(len(sample), os.cpu_count() or 1), want_witness=True)
    assert (res[:, 0] == 0).all()
    for k, i in enumerate(sample):
        got = [int.from_bytes(out[(i * n_tail + j) * 32:(i * n_tail + j + 1) * 32], "big") for j in range(n_tail)]
        exp = [int.from_bytes(ow[k, w].tobytes(), "little") for w in tail]
        assert got == exp, f"instance {i}"
    info = circ.info
    circ.close()
    return info


def _scale():
    return int(os.environ.get("ACVMB_FULL_SCALE", 1))


def test_config3_full_size_sampled_parity(ctx):
    """configs[3]: 2^14 chained hash calls alternating SHA256 / Keccak256 over 64 byte-witnesses, batch 4096."""
    n, batch = (1 << 14) // _scale(), 4096
    data, inputs, nw = ab.hash_chain_circuit(n)
    rng = np.random.default_rng(3)
    arr = np.zeros((batch, len(inputs), 32), dtype=np.uint8)
    arr[:, :, 31] = rng.integers(0, 256, size=(batch, len(inputs)), dtype=np.uint8)
    info = _sampled_tail_parity(ctx, data, inputs, nw, batch, arr.tobytes(), 32, [0, 1, batch // 2, batch - 1])
    assert info["n_hash"] == n


def test_config2_full_size_sampled_parity(pctx):
    """configs[2]: 2^16 chained Pedersen{[prev.x, fresh_i], domain 0} calls, batch 4096 -- on the opt-in kernel: the values are
    those of the structure restatement (oracle/pedersen.py, parity with barretenberg unpinned), checked against its C++ twin."""
    n, batch = (1 << 16) // _scale(), 4096
    data, inputs, nw = ab.pedersen_chain_circuit(n)
    rng = np.random.default_rng(2)
    arr = rng.integers(0, 256, size=(batch, len(inputs), 32), dtype=np.uint8)
    arr[:, :, 0] &= 0x1F
    info = _sampled_tail_parity(pctx, data, inputs, nw, batch, arr.tobytes(), 2, [0, batch // 2 + 1, batch - 1])
    assert info["n_curve"] == n


def test_config4_full_size_sampled_parity(pctx):
    """configs[4] (one GPU's share): 2^20 mixed opcodes -- 93% dense arithmetic, 4% RANGE/AND/XOR(32), 2% SHA256/Keccak256 over
    64 bytes, 1% Pedersen(2)/FixedBaseScalarMul -- batch 8192, in sub-batches."""
    n, batch = (1 << 20) // _scale(), 8192
    data, inputs, nw, counts = ab.mixed_circuit(n)
    inp = ab.synthetic_inputs(batch, seed_id=4)
    info = _sampled_tail_parity(pctx, data, inputs, nw, batch, inp, 64, [0, 1, batch // 3, batch // 2, batch - 2, batch - 1])
    assert info["n_opcodes"] == n and info["n_curve"] == counts["pedersen"] + counts["fixed_base"]


@pytest.mark.parametrize("mode,coeffs", [("global", "noir-like"), ("local", "noir-like"), ("global", "dense")])
def test_config1_other_modes_full_size_sampled_parity(ctx, mode, coeffs):
    """The operand / coefficient variants of configs[1] that DESIGN.md also reports, at the full 2^20 gates (one tile-aligned
    slice of the batch is enough here: the full batch runs in bench.py and in the local/dense test above)."""
    gates = (1 << 20) // _scale()
    batch = 1184
    data, inputs, nw = ab.synthetic_arith_circuit(gates, seed_id=1, mode=mode, coeffs=coeffs)
    inp = ab.synthetic_inputs(batch, seed_id=1)
    _sampled_tail_parity(ctx, data, inputs, nw, batch, inp, 32, [0, 7, batch // 2, batch - 1])
