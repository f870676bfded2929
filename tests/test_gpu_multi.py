"""Multi-device context behind the C ABI (acvmb_ctx_create_multi): the plan is compiled once and broadcast to
every device, acvmb_solve_batch shards the batch contiguously and fills ONE host buffer.  With a single visible GPU the same
calls run with n = 1 (the sharding code is then the identity); on a multi-GPU box every device takes part."""
import pytest

import acvm_b200
from acvm_b200 import acir_builder as ab
from conftest import inputs_to_dicts, witness_rows
from oracle import acir, pwg

pytestmark = pytest.mark.gpu


def _devices():
    import torch
    return list(range(min(torch.cuda.device_count(), 8)))


def test_multi_device_context_matches_single_device_and_oracle():
    devs = _devices()
    mctx = acvm_b200.Context(devs)
    sctx = acvm_b200.Context(devs[-1])
    try:
        assert mctx.n_devices() == len(devs)
        data, inputs, _ = ab.synthetic_arith_circuit(3000, mode="local", coeffs="dense", seed_id=8)
        batch = 8 * 37 + 5            # not a whole number of tiles per device
        inp = ab.synthetic_inputs(batch, seed_id=8)
        mc = acvm_b200.CompiledCircuit(mctx, data, inputs)
        assert mctx.broadcast_backend() in (("nccl", "peer-copy") if len(devs) > 1 else ("none",))
        sc = acvm_b200.CompiledCircuit(sctx, data, inputs)
        tail = list(range(mc.num_witnesses - 40, mc.num_witnesses))
        m_out, m_st = mc.solve_batch(inp, batch, out_ids=tail)
        s_out, s_st = sc.solve_batch(inp, batch, out_ids=tail)
        assert m_out == s_out and [(s.status, s.error, s.opcode_index) for s in m_st] == [(s.status, s.error, s.opcode_index) for s in s_st]
        ri = mc.run_info()
        assert ri["kernel_launches"] >= 3 * min(len(devs), -(-batch // 8))
        # a few instances against the oracle, first and last shard included
        oc = acir.decode_circuit(data)
        rows = witness_rows(m_out, batch, len(tail))
        for i in (0, batch // 2, batch - 1):
            iw = inputs_to_dicts(inp, batch, inputs)[i]
            ost, owm, _ = pwg.solve_circuit(oc, iw)
            assert ost == "Solved" and rows[i] == [owm[w] for w in tail]
    finally:
        mctx.close()
        sctx.close()


def test_multi_device_context_reports_per_instance_failures_in_place():
    devs = _devices()
    ctx = acvm_b200.Context(devs)
    try:
        b = ab.CircuitBuilder()
        b.arithmetic([(1, 1, 2)], [(ab.P - 1, 3)], 0)       # w3 = w1 * w2
        b.arithmetic([], [(1, 1)], ab.P - 2)                # check w1 == 2
        b.logic("AND", (3, 64), (2, 64), 4)
        data = b.to_bytes()
        batch = 301
        vals = [(2 if i % 7 else 9, i + 1) for i in range(batch)]
        inp = b"".join(a.to_bytes(32, "big") + c.to_bytes(32, "big") for a, c in vals)
        circ = acvm_b200.CompiledCircuit(ctx, data, [1, 2])
        out, st = circ.solve_batch(inp, batch)
        nw = circ.num_witnesses
        for i, (a, c) in enumerate(vals):
            if a == 2:
                assert st[i].status == "Solved"
                assert int.from_bytes(out[(i * nw + 4) * 32:(i * nw + 5) * 32], "big") == ((a * c) & c)
            else:
                assert (st[i].status, st[i].error, st[i].opcode_index) == ("Failure", "UnsatisfiedConstrain", 1)
    finally:
        ctx.close()


def test_multi_device_context_runs_host_segments_on_every_shard():
    """A circuit with host segments (Brillig on the host VM between device segments) and directives: every shard of a
    multi-device context carries its own copy of the circuit for the host VM; results must equal the single-device ones."""
    import test_host_logic as thl
    devs = _devices()
    mctx = acvm_b200.Context(devs)
    sctx = acvm_b200.Context(devs[0])
    try:
        data = thl._brillig_circuit()
        rows, one = thl._brillig_inputs()
        reps = 40                                   # 240 instances: several tiles per device
        inp = one * reps
        batch = len(rows) * reps
        mc = acvm_b200.CompiledCircuit(mctx, data, [1, 2, 3])
        sc = acvm_b200.CompiledCircuit(sctx, data, [1, 2, 3])
        assert mc.info["n_host_segments"] >= 1
        m_out, m_st, m_pr = mc.solve_batch(inp, batch, want_present=True)
        s_out, s_st, s_pr = sc.solve_batch(inp, batch, want_present=True)
        assert [(s.status, s.error, s.opcode_index, s.aux) for s in m_st] == [(s.status, s.error, s.opcode_index, s.aux) for s in s_st]
        assert m_pr == s_pr and m_out == s_out
        assert len({s.status for s in m_st}) > 1     # the input rows mix solved and failing instances
    finally:
        mctx.close()
        sctx.close()
