"""GPU parity: arithmetic-gate path through the C ABI vs the oracle (bit-exact canonical values)."""
import pytest

import acvm_b200
from acvm_b200 import acir_builder as ab
from conftest import inputs_to_dicts, witness_rows
from oracle import acir, pwg

pytestmark = pytest.mark.gpu


def _check(ctx, data, input_witnesses, batch, inp, expect_all_solved=True):
    circ = acvm_b200.CompiledCircuit(ctx, data, input_witnesses)
    out, st, pres = circ.solve_batch(inp, batch, want_present=True)
    rows = witness_rows(out, batch, circ.num_witnesses)
    oc = acir.decode_circuit(data)
    assign = circ.assign_opcodes()
    nw = circ.num_witnesses
    for i, iw in enumerate(inputs_to_dicts(inp, batch, input_witnesses)):
        if batch > 64 and i >= 32 and i % 8 and i < batch - 8:   # large batches: the pure-Python oracle checks a sample
            continue
        ost, owm, oerr = pwg.solve_circuit(oc, iw)
        assert st[i].status == ost, (i, st[i], oerr)
        if ost == "Failure":
            assert st[i].error == oerr.kind, (i, st[i], oerr)
            if oerr.opcode_location is not None:
                assert st[i].opcode_index == oerr.opcode_location
        limit = 0xFFFFFFFF if ost == "Solved" else st[i].opcode_index
        got = {w: rows[i][w] for w in range(nw) if pres[i * nw + w]}
        assert got == owm, f"instance {i}: witness map differs"
        if 0xFFFFFFFD not in assign:  # static plans: presence is derivable from the plan alone
            assert got.keys() == {w for w in range(nw) if assign[w] == 0xFFFFFFFE or (assign[w] != 0xFFFFFFFF and assign[w] < limit)}
    circ.close()
    return st


@pytest.mark.parametrize("batch", [1, 33, 257])
@pytest.mark.parametrize("mode,coeffs", [("local", "dense"), ("global", "noir-like")])
def test_synthetic_1k(ctx, batch, mode, coeffs):
    data, inputs, _ = ab.synthetic_arith_circuit(1024, mode=mode, coeffs=coeffs)
    _check(ctx, data, inputs, batch, ab.synthetic_inputs(batch))


@pytest.mark.parametrize("scaled", [0, 1])
@pytest.mark.parametrize("mode,coeffs", [("local", "dense"), ("local", "noir-like")])
def test_scaled_and_canonical_columns_agree(scaled, mode, coeffs):
    """Columns written and read only by gates hold lambda_w * value (one Montgomery reduction per multiplicative gate); with
    the option off every column is canonical (two reductions).  Both must return the oracle's canonical witness map, and the
    device checksum (which un-scales on the device) must be the same number."""
    c = acvm_b200.Context(0)
    try:
        c.set_option("scaled_columns", scaled)
        data, inputs, _ = ab.synthetic_arith_circuit(2000, mode=mode, coeffs=coeffs, seed_id=5)
        circ = acvm_b200.CompiledCircuit(c, data, inputs)
        assert circ.info["scaled_columns"] == scaled
        assert (circ.info["n_gate_one_reduction"] > 0) == bool(scaled)
        circ.close()
        batch = 19
        inp = ab.synthetic_inputs(batch, seed_id=5)
        _check(c, data, inputs, batch, inp)
        circ = acvm_b200.CompiledCircuit(c, data, inputs)
        b = acvm_b200.DeviceBatch(circ, batch)
        b.stage_inputs(0, inp)
        b.run_staged(0)
        sums = b.checksums()
        b.close()
        circ.close()
    finally:
        c.close()
    assert _CHECKSUMS.setdefault((mode, coeffs), sums) == sums


_CHECKSUMS = {}


def test_ring_of_recent_values_kernel_variant():
    """Opt-in shared-memory ring (context option ring_bytes): operands produced by recent gate / logic micro-ops are read from
    shared memory.  Same results as the default kernel, bit-exact vs the oracle."""
    c = acvm_b200.Context(0)
    try:
        c.set_option("ring_bytes", 24576)
        data, inputs, _ = ab.synthetic_arith_circuit(3000, mode="local", coeffs="dense", seed_id=6)
        circ = acvm_b200.CompiledCircuit(c, data, inputs)
        assert circ.info["ring_slots"] == 96 and circ.info["n_ring_reads"] > 0.8 * circ.info["n_operand_reads"]
        circ.close()
        _check(c, data, inputs, 41, ab.synthetic_inputs(41, seed_id=6))
        b = ab.CircuitBuilder()
        b.logic("AND", (1, 64), (2, 64), 10)
        b.logic("XOR", (10, 64), (1, 64), 11)
        b.range((11, 64))
        b.arithmetic([(3, 10, 11)], [(1, 1), (ab.P - 1, 12)], 7)
        b.arithmetic([], [(1, 12), (1, 11), (ab.P - 1, 13)], 0)
        _check(c, b.to_bytes(), [1, 2], 9, ab.synthetic_inputs(9, n_inputs=2, seed_id=6))
    finally:
        c.close()


def test_addition_golden(ctx, golden):
    fx = golden["acvm_js_shared"]["addition"]
    data = bytes(fx["bytecode"])
    iw = {int(k): int(v, 16) for k, v in fx["initialWitnessMap"].items()}
    vm = acvm_b200.ACVM(ctx, data, iw)
    st = vm.solve()
    assert st.status == "Solved"
    wm = vm.finalize()
    assert wm[fx["resultWitness"]] == int(fx["expectedResult"], 16)
    assert wm == {1: 1, 2: 2, 3: 3}


def test_unsatisfied_constraint_location(ctx):
    # acvm/tests/solver.rs:490-525 : a == b with a != b  -> UnsatisfiedConstrain at Acir(0)
    b = ab.CircuitBuilder()
    b.arithmetic([], [(1, 1), (ab.P - 1, 2)], 0)
    data = b.to_bytes()
    inp = (5).to_bytes(32, "big") + (6).to_bytes(32, "big") + (7).to_bytes(32, "big") + (7).to_bytes(32, "big")
    circ = acvm_b200.CompiledCircuit(ctx, data, [1, 2])
    _, st = circ.solve_batch(inp, 2, want_witness=False)
    assert (st[0].status, st[0].error, st[0].opcode_index) == ("Failure", "UnsatisfiedConstrain", 0)
    assert st[1].status == "Solved"


def test_mixed_failures_lowest_opcode_wins(ctx):
    # instance-dependent failures at different opcodes; reordering by the scheduler must not change which one is reported
    b = ab.CircuitBuilder()
    b.arithmetic([(1, 1, 2)], [(ab.P - 1, 3)], 0)        # w3 = w1*w2
    b.arithmetic([], [(1, 3), (ab.P - 1, 4)], 5)         # w4 = w3 + 5
    b.arithmetic([], [(1, 1)], ab.P - 2)                 # check w1 == 2
    b.arithmetic([(1, 4, 4)], [(ab.P - 1, 5)], 0)        # w5 = w4^2
    b.arithmetic([], [(1, 2)], ab.P - 3)                 # check w2 == 3
    data = b.to_bytes()
    cases = [(2, 3), (9, 3), (2, 9), (9, 9)]
    inp = b"".join(a.to_bytes(32, "big") + c.to_bytes(32, "big") for a, c in cases)
    _check(ctx, data, [1, 2], len(cases), inp)


def test_wide_expression_chain(ctx):
    # pre-CSat style gate with many terms: exercises chained micro-gates through temporaries
    rng = ab.SplitMix64(7)
    b = ab.CircuitBuilder()
    lin = [(rng.nonzero_field(), w) for w in range(1, 10)]
    mul = [(rng.nonzero_field(), 1 + rng.below(9), 1 + rng.below(9)) for _ in range(3)]
    b.arithmetic(mul[:1], lin + [(rng.nonzero_field(), 10)], rng.field())
    b.arithmetic([], [(3, 10), (5, 11)], 1)
    b.arithmetic(mul[1:2], [(1, 10), (1, 11), (1, 12)], 0)
    data = b.to_bytes()
    batch = 5
    inp = ab.synthetic_inputs(batch, n_inputs=9, seed_id=3)
    _check(ctx, data, list(range(1, 10)), batch, inp)


def test_too_many_unknowns_static(ctx):
    b = ab.CircuitBuilder()
    b.arithmetic([], [(1, 1), (1, 2), (1, 3)], 0)  # w2, w3 unknown
    data = b.to_bytes()
    circ = acvm_b200.CompiledCircuit(ctx, data, [1])
    _, st = circ.solve_batch((4).to_bytes(32, "big"), 1, want_witness=False)
    assert (st[0].status, st[0].error, st[0].opcode_index) == ("Failure", "OpcodeNotSolvable.ExpressionHasTooManyUnknowns", 0)


def test_logic_and_range(ctx):
    b = ab.CircuitBuilder()
    b.logic("AND", (1, 32), (2, 32), 3)
    b.logic("XOR", (1, 32), (2, 32), 4)
    b.logic("XOR", (1, 254), (2, 254), 5)
    b.logic("AND", (1, 13), (2, 13), 6)
    b.range((3, 32))
    b.range((1, 64))
    data = b.to_bytes()
    vals = [(0xDEADBEEF12345678, 0x0F0F0F0FF0F0F0F0), (ab.P - 1, ab.P - 2), (1 << 64, 3), (0, 0), ((1 << 64) - 1, 1 << 253)]
    inp = b"".join(a.to_bytes(32, "big") + c.to_bytes(32, "big") for a, c in vals)
    _check(ctx, data, [1, 2], len(vals), inp)


def test_no_gpu_fallback_symbols():
    # the library must be the thing that ran: a run record with kernel launches exists
    c = acvm_b200.Context(0)
    data, inputs, _ = ab.synthetic_arith_circuit(64)
    circ = acvm_b200.CompiledCircuit(c, data, inputs)
    circ.solve_batch(ab.synthetic_inputs(4), 4)
    ri = circ.run_info()
    assert ri["kernel_launches"] >= 3 and ri["kernel_ms"] > 0


def test_value_dependent_gates(ctx):
    # the unknown is a multiplication operand: per-instance known-set divergence (SURVEY 8a quirks 2+3)
    from test_host_logic import _vd_circuit, _vd_inputs
    rows, inp = _vd_inputs()
    st = _check(ctx, _vd_circuit(), [1, 2, 3, 4], len(rows), inp)
    assert {s.status for s in st} == {"Solved", "Failure"}


def test_inverse_gate_pattern(ctx):
    # x * x_inv - 1 = 0 solved for x_inv (per-lane field inversion), then used downstream
    b = ab.CircuitBuilder()
    b.arithmetic([(1, 1, 2)], [], ab.P - 1)
    b.arithmetic([(1, 2, 2)], [(ab.P - 1, 3)], 0)
    data = b.to_bytes()
    vals = [5, 1, ab.P - 1, 123456789, 0]
    inp = b"".join(v.to_bytes(32, "big") for v in vals)
    st = _check(ctx, data, [1], len(vals), inp)
    assert [s.status for s in st] == ["Solved"] * 4 + ["Failure"]
    vm = acvm_b200.ACVM(ctx, data, {1: 5})
    assert vm.solve().status == "Solved"
    assert vm.finalize()[2] == int("135b52945a13d9aa49b9b57c33cd568ba9ae5ce9ca4a2d06e7f3fbd4c6666667", 16)  # 1/5, foreign_call.ts:20-27


def test_directives_and_memory(ctx, golden):
    from test_host_logic import _dir_mem_circuit, _dir_mem_inputs
    rows, inp = _dir_mem_inputs()
    st = _check(ctx, _dir_mem_circuit(), [1, 2, 3], len(rows), inp)
    assert "Solved" in {s.status for s in st}
    fx = golden["acvm_js_shared"]["memory_op"]   # acvm_js/test/shared/memory_op.ts
    vm = acvm_b200.ACVM(ctx, bytes(fx["bytecode"]), {int(k): int(v, 16) for k, v in fx["initialWitnessMap"].items()})
    assert vm.solve().status == "Solved"
    assert vm.finalize() == {int(k): int(v, 16) for k, v in fx["expectedWitnessMap"].items()}


def test_brillig_on_host_between_device_segments(ctx, golden):
    # north star: "Brillig opcodes execute on the host brillig_vm with results DMA'd back into the device WitnessMap"
    from test_host_logic import _brillig_circuit, _brillig_inputs
    rows, inp = _brillig_inputs()
    st = _check(ctx, _brillig_circuit(), [1, 2, 3], len(rows), inp)
    assert {"Solved", "Failure"} <= {s.status for s in st}
    # a larger batch through the same plan (threads on the host side, tiles on the device side)
    big = inp * 40
    _check(ctx, _brillig_circuit(), [1, 2, 3], len(rows) * 40, big)
    # the reference's foreign-call fixtures stop at RequiresForeignCall with ip on the Brillig opcode (mod.rs:267)
    fx = golden["acvm_js_shared"]["foreign_call"]
    vm = acvm_b200.ACVM(ctx, bytes(fx["bytecode"]), {1: 5})
    s = vm.solve()
    assert (s.status, s.opcode_index) == ("RequiresForeignCall", 0)
    assert vm.witness_map() == {1: 5}


def test_foreign_call_resolution_golden(ctx, golden):
    # acvm_js/test/shared/{foreign_call,complex_foreign_call}.ts: stall, hand back inputs, resolve, finish
    sh = golden["acvm_js_shared"]
    fx = sh["foreign_call"]
    vm = acvm_b200.ACVM(ctx, bytes(fx["bytecode"]), {1: 5})
    assert vm.solve().status == "RequiresForeignCall" and vm.instruction_pointer() == 0
    assert vm.get_pending_foreign_call() == ("invert", [[5]])
    vm.resolve_pending_foreign_call([int(golden["kats"]["inv5"], 16)])
    assert vm.solve().status == "Solved"
    assert vm.finalize() == {int(k): int(v, 16) for k, v in fx["expectedWitnessMap"].items()}
    fx = sh["complex_foreign_call"]
    vm = acvm_b200.ACVM(ctx, bytes(fx["bytecode"]), {1: 1, 2: 2, 3: 3})
    assert vm.solve().status == "RequiresForeignCall"
    assert vm.get_pending_foreign_call() == ("complex", [[1, 2, 3], [6]])
    vm.resolve_pending_foreign_call([[2, 6, 12], 6, 12])
    assert vm.solve().status == "Solved"
    assert vm.finalize() == {int(k): int(v, 16) for k, v in fx["expectedWitnessMap"].items()}


def test_device_checksum_matches_host_definition(ctx):
    data, inputs, _ = ab.synthetic_arith_circuit(300)
    circ = acvm_b200.CompiledCircuit(ctx, data, inputs)
    batch = 19
    inp = ab.synthetic_inputs(batch, seed_id=5)
    b = acvm_b200.DeviceBatch(circ, batch)
    b.upload(inp)
    b.run()
    sums = b.checksums()
    out = b.download()
    nw = circ.num_witnesses
    assign = circ.assign_opcodes()
    for i in range(batch):
        wm = {w: int.from_bytes(out[(i * nw + w) * 32:(i * nw + w + 1) * 32], "big") for w in range(nw) if assign[w] != 0xFFFFFFFF}
        assert acvm_b200.witness_checksum(wm) == sums[i]


def test_forced_subbatching_and_ragged_tail(ctx):
    # the sub-batch loop of acvmb_solve_batch (used when the witness columns do not fit in HBM) with a ragged last sub-batch
    data, inputs, _ = ab.synthetic_arith_circuit(256, seed_id=12)
    # budget = fixed overhead (64 MiB + four staging buffers) + 40 instances of columns: two half-size column buffers, the VM
    # kernel of sub-batch k+1 overlapping the gather + D2H of sub-batch k (runtime.cu acvmb_solve_batch_ex)
    c2 = acvm_b200.Context(0)
    circ = acvm_b200.CompiledCircuit(c2, data, inputs)
    c2.set_option("max_resident_bytes", (64 << 20) + 4 * 103 * circ.num_witnesses * 32 + 40 * (circ.info["n_slots"] * 32 + 8 + len(inputs) * 32))
    batch = 103
    inp = ab.synthetic_inputs(batch, seed_id=12)
    out, st = circ.solve_batch(inp, batch)
    info = circ.run_info()
    assert info["n_subbatches"] >= 3 and 16 <= info["resident_instances"] < batch // 2, info
    ref = acvm_b200.CompiledCircuit(ctx, data, inputs)
    out_ref, st_ref = ref.solve_batch(inp, batch)
    assert out == out_ref and [s.status for s in st] == [s.status for s in st_ref] == ["Solved"] * batch
    c2.close()


def test_plan_blob_roundtrip_with_host_segments(ctx):
    # multi-GPU distribution path: rank 0 serialises the compiled plan (which embeds the ACIR bytes when Brillig opcodes need
    # the host VM), other ranks deserialise and must produce identical results
    from test_host_logic import _brillig_circuit, _brillig_inputs
    rows, inp = _brillig_inputs()
    data = _brillig_circuit()
    a = acvm_b200.CompiledCircuit(ctx, data, [1, 2, 3])
    blob = a.serialize()
    b = acvm_b200.CompiledCircuit.from_blob(ctx, blob, [1, 2, 3])
    oa, sa, pa = a.solve_batch(inp, len(rows), want_present=True)
    ob, sb, pb = b.solve_batch(inp, len(rows), want_present=True)
    assert (oa, pa) == (ob, pb) and [(s.status, s.error, s.opcode_index) for s in sa] == [(s.status, s.error, s.opcode_index) for s in sb]


@pytest.mark.parametrize("n,tup,sort_by", [(5, 1, [0]), (8, 2, [1, 0]), (33, 2, [0]), (2, 1, [0]), (1, 1, [0])])
def test_permutation_sort_directive(ctx, n, tup, sort_by):
    """Directive::PermutationSort runs as a host segment between two device segments (directives/mod.rs:88-121)."""
    from sort_cases import sort_circuit, sort_rows
    from test_gpu_blackbox import _check_circuit
    data, inputs = sort_circuit(n, tup, sort_by)
    batch = 37
    st = _check_circuit(ctx, data, inputs, batch, sort_rows(n, tup, batch))
    assert all(s.status == "Solved" for s in st)
    data, _ = sort_circuit(n, tup, sort_by, preassign_bit=True)
    _check_circuit(ctx, data, inputs, batch, sort_rows(n, tup, batch))


def test_solve_batch_reuses_cached_columns(ctx):
    # repeated calls on one circuit keep the column buffers (runtime.cu cached_batch); another circuit or size drops them
    data, inputs, _ = ab.synthetic_arith_circuit(200, seed_id=3)
    data2, inputs2, _ = ab.synthetic_arith_circuit(120, seed_id=4)
    a = acvm_b200.CompiledCircuit(ctx, data, inputs)
    b = acvm_b200.CompiledCircuit(ctx, data2, inputs2)
    oc = acir.decode_circuit(data)
    for seed, batch in ((1, 9), (2, 9), (3, 17), (4, 17)):
        inp = ab.synthetic_inputs(batch, seed_id=seed)
        out, st = a.solve_batch(inp, batch)
        if seed == 2:
            b.solve_batch(ab.synthetic_inputs(5, seed_id=9), 5)   # evicts a's buffers
        rows = witness_rows(out, batch, a.num_witnesses)
        for i, iw in enumerate(inputs_to_dicts(inp, batch, inputs)):
            ost, owm, _ = pwg.solve_circuit(oc, iw)
            assert st[i].status == ost == "Solved" and all(rows[i][w] == v for w, v in owm.items())
    a.close()       # owner of the cached buffers goes away first
    out, st = b.solve_batch(ab.synthetic_inputs(5, seed_id=9), 5)
    assert all(s.status == "Solved" for s in st)
    b.close()
