"""The reference's own integration scenarios (acvm/tests/solver.rs) re-stated as circuits in the reference wire format and
run through BOTH the oracle (CPU, always) and the C ABI on the GPU (-m gpu), so the two sets of tests read alike."""
import pytest

import acvm_b200
from acvm_b200 import acir_builder as ab
from oracle import acir, field as F, pwg

R = lambda r: ("Register", r)
M1 = ab.P - 1


def _oracle_dependent_execution():  # solver.rs:308-426
    b = ab.CircuitBuilder()
    b.arithmetic([], [(M1, 1), (1, 2)], 0)                                     # x == y
    b.brillig([("Single", ab.wexpr(1)), ("Single", ab.cexpr(0)), ("Single", ab.wexpr(2))],
              [("Simple", 1), ("Simple", 3), ("Simple", 2), ("Simple", 4)], [
        dict(op="ForeignCall", function="invert", destinations=[R(1)], inputs=[R(0)]),
        dict(op="ForeignCall", function="invert", destinations=[R(3)], inputs=[R(2)]),
    ])
    b.arithmetic([], [(M1, 3), (1, 4)], 0)                                     # x_inv == y_inv
    return b.to_bytes(current_witness_index=4)


def _brillig_oracle_predicate():  # solver.rs:428-487: predicate is the constant 0 -> outputs zeroed, no foreign call
    b = ab.CircuitBuilder()
    b.brillig([("Single", ([], [(1, 1), (1, 2)], 0)), ("Single", ab.cexpr(0))],
              [("Simple", 4), ("Simple", 3), ("Simple", 5), ("Simple", 6)], [
        dict(op="BinaryFieldOp", destination=2, bop=4, lhs=0, rhs=1),
        dict(op="ForeignCall", function="invert", destinations=[R(1)], inputs=[R(0)]),
    ], predicate=ab.cexpr(0))
    return b.to_bytes()


def _unsatisfied_opcode_resolved():  # solver.rs:489-524
    b = ab.CircuitBuilder()
    b.arithmetic([], [(1, 0), (M1, 1), (M1, 2), (M1, 3)], 0)
    return b.to_bytes(current_witness_index=3)


def _unsatisfied_opcode_resolved_brillig():  # solver.rs:526-608: trap at brillig index 2 of acir opcode 0
    b = ab.CircuitBuilder()
    b.brillig([("Single", ab.wexpr(4)), ("Single", ab.wexpr(5))], [("Simple", 6)], [
        dict(op="BinaryFieldOp", destination=2, bop=4, lhs=0, rhs=1),
        dict(op="JumpIf", condition=2, location=3),
        dict(op="Trap"),
        dict(op="Stop"),
    ], predicate=ab.cexpr(1))
    b.arithmetic([], [(1, 0), (M1, 1), (M1, 2), (M1, 3)], 0)
    return b.to_bytes()


def _memory_operations():  # solver.rs:610-648
    b = ab.CircuitBuilder()
    b.memory_init(0, [1, 2, 3, 4, 5])
    b.memory_op(0, ab.cexpr(0), ab.wexpr(6), ab.wexpr(7))
    b.arithmetic([], [(1, 7), (M1, 8)], 1)
    return b.to_bytes()


# ---- oracle (CPU) ------------------------------------------------------------------------------
def test_oracle_reference_scenarios():
    c = acir.decode_circuit(_oracle_dependent_execution())
    vm = pwg.ACVM(pwg.StubbedBackend(), c.opcodes, {1: 2, 2: 2})
    assert vm.solve() == "RequiresForeignCall" and vm.instruction_pointer == 1
    vm.resolve_pending_foreign_call([("Single", F.inverse(2))])
    assert vm.solve() == "RequiresForeignCall" and vm.instruction_pointer == 1
    vm.resolve_pending_foreign_call([("Single", F.inverse(2))])
    assert vm.solve() == "Solved"

    c = acir.decode_circuit(_brillig_oracle_predicate())
    st, wm, _ = pwg.solve_circuit(c, {1: 2, 2: 3}, pwg.StubbedBackend())
    assert st == "Solved" and [wm[w] for w in (3, 4, 5, 6)] == [0, 0, 0, 0]

    c = acir.decode_circuit(_unsatisfied_opcode_resolved())
    st, _, err = pwg.solve_circuit(c, {0: 4, 1: 2, 2: 1, 3: 2}, pwg.StubbedBackend())
    assert (st, err.kind, err.opcode_location) == ("Failure", "UnsatisfiedConstrain", 0)

    c = acir.decode_circuit(_unsatisfied_opcode_resolved_brillig())
    st, _, err = pwg.solve_circuit(c, {0: 4, 1: 2, 2: 1, 3: 2, 4: 0, 5: 1, 6: 0}, pwg.StubbedBackend())
    assert (st, err.kind, err.message, err.index) == ("Failure", "BrilligFunctionFailed", "explicit trap hit in brillig", 2)

    c = acir.decode_circuit(_memory_operations())
    st, wm, _ = pwg.solve_circuit(c, {1: 1, 2: 2, 3: 3, 4: 4, 5: 5, 6: 4}, pwg.StubbedBackend())
    assert st == "Solved" and wm[8] == 6


# ---- the same scenarios through the C ABI on the GPU ------------------------------------------
@pytest.mark.gpu
def test_gpu_oracle_dependent_execution(ctx):
    vm = acvm_b200.ACVM(ctx, _oracle_dependent_execution(), {1: 2, 2: 2})
    s = vm.solve()
    assert s.status == "RequiresForeignCall" and vm.instruction_pointer() == 1, "should stall on brillig"
    fn, inputs = vm.get_pending_foreign_call()
    assert fn == "invert" and len(inputs) == 1, "Should be waiting for a single input"
    vm.resolve_pending_foreign_call([F.inverse(inputs[0][0])])
    s = vm.solve()
    assert s.status == "RequiresForeignCall" and vm.instruction_pointer() == 1
    fn, inputs = vm.get_pending_foreign_call()
    vm.resolve_pending_foreign_call([F.inverse(inputs[0][0])])
    assert vm.solve().status == "Solved", "should be fully solved"
    wm = vm.finalize()
    assert wm[3] == wm[4] == F.inverse(2)


@pytest.mark.gpu
def test_gpu_brillig_oracle_predicate(ctx):
    vm = acvm_b200.ACVM(ctx, _brillig_oracle_predicate(), {1: 2, 2: 3})
    assert vm.solve().status == "Solved"
    wm = vm.finalize()
    assert [wm[w] for w in (3, 4, 5, 6)] == [0, 0, 0, 0]


@pytest.mark.gpu
def test_gpu_unsatisfied_opcode_resolved(ctx):
    vm = acvm_b200.ACVM(ctx, _unsatisfied_opcode_resolved(), {0: 4, 1: 2, 2: 1, 3: 2})
    s = vm.solve()
    assert (s.status, s.error, s.opcode_index) == ("Failure", "UnsatisfiedConstrain", 0)
    with pytest.raises(acvm_b200.AcvmError) as e:   # finalize() panics unless Solved (mod.rs:176-181)
        vm.finalize()
    assert e.value.rc == -7


@pytest.mark.gpu
def test_gpu_unsatisfied_opcode_resolved_brillig(ctx):
    vm = acvm_b200.ACVM(ctx, _unsatisfied_opcode_resolved_brillig(), {0: 4, 1: 2, 2: 1, 3: 2, 4: 0, 5: 1, 6: 0})
    s = vm.solve()
    # BrilligFunctionFailed { call_stack: [Brillig { acir_index: 0, brillig_index: 2 }] }
    assert (s.status, s.error, s.opcode_index, s.aux) == ("Failure", "BrilligFunctionFailed", 0, 2)


@pytest.mark.gpu
def test_gpu_memory_operations(ctx):
    vm = acvm_b200.ACVM(ctx, _memory_operations(), {1: 1, 2: 2, 3: 3, 4: 4, 5: 5, 6: 4})
    assert vm.solve().status == "Solved"
    assert vm.finalize()[8] == 6


def _lockstep(ctx, data, initial, probe):
    """acvmb_vm_solve_opcode against the oracle's ACVM::solve_opcode, call by call: status, instruction pointer and the
    witnesses visible through witness_map() must agree after every step."""
    from oracle import acir as oacir
    vm = acvm_b200.ACVM(ctx, data, initial)
    ovm = pwg.ACVM(pwg.OracleBackend(), oacir.decode_circuit(data).opcodes, dict(initial))
    steps = 0
    while ovm.status == "InProgress":
        ost = ovm.solve_opcode()
        st = vm.solve_opcode()
        steps += 1
        assert st.status == ost, (steps, st, ost)
        assert vm.instruction_pointer() == ovm.instruction_pointer
        wm = vm.witness_map()
        for w in probe:
            assert wm.get(w) == ovm.witness_map.get(w), (steps, w)
        if ost == "Failure":
            assert st.error == ovm.error.kind
    return vm, ovm, steps


@pytest.mark.gpu
def test_gpu_solve_opcode_steps_like_the_reference(ctx):
    # five arithmetic opcodes + a logic op; the check at opcode 4 fails unless w1 == 2
    b = ab.CircuitBuilder()
    b.arithmetic([(1, 1, 2)], [(ab.P - 1, 3)], 0)        # w3 = w1*w2
    b.arithmetic([], [(1, 3), (ab.P - 1, 4)], 5)         # w4 = w3 + 5
    b.logic("XOR", (1, 64), (2, 64), 5)
    b.arithmetic([(1, 4, 5)], [(ab.P - 1, 6)], 0)        # w6 = w4*w5
    b.arithmetic([], [(1, 1)], ab.P - 2)                 # check w1 == 2
    b.arithmetic([(1, 6, 6)], [(ab.P - 1, 7)], 0)        # w7 = w6^2
    data = b.to_bytes()
    vm, ovm, steps = _lockstep(ctx, data, {1: 2, 2: 9}, range(1, 8))
    assert steps == 6 and vm.get_status().status == "Solved" and vm.finalize() == ovm.finalize()
    with pytest.raises(acvm_b200.AcvmError) as e:        # one more step indexes past the opcodes: the reference panics
        vm.solve_opcode()
    assert e.value.rc == -7
    vm, ovm, steps = _lockstep(ctx, data, {1: 3, 2: 9}, range(1, 8))
    assert steps == 5 and vm.instruction_pointer() == 4
    st = vm.solve_opcode()                               # failure is terminal: the same opcode fails again
    assert (st.status, st.error, st.opcode_index) == ("Failure", "UnsatisfiedConstrain", 4)
    # solve() after a few single steps finishes the job
    vm = acvm_b200.ACVM(ctx, data, {1: 2, 2: 9})
    assert vm.solve_opcode().status == "InProgress" and vm.instruction_pointer() == 1
    assert 4 not in vm.witness_map() and vm.witness_map()[3] == 18
    assert vm.solve().status == "Solved" and vm.instruction_pointer() == 6


@pytest.mark.gpu
def test_gpu_solve_opcode_stops_at_a_foreign_call(ctx):
    data = _oracle_dependent_execution()
    vm = acvm_b200.ACVM(ctx, data, {1: 2, 2: 2})
    assert vm.solve_opcode().status == "InProgress" and vm.instruction_pointer() == 1
    st = vm.solve_opcode()
    assert st.status == "RequiresForeignCall" and vm.instruction_pointer() == 1
    assert vm.solve_opcode().status == "RequiresForeignCall"      # unresolved: the Brillig opcode waits again
    fn, inputs = vm.get_pending_foreign_call()
    vm.resolve_pending_foreign_call([F.inverse(inputs[0][0])])
    assert vm.instruction_pointer() == 1
    st = vm.solve_opcode()
    assert st.status == "RequiresForeignCall" and vm.instruction_pointer() == 1   # second foreign call of the same opcode
    fn, inputs = vm.get_pending_foreign_call()
    vm.resolve_pending_foreign_call([F.inverse(inputs[0][0])])
    assert vm.solve().status == "Solved"
