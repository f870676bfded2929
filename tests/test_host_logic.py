"""CPU tests of the product's host side: library loads and exports the header's symbols, the C++ decoder
and plan compiler agree with the oracle (through the test-only plan interpreter), field limb algorithms."""
import ctypes
import hashlib
import os
import random
import re
import subprocess

import pytest

import acvm_b200
from acvm_b200 import acir_builder as ab
from conftest import ROOT, inputs_to_dicts
from oracle import acir, pwg
import plan_interp


def test_library_exports_every_declared_symbol():
    lib = acvm_b200.lib()
    hdr = open(os.path.join(ROOT, "include", "acvm_b200.h")).read()
    declared = set(re.findall(r"\b(acvmb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/acvm_b200.h but not exported"
    assert declared <= set(lib._acvmb_signatures) | {"acvmb_last_error"}


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(acvm_b200.AcvmError) as e:
        acvm_b200.Context(0)
    assert e.value.rc == -2 and "no CPU fallback" in str(e.value)


def _interp_vs_oracle(data, inputs, inp, batch, S=16, **plan_options):
    info, blob = acvm_b200.compile_plan_host(data, inputs, S, **plan_options)
    plan = plan_interp.PlanBlob(blob)
    oc = acir.decode_circuit(data)
    n = len(inputs)
    for i in range(batch):
        iw = {w: int.from_bytes(inp[(i * n + k) * 32:(i * n + k + 1) * 32], "big") for k, w in enumerate(inputs)}
        st, wm = plan_interp.run_plan(plan, iw, circuit=oc)
        ost, owm, oerr = pwg.solve_circuit(oc, iw)
        assert st[0] == ost, (st, ost, oerr)
        if ost == "Failure":
            assert acvm_b200.solver.ERR_NAMES[st[1]] == oerr.kind
            if oerr.opcode_location is not None:
                assert st[2] == oerr.opcode_location
        assert wm == owm
    return info


@pytest.mark.parametrize("S", [1, 4, 16, 32])
@pytest.mark.parametrize("mode,coeffs", [("local", "dense"), ("global", "noir-like")])
def test_plan_matches_oracle_synthetic(S, mode, coeffs):
    data, inputs, _ = ab.synthetic_arith_circuit(600, mode=mode, coeffs=coeffs)
    info = _interp_vs_oracle(data, inputs, ab.synthetic_inputs(2), 2, S)
    assert info["n_micro_ops"] == 600 and info["ref_fr_inv"] == 600 - 600 // 16


def test_plan_batched_inversion_path():
    # circuits above 2048 opcodes compile in two passes with one batched (Montgomery-trick) inversion in between; the record
    # pass skips the column-scale arithmetic (plan.cpp lower_sum), so the two passes must still ask for the same inverses
    data, inputs, _ = ab.synthetic_arith_circuit(3000, mode="local", coeffs="dense")
    info = _interp_vs_oracle(data, inputs, ab.synthetic_inputs(1), 1, 16)
    assert info["n_opcodes"] == 3000
    data, inputs, _ = ab.synthetic_arith_circuit(2500, mode="global", coeffs="noir-like")
    info = _interp_vs_oracle(data, inputs, ab.synthetic_inputs(1), 1, 16)
    assert info["n_opcodes"] == 2500
    # wide expressions (partial sums through temporaries), a product whose operands also appear linearly, and at the end a
    # value-dependent gate: the replay pass meets scaled columns there and the circuit is recompiled with canonical ones
    for value_dependent in (False, True):
        rng = ab.SplitMix64(77)
        b = ab.CircuitBuilder()
        nxt = 9
        for i in range(2100):
            lo = max(1, nxt - 40)
            ws = [lo + rng.below(nxt - lo) for _ in range(6)]
            lin = [(rng.nonzero_field(), w) for w in dict.fromkeys(ws)]
            mul = [(rng.nonzero_field(), ws[0], ws[1])] if i % 3 else []
            b.arithmetic(mul, lin + [(rng.nonzero_field(), nxt)], rng.field())
            nxt += 1
        if value_dependent:
            b.arithmetic([(1, nxt - 1, nxt)], [(3, nxt - 2)], 5)     # w_last * w_new + 3 w + 5 = 0: solvable only per value
        info = _interp_vs_oracle(b.to_bytes(), list(range(1, 9)), ab.synthetic_inputs(2, n_inputs=8, seed_id=5), 2, 16)
        assert info["scaled_columns"] == (0 if value_dependent else 1)


def test_plan_failures_and_chains():
    rng = ab.SplitMix64(11)
    b = ab.CircuitBuilder()
    lin = [(rng.nonzero_field(), w) for w in range(1, 10)]
    b.arithmetic([(rng.nonzero_field(), 2, 3)], lin + [(rng.nonzero_field(), 10)], rng.field())
    b.arithmetic([(rng.nonzero_field(), 10, 10), ], [(3, 10), (5, 11)], 1)
    b.arithmetic([], [(1, 1)], ab.P - 2)  # check w1 == 2  (fails unless input is 2)
    b.logic("XOR", (10, 64), (11, 64), 12)
    b.range((12, 64))
    b.range((11, 8))                      # almost surely fails
    data = b.to_bytes()
    inp = ab.synthetic_inputs(3, n_inputs=9, seed_id=5)
    inp = (2).to_bytes(32, "big") + inp[32:]   # instance 0 passes the w1 == 2 check
    info = _interp_vs_oracle(data, list(range(1, 10)), inp, 3)
    assert info["n_temps"] > 0


def test_static_failures():
    b = ab.CircuitBuilder()
    b.arithmetic([], [(1, 1), (1, 2), (1, 3)], 0)
    info, _ = acvm_b200.compile_plan_host(b.to_bytes(), [1], 16)
    assert (info["static_fail_present"], info["static_fail_kind"], info["static_fail_opcode"]) == (1, 2, 0)
    b = ab.CircuitBuilder()
    b.arithmetic([], [(1, 1), (ab.P - 1, 2)], 0)
    b.logic("AND", (2, 8), (3, 8), 4)     # witness 3 never assigned -> MissingAssignment(3) at opcode 1
    _interp_vs_oracle(b.to_bytes(), [1], (9).to_bytes(32, "big"), 1)
    info, _ = acvm_b200.compile_plan_host(b.to_bytes(), [1], 16)
    assert (info["static_fail_kind"], info["static_fail_opcode"], info["static_fail_aux"]) == (1, 1, 3)


def test_decoder_rejects_garbage_and_accepts_golden(golden):
    with pytest.raises(acvm_b200.AcvmError) as e:
        acvm_b200.compile_plan_host(b"\x1f\x8b\x08\x00garbage", [], 16)
    assert e.value.rc == -4
    good = bytes(golden["rust_serialization"]["addition_circuit"])
    for cut in (len(good) // 2, len(good) - 1, 11):     # a truncated gzip stream is an error, not a hang or a short read
        with pytest.raises(acvm_b200.AcvmError) as e:
            acvm_b200.compile_plan_host(good[:cut], [1, 2], 16)
        assert e.value.rc == -4
    info, _ = acvm_b200.compile_plan_host(bytes(golden["rust_serialization"]["addition_circuit"]), [1, 2], 16)
    assert info["n_opcodes"] == 1 and info["num_witnesses"] == 5 and info["n_gate_assign"] == 1
    # opcodes outside the supported scope decode fine and are refused loudly, never silently skipped
    with pytest.raises(acvm_b200.AcvmError) as e:
        acvm_b200.compile_plan_host(bytes(golden["rust_serialization"]["schnorr_verify_circuit"]), list(range(1, 77)), 16)
    assert e.value.rc == -5


def test_fr_limb_algorithms_on_host(tmp_path):
    so = tmp_path / "fr_host_shim.so"
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", str(so), os.path.join(ROOT, "tests", "host", "fr_host_shim.cpp")], check=True)
    L = ctypes.CDLL(str(so))
    P = ab.P
    Rinv = pow(1 << 256, -1, P)
    rnd = random.Random(1)

    def arr(vals):
        return (ctypes.c_uint32 * (8 * len(vals)))(*[(v >> (32 * i)) & 0xFFFFFFFF for v in vals for i in range(8)])

    def val(a):
        return sum(int(a[i]) << (32 * i) for i in range(8))

    for it in range(3000):
        k = rnd.choice([1, 2, 3, 4])
        A = [rnd.choice([0, 1, P - 1, rnd.randrange(P)]) for _ in range(k)]
        B = [rnd.choice([0, 1, P - 1, rnd.randrange(P)]) for _ in range(k)]
        if k <= 3 and rnd.random() < 0.3:
            A[0] += P // 6  # lazily reduced operand (< 1.19 p)
        r = (ctypes.c_uint32 * 8)()
        L.t_mont_dot(k, arr(A), arr(B), r)
        got = val(r)
        assert got < 2 * P and got % P == sum(x * y for x, y in zip(A, B)) * Rinv % P
    for it in range(2000):  # gate constant folded into the reduction as the initial accumulator (any value < 2^256)
        k = rnd.choice([1, 2, 3])
        A = [rnd.choice([0, 1, P - 1, rnd.randrange(P)]) for _ in range(k)]
        B = [rnd.choice([0, 1, P - 1, rnd.randrange(P)]) for _ in range(k)]
        if k <= 2 and rnd.random() < 0.3:
            A[0] += (3 * P) // 4   # lazily reduced operand (< 1.76 p)
        init = rnd.choice([0, 1, P - 1, (1 << 256) - 1, rnd.randrange(P)])
        r = (ctypes.c_uint32 * 8)()
        L.t_mont_dot_init(k, arr(A), arr(B), arr([init]), r)
        got = val(r)
        assert got < 2 * P and got % P == (init + sum(x * y for x, y in zip(A, B))) * Rinv % P
    for it in range(1500):  # every FMA/ALU pipe split level of the row helpers computes the same product
        x, y = rnd.choice([0, 1, P - 1, rnd.randrange(P)]), rnd.choice([0, 1, P - 1, rnd.randrange(P)])
        for split in range(5):
            r = (ctypes.c_uint32 * 8)()
            L.t_mont_mul_split(split, arr([x]), arr([y]), r)
            assert val(r) == x * y * Rinv % P, (split, hex(x), hex(y))
    # lazy-reduction bounds the gate kernel relies on: (x+alpha)(y+beta) with both factors < 2p, then <u, w1>.<m0, m1>
    for it in range(3000):
        a, b = rnd.choice([2 * P - 2, rnd.randrange(2 * P - 1)]), rnd.choice([2 * P - 2, rnd.randrange(2 * P - 1)])
        r = (ctypes.c_uint32 * 8)()
        L.t_mont_dot(1, arr([a]), arr([b]), r)
        u = val(r)
        assert u % P == a * b * Rinv % P and u < 1.76 * P + 1
        w, m0, m1 = (rnd.choice([P - 1, rnd.randrange(P)]) for _ in range(3))
        L.t_mont_dot(2, arr([u, w]), arr([m0, m1]), r)
        assert val(r) % P == (u * m0 + w * m1) * Rinv % P and val(r) < 2 * P
    # carry-free 9 x 29-bit representation (fr29.cuh): conversions round-trip and the K-term dot product is exact
    Rinv261 = pow(1 << 261, -1, P)
    for it in range(3000):
        z = rnd.randrange(1 << 256)
        r = (ctypes.c_uint32 * 8)()
        L.t_roundtrip9(arr([z]), r)
        assert val(r) == z
        k = rnd.choice([1, 2, 3, 4])
        A = [rnd.choice([0, 1, P - 1, 2 * P - 1, (1 << 256) - 1, rnd.randrange(P)]) for _ in range(k)]
        B = [rnd.choice([0, 1, P - 1, 2 * P - 1, (1 << 256) - 1, rnd.randrange(P)]) for _ in range(k)]
        L.t_dot9(k, arr(A), arr(B), r)
        ssum = sum(x * y for x, y in zip(A, B))
        assert val(r) % P == ssum * Rinv261 % P and val(r) < P + ssum // (1 << 261) + 1
    for it in range(1000):
        x, y, z = rnd.randrange(P), rnd.randrange(P), rnd.randrange(1 << 256)
        r = (ctypes.c_uint32 * 8)()
        L.t_add(arr([x]), arr([y]), r)
        assert val(r) == (x + y) % P
        L.t_sub(arr([x]), arr([y]), r)
        assert val(r) == (x - y) % P
        a = arr([z])
        L.t_reduce(a)
        assert val(a) == z % P
        assert L.t_num_bits(arr([z])) == z.bit_length()


def test_binary_euclid_inverse_on_host(tmp_path):
    """fr::inv_bea (the inversion of every curve finaliser) compiled for the host, against pow(a, -1, p)."""
    so = tmp_path / "fr_host_shim.so"
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", str(so), os.path.join(ROOT, "tests", "host", "fr_host_shim.cpp")], check=True)
    L = ctypes.CDLL(str(so))
    P = ab.P
    rnd = random.Random(4)
    arr = lambda v: (ctypes.c_uint32 * 8)(*[(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)])
    cases = [1, 2, 3, P - 1, P - 2, 1 << 32, 3 << 32, 5 << 96, 1 << 200, 1 << 253, (P - 1) // 2, (P + 1) // 2, 0xFFFFFFFF]
    cases += [rnd.randrange(1, P) for _ in range(3000)]
    cases += [(rnd.randrange(1, 1 << 60) << rnd.randrange(0, 190)) % P or 1 for _ in range(300)]   # long runs of trailing zeros
    for a in cases:
        r = (ctypes.c_uint32 * 8)()
        L.t_inv_bea(arr(a), r)
        assert sum(int(r[i]) << (32 * i) for i in range(8)) == pow(a, -1, P), hex(a)


def test_plan_heavy_ops_match_oracle():
    b = ab.CircuitBuilder()
    b.hash256("SHA256", [(1, 8), (2, 16), (3, 254)], list(range(10, 42)))
    b.hash256("Keccak256", [(w, 8) for w in range(10, 42)], list(range(50, 82)))
    b.logic("AND", (1, 100), (2, 100), 90)
    b.logic("AND", (3, 120), (2, 120), 91)
    b.fixed_base_scalar_mul((90, 128), (91, 128), (92, 93))
    b.keccak_var([(w, 8) for w in range(50, 60)], (94, 32), list(range(100, 132)))
    b.hash256("Blake2s", [(w, 8) for w in range(100, 132)] + [(1, 254), (2, 254), (3, 70)], list(range(140, 172)))
    b.hash_to_field([(w, 8) for w in range(140, 172)], 180)
    b.hash_to_field([], 181)
    data = b.to_bytes()
    inp = ab.synthetic_inputs(2, n_inputs=3, seed_id=4)
    rows = [inp[i * 96:(i + 1) * 96] + v.to_bytes(32, "big") for i, v in enumerate((4, 11))]
    info = _interp_vs_oracle(data, [1, 2, 3, 94], b"".join(rows), 2)
    assert info["needs_full_kernel"] == 1 and info["n_hash"] == 6 and info["n_curve"] == 1


def _vd_circuit():
    """Value-dependent gates: the unknown is a multiplication operand (arithmetic.rs:217-221 quirks)."""
    b = ab.CircuitBuilder()
    b.arithmetic([(1, 1, 5)], [], ab.P - 7)                    # w1*w5 = 7        : w5 = 7/w1, or (w1==0) unsat
    b.arithmetic([(3, 2, 6)], [(1, 3)], 0)                     # 3*w2*w6 + w3 = 0 : w6 assigned unless w2 == 0 (then needs w3 == 0)
    b.arithmetic([(1, 3, 7)], [(5, 7), (1, 4)], 0)             # w3*w7 + 5*w7 + w4: two entries for w7 unless w3 == 0
    b.arithmetic([], [(1, 6), (ab.P - 1, 8)], 0)               # w8 = w6           : fails where w6 never got assigned
    b.logic("XOR", (5, 254), (1, 254), 9)                      # blackbox over a conditionally assigned input
    b.arithmetic([(1, 9, 9)], [(ab.P - 1, 10)], 0)             # w10 = w9^2
    return b.to_bytes()


def _vd_inputs():
    rows = [(3, 4, 0, 6), (3, 4, 5, 6), (0, 4, 0, 6), (3, 0, 0, 6), (3, 0, 5, 6), (1, ab.P - 5, 0, 9), (2, 0, 0, 0), (7, 9, 0, 0)]
    return rows, b"".join(int(v).to_bytes(32, "big") for r in rows for v in r)


def test_value_dependent_gates_plan_vs_oracle():
    rows, inp = _vd_inputs()
    info = _interp_vs_oracle(_vd_circuit(), [1, 2, 3, 4], inp, len(rows))
    assert info["needs_full_kernel"] == 1


def _dir_mem_circuit():
    """Directives + memory blocks: bit/byte decomposition, euclidean division, dynamic array access."""
    b = ab.CircuitBuilder()
    b.logic("AND", (1, 40), (1, 40), 10)                                            # w10 = low 40 bits of w1
    b.directive_to_le_radix(ab.wexpr(10), list(range(20, 60)), 2)                   # 40 bits
    b.directive_to_le_radix(([], [(1, 10)], 3), list(range(60, 66)), 256)           # bytes of w10 + 3
    b.directive_to_le_radix(ab.wexpr(10), list(range(66, 80)), 10)                  # decimal digits (general radix)
    b.directive_quotient(ab.wexpr(1), ab.wexpr(2), 80, 81)                          # w1 / w2
    b.directive_quotient(ab.wexpr(1), ([], [(1, 10)], 1), 82, 83, predicate=ab.wexpr(3))
    b.arithmetic([(1, 80, 2)], [(1, 81), (ab.P - 1, 1)], 0)                         # check q*b + r == a
    b.memory_init(0, [20, 21, 22, 23, 60, 61, 62, 63])
    b.logic("AND", (10, 3), (10, 3), 84)                                            # index in 0..7
    b.memory_op(0, ab.cexpr(0), ab.wexpr(84), ab.wexpr(85))                         # w85 = mem[w84]
    b.memory_op(0, ab.cexpr(1), ab.wexpr(84), ([], [(5, 85)], 1))                   # mem[w84] = 5*w85 + 1
    b.memory_op(0, ab.cexpr(0), ab.wexpr(84), ab.wexpr(86))                         # w86 = mem[w84]
    b.memory_op(0, ab.cexpr(0), ab.cexpr(3), ab.wexpr(87), predicate=ab.wexpr(3))   # predicated read at a constant index
    b.memory_op(0, ab.cexpr(1), ([], [(1, 84)], 4), ab.wexpr(86), predicate=ab.wexpr(3))  # may go out of bounds (index+4)
    b.memory_op(0, ab.cexpr(0), ab.cexpr(7), ab.wexpr(88))
    return b.to_bytes()


def _dir_mem_inputs():
    rows = [(0x1234567890ABCDEF, 77, 1), (0xFFFFFFFFFF, 1, 0), (5, 0, 1), (ab.P - 1, 3, 7), (0, 9, 0), ((1 << 200) + 12345, (1 << 100) + 7, 1)]
    return rows, b"".join(int(v).to_bytes(32, "big") for r in rows for v in r)


def test_directives_and_memory_plan_vs_oracle(golden):
    rows, inp = _dir_mem_inputs()
    _interp_vs_oracle(_dir_mem_circuit(), [1, 2, 3], inp, len(rows))
    # the reference's own memory_op fixture (acvm_js/test/shared/memory_op.ts)
    fx = golden["acvm_js_shared"]["memory_op"]
    iw = {int(k): int(v, 16) for k, v in fx["initialWitnessMap"].items()}
    keys = sorted(iw)
    _interp_vs_oracle(bytes(fx["bytecode"]), keys, b"".join(iw[k].to_bytes(32, "big") for k in keys), 1)
    # too few output witnesses for the value -> UnsatisfiedConstrain (directives/mod.rs:67-71)
    b = ab.CircuitBuilder()
    b.directive_to_le_radix(ab.wexpr(1), [2, 3, 4], 2)
    _interp_vs_oracle(b.to_bytes(), [1], (7).to_bytes(32, "big") + (8).to_bytes(32, "big") + (0).to_bytes(32, "big"), 3)


def _brillig_circuit():
    """Brillig on the host between device segments: field/int ops, memory, calls, a trap, predicates, recorded foreign calls."""
    R = lambda r: ("Register", r)
    b = ab.CircuitBuilder()
    # w10 = 1/w1 (field div), w11 = w1 + w2 ; then constrained on the device
    b.brillig([("Single", ab.wexpr(1)), ("Single", ab.wexpr(2))], [("Simple", 10), ("Simple", 11)], [
        dict(op="Const", destination=2, value=1),
        dict(op="BinaryFieldOp", destination=3, bop=3, lhs=2, rhs=0),      # r3 = 1 / r0
        dict(op="BinaryFieldOp", destination=1, bop=0, lhs=0, rhs=1),      # r1 = r0 + r1
        dict(op="Mov", destination=0, source=3),
        dict(op="Stop"),
    ])
    b.arithmetic([(1, 1, 10)], [], ab.P - 1)                                # check w1 * w10 == 1 (fails for w1 == 0)
    # array in / array out, integer ops with bit sizes, a loop with Call/Return, Load/Store
    b.brillig([("Array", [ab.wexpr(1), ab.wexpr(2), ([], [(1, 11)], 5)]), ("Single", ab.wexpr(3))],
              [("Array", [12, 13, 14]), ("Simple", 15)], [
        dict(op="Const", destination=2, value=0),                            # i = 0
        dict(op="Const", destination=3, value=3),                            # n = 3
        dict(op="Const", destination=4, value=1),
        dict(op="Const", destination=7, value=100),                          # out base pointer
        dict(op="BinaryIntOp", destination=5, bop=6, bit_size=32, lhs=2, rhs=3),   # 4: i < n
        dict(op="JumpIfNot", condition=5, location=9),
        dict(op="Call", location=12),
        dict(op="BinaryIntOp", destination=2, bop=0, bit_size=32, lhs=2, rhs=4),   # i += 1
        dict(op="Jump", location=4),
        dict(op="Mov", destination=0, source=7),                             # 9: r0 = pointer to the output array
        dict(op="BinaryIntOp", destination=1, bop=4, bit_size=64, lhs=1, rhs=3),   # r1 = (w3 % 2^64) / 3
        dict(op="Stop"),
        dict(op="BinaryIntOp", destination=8, bop=0, bit_size=64, lhs=0, rhs=2),   # 12: src = in_ptr + i
        dict(op="Load", destination=9, source_pointer=8),
        dict(op="BinaryIntOp", destination=9, bop=2, bit_size=16, lhs=9, rhs=9),   # square mod 2^16
        dict(op="BinaryIntOp", destination=9, bop=11, bit_size=16, lhs=9, rhs=4),  # << 1
        dict(op="BinaryIntOp", destination=10, bop=0, bit_size=64, lhs=7, rhs=2),
        dict(op="Store", destination_pointer=10, source=9),
        dict(op="Return"),
    ])
    b.arithmetic([], [(1, 12), (1, 13), (ab.P - 1, 16)], 0)                 # w16 = w12 + w13
    # predicate 0 -> outputs zeroed ; trap when w3 == 7
    b.brillig([("Single", ab.wexpr(3))], [("Simple", 17)], [
        dict(op="Const", destination=1, value=7),
        dict(op="BinaryFieldOp", destination=2, bop=4, lhs=0, rhs=1),
        dict(op="JumpIf", condition=2, location=5),
        dict(op="BinaryFieldOp", destination=0, bop=2, lhs=0, rhs=0),
        dict(op="Stop"),
        dict(op="Trap"),
    ], predicate=ab.wexpr(2))
    # recorded foreign call results are replayed (mod.rs:214-228); hash blackbox inside brillig
    b.brillig([("Array", [([], [(1, 17)], 0), ab.wexpr(1)])], [("Array", list(range(20, 52))), ("Simple", 52)], [
        dict(op="ForeignCall", function="double", destinations=[R(5)], inputs=[R(0)]),
        dict(op="Const", destination=1, value=2),                            # message length
        dict(op="Const", destination=2, value=200),
        dict(op="BlackBox", bb=dict(name="Sha256", message=(0, 1), output=(2, 32))),
        dict(op="Mov", destination=0, source=2),
        dict(op="Mov", destination=1, source=5),
        dict(op="Stop"),
    ], foreign_call_results=[[("Single", 424242)]])
    return b.to_bytes()


def _brillig_inputs():
    rows = [(5, 6, 9), (0, 6, 9), (3, 0, 7), (3, 1, 7), (ab.P - 1, 2, (1 << 70) + 5), (77, 88, 99)]
    return rows, b"".join(int(v).to_bytes(32, "big") for r in rows for v in r)


def test_brillig_segments_plan_vs_oracle(golden):
    rows, inp = _brillig_inputs()
    _interp_vs_oracle(_brillig_circuit(), [1, 2, 3], inp, len(rows))
    # the reference's foreign-call fixtures stop at RequiresForeignCall when no result is recorded
    for name, keys in (("foreign_call", [1]), ("complex_foreign_call", [1, 2, 3])):
        fx = golden["acvm_js_shared"][name]
        iw = {int(k): int(v, 16) for k, v in fx["initialWitnessMap"].items()}
        _interp_vs_oracle(bytes(fx["bytecode"]), keys, b"".join(iw[k].to_bytes(32, "big") for k in keys), 1)


def test_oracle_brillig_foreign_call_fixtures(golden):
    """acvm_js/test/shared/{foreign_call,complex_foreign_call}.ts: resolve the call like the JS test does and compare maps."""
    sh = golden["acvm_js_shared"]
    c = acir.decode_circuit(bytes(sh["foreign_call"]["bytecode"]))
    vm = pwg.ACVM(pwg.OracleBackend(), c.opcodes, {1: 5})
    assert vm.solve() == "RequiresForeignCall" and vm.pending_foreign_call["function"] == "invert"
    vm.resolve_pending_foreign_call([("Single", int(golden["kats"]["inv5"], 16))])
    assert vm.solve() == "Solved"
    assert vm.finalize() == {int(k): int(v, 16) for k, v in sh["foreign_call"]["expectedWitnessMap"].items()}
    c = acir.decode_circuit(bytes(sh["complex_foreign_call"]["bytecode"]))
    vm = pwg.ACVM(pwg.OracleBackend(), c.opcodes, {1: 1, 2: 2, 3: 3})
    assert vm.solve() == "RequiresForeignCall"
    vm.resolve_pending_foreign_call([("Array", [2, 6, 12]), ("Single", 6), ("Single", 12)])
    assert vm.solve() == "Solved"
    assert vm.finalize() == {int(k): int(v, 16) for k, v in sh["complex_foreign_call"]["expectedWitnessMap"].items()}


def _brillig_run_cpp(data, idx, in_vals, n_out):
    import ctypes as C
    buf = b"".join(int(v).to_bytes(32, "big") for v in in_vals)
    out = C.create_string_buffer(max(1, n_out) * 32)
    st, pc = C.c_uint32(), C.c_uint32()
    lib = acvm_b200.lib()
    rc = lib.acvmb_brillig_run_host(data, len(data), idx, buf, len(in_vals), out, n_out, C.byref(st), C.byref(pc))
    assert rc == 0, lib.acvmb_last_error()
    return st.value, pc.value, [int.from_bytes(out.raw[i * 32:(i + 1) * 32], "big") for i in range(n_out)]

def _brillig_run_oracle(br, in_vals):
    from oracle import brillig_vm as obv
    regs, mem, pos = [], [], 0
    for kind, e in br["inputs"]:
        if kind == "Single":
            regs.append(in_vals[pos]); pos += 1
        else:
            regs.append(len(mem)); mem += in_vals[pos:pos + len(e)]; pos += len(e)
    vm = obv.VM(regs, mem, br["bytecode"], br["foreign_call_results"], pwg.OracleBackend())
    try:
        st = vm.process_opcodes()
    except pwg.ReferencePanic:
        return 3, 0, None
    if st[0] == "Failure":
        return 1, st[2][-1], None
    if st[0] == "ForeignCallWait":
        return 2, 0, None
    outs = []
    for i, (kind, o) in enumerate(br["outputs"]):
        reg = vm.get(i)
        if kind == "Simple":
            outs.append(reg)
        else:
            if reg.bit_length() > 64 or reg + len(o) > len(vm.memory):
                return 3, 0, None
            outs += vm.memory[reg:reg + len(o)]
    return 0, 0, outs


def test_cpp_brillig_vm_matches_oracle_vm():
    """The C++ host VM (csrc/brillig_host.hpp) against the oracle's VM on every Brillig opcode of the test circuit and on
    randomly generated integer/field programs (all BinaryIntOp kinds, several bit sizes)."""
    run_cpp, run_oracle = _brillig_run_cpp, _brillig_run_oracle
    rnd = random.Random(3)
    P = ab.P
    # random straight-line programs over 6 registers
    for trial in range(300):
        code = []
        for _ in range(rnd.randrange(1, 12)):
            if rnd.random() < 0.3:
                code.append(dict(op="BinaryFieldOp", destination=rnd.randrange(6), bop=rnd.randrange(5), lhs=rnd.randrange(6), rhs=rnd.randrange(6)))
            else:
                code.append(dict(op="BinaryIntOp", destination=rnd.randrange(6), bop=rnd.randrange(13),
                                 bit_size=rnd.choice([1, 8, 16, 32, 64, 127, 128, 200, 254]), lhs=rnd.randrange(6), rhs=rnd.randrange(6)))
        code.append(dict(op="Stop"))
        b = ab.CircuitBuilder()
        b.brillig([("Single", ab.wexpr(k + 1)) for k in range(4)], [("Simple", 10 + k) for k in range(6)], code)
        data = b.to_bytes()
        br = acir.decode_circuit(data).opcodes[0].body
        ins = [rnd.choice([0, 1, 2, 255, (1 << 64) - 1, 1 << 127, P - 1, rnd.randrange(P), rnd.randrange(1 << 20)]) for _ in range(4)]
        so, pco, oo = run_oracle(br, ins)
        sc, pcc, oc_ = run_cpp(data, 0, ins, 6)
        assert sc == so, (trial, code, ins, sc, so)
        if so == 0:
            assert oc_ == oo, (trial, code, ins)
    # the hand-written programs of the segment test
    data = _brillig_circuit()
    circ = acir.decode_circuit(data)
    for idx, op in enumerate(circ.opcodes):
        if op.kind != "Brillig":
            continue
        n_in = sum(1 if k == "Single" else len(e) for k, e in op.body["inputs"])
        n_out = sum(1 if k == "Simple" else len(o) for k, o in op.body["outputs"])
        for _ in range(20):
            ins = [rnd.choice([0, 1, 7, rnd.randrange(1 << 16), rnd.randrange(P)]) for _ in range(n_in)]
            so, pco, oo = run_oracle(op.body, ins)
            sc, pcc, oc_ = run_cpp(data, idx, ins, n_out)
            assert (sc, pcc if sc == 1 else 0) == (so, pco if so == 1 else 0), (idx, ins)
            if so == 0:
                assert oc_ == oo, (idx, ins)


def test_witness_map_compression_golden(golden):  # acvm_js/test/shared/witness_compression.ts
    fx = golden["acvm_js_shared"]["witness_compression"]
    expected = {int(k): int(v, 16) for k, v in fx["expectedWitnessMap"].items()}
    assert acvm_b200.decompress_witness_map(bytes(fx["expectedCompressedWitnessMap"])) == expected
    ours = acvm_b200.compress_witness_map(expected)
    assert acvm_b200.decompress_witness_map(ours) == expected
    assert acir.decode_witness_map(ours) == expected        # the oracle's independent decoder accepts our bytes


def test_ecdsa_device_routine_on_host(tmp_path, golden):
    """acvm_b200/csrc/ecdsa.cuh is plain C++: the routine each GPU lane runs, compiled for the host, against the oracle."""
    import ecdsa_cases
    from oracle import ecdsa
    so = tmp_path / "ecdsa_host_shim.so"
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", str(so), os.path.join(ROOT, "tests", "host", "ecdsa_host_shim.cpp")], check=True)
    L = ctypes.CDLL(str(so))
    for ci, name in enumerate(("EcdsaSecp256k1", "EcdsaSecp256r1")):
        k = golden["kats"]["ecdsa_valid"][name]
        assert L.t_ecdsa_verify(ci, *[bytes.fromhex(k[a]) for a in ("hashed_message", "pub_key_x", "pub_key_y", "signature")]) == 1
        seen = set()
        for (label, hm, px, py, sig, exp) in ecdsa_cases.cases(name, seed=2, n_random=6):
            assert L.t_ecdsa_verify(ci, hm, px, py, sig) == {True: 1, False: 0, "panic": 2}[exp], (name, label)
            seen.add(exp)
        assert seen == {True, False, "panic"}


@pytest.mark.parametrize("name", ["EcdsaSecp256k1", "EcdsaSecp256r1"])
def test_ecdsa_and_recursive_aggregation_plan_vs_oracle(name):
    import ecdsa_cases
    rnd = random.Random(9)
    cs = ecdsa_cases.cases(name, seed=3, n_random=1)
    inp = b"".join(ecdsa_cases.input_row(c, rnd) for c in cs)
    info = _interp_vs_oracle(ecdsa_cases.circuit(name), ecdsa_cases.INPUTS, inp, len(cs))
    assert info["needs_full_kernel"] == 1 and info["n_curve"] == 1 and not info["static_fail_present"]
    # RecursiveAggregation output that an earlier opcode already assigned: insert_value compares with 0 (mod.rs:154-161)
    _interp_vs_oracle(ecdsa_cases.circuit(name, preassigned_out=True), ecdsa_cases.INPUTS, inp, 3)
    # malformed opcodes: size errors are BlackBoxFunctionFailed, a short hash panics in GenericArray::from_slice
    info = _interp_vs_oracle(ecdsa_cases.circuit(name, n_pkx=31), ecdsa_cases.INPUTS, inp, 1)
    assert info["static_fail_present"] and acvm_b200.solver.ERR_NAMES[info["static_fail_kind"]] == "BlackBoxFunctionFailed"
    info = _interp_vs_oracle(ecdsa_cases.circuit(name, n_hm=31), ecdsa_cases.INPUTS, inp, 1)
    assert info["static_fail_present"] and acvm_b200.solver.ERR_NAMES[info["static_fail_kind"]] == "ReferencePanic"


def test_permutation_route_host_matches_oracle(golden):
    from oracle import sorting
    for k in golden["kats"]["permutation_route"]:
        base = sorted(k["inputs"])
        assert acvm_b200.solver.permutation_route_host([base.index(v) for v in k["outputs"]]) == [bool(b) for b in k["bits"]]
    rnd = random.Random(8)
    assert acvm_b200.solver.permutation_route_host([]) == [] and acvm_b200.solver.permutation_route_host([0]) == []
    for n in list(range(2, 40)) + [63, 64, 65, 200, 1000]:
        for _ in range(3):
            b = list(range(n))
            rnd.shuffle(b)
            assert acvm_b200.solver.permutation_route_host(b) == sorting.route(list(range(n)), b), n
    with pytest.raises(acvm_b200.AcvmError):
        acvm_b200.solver.permutation_route_host([0, 0, 1])


from sort_cases import sort_circuit as _sort_circuit, sort_rows as _sort_rows  # noqa: E402


@pytest.mark.parametrize("n,tup,sort_by", [(5, 1, [0]), (8, 2, [1, 0]), (7, 2, [0]), (2, 1, [0]), (1, 1, [0]), (12, 3, [2, 3]), (6, 1, [1])])
def test_permutation_sort_plan_vs_oracle(n, tup, sort_by):
    data, inputs = _sort_circuit(n, tup, sort_by)
    batch = 6
    inp = _sort_rows(n, tup, batch)
    _interp_vs_oracle(data, inputs, inp, batch)
    # fewer / more bit witnesses than switches: zip() truncates (directives/mod.rs:115)
    for nb in (0, 1, 40):
        data2, _ = _sort_circuit(n, tup, sort_by, n_bits=nb)
        _interp_vs_oracle(data2, inputs, inp, 2)
    data3, _ = _sort_circuit(n, tup, sort_by, preassign_bit=True)
    _interp_vs_oracle(data3, inputs, inp, batch)


def test_cpp_brillig_vm_blackbox_ops(golden):
    """Brillig BlackBox ops on the host VM (brillig_vm/src/black_box.rs:42-165): the three hashes, HashToField128Security
    and both ECDSA curves, against the oracle's VM."""
    import ecdsa_cases
    rnd = random.Random(21)
    # message in memory[0..n), digest to memory[100..132), field hash to a register
    for n in (0, 1, 55, 64, 65, 130):
        msg = [rnd.randrange(1 << 20) for _ in range(n)]     # only the low byte of each value is hashed
        code = [dict(op="Const", destination=1, value=n), dict(op="Const", destination=2, value=200)]
        for k, name in enumerate(("Sha256", "Blake2s", "Keccak256")):
            code.append(dict(op="Const", destination=3 + k, value=200 + 32 * k))
            code.append(dict(op="BlackBox", bb=dict(name=name, message=(0, 1), output=(3 + k, 32))))
        code.append(dict(op="BlackBox", bb=dict(name="HashToField128Security", message=(0, 1), output=7)))
        code.append(dict(op="Mov", destination=0, source=2))
        code.append(dict(op="Mov", destination=1, source=7))
        code.append(dict(op="Stop"))
        b = ab.CircuitBuilder()
        # memory must exist up to 296: pass a 300-element array as the first input (register 0 = pointer 0)
        b.brillig([("Array", [ab.wexpr(k + 1) for k in range(300)])], [("Array", list(range(400, 496))), ("Simple", 500)], code)
        data = b.to_bytes()
        br = acir.decode_circuit(data).opcodes[0].body
        ins = msg + [0] * (300 - n)
        so, _, oo = _brillig_run_oracle(br, ins)
        sc, _, oc_ = _brillig_run_cpp(data, 0, ins, 97)
        assert (so, sc) == (0, 0) and oc_ == oo, n
    # ECDSA: memory = pkx | pky | sig | hashed
    for name in ("EcdsaSecp256k1", "EcdsaSecp256r1"):
        code = [dict(op="Const", destination=1, value=32), dict(op="Const", destination=2, value=64), dict(op="Const", destination=3, value=128),
                dict(op="Const", destination=4, value=32),
                dict(op="BlackBox", bb=dict(name=name, hashed_msg=(3, 4), public_key_x=(0, 32), public_key_y=(1, 32), signature=(2, 64), result=5)),
                dict(op="Mov", destination=0, source=5), dict(op="Stop")]
        b = ab.CircuitBuilder()
        b.brillig([("Array", [ab.wexpr(k + 1) for k in range(160)])], [("Simple", 300)], code)
        data = b.to_bytes()
        br = acir.decode_circuit(data).opcodes[0].body
        seen = set()
        for c in ecdsa_cases.cases(name, seed=13, n_random=1):
            ins = list(c[2]) + list(c[3]) + list(c[4]) + list(c[1])
            so, _, oo = _brillig_run_oracle(br, ins)
            sc, _, oc_ = _brillig_run_cpp(data, 0, ins, 1)
            assert sc == so == (3 if c[5] == "panic" else 0), c[0]
            if so == 0:
                assert oc_ == oo == [1 if c[5] else 0], c[0]
            seen.add(c[5])
        assert seen == {True, False, "panic"}
        # wrong array size: BlackBoxResolutionError::Failed -> VM failure, not a panic
        code[4] = dict(op="BlackBox", bb=dict(name=name, hashed_msg=(3, 4), public_key_x=(0, 31), public_key_y=(1, 32), signature=(2, 64), result=5))
        b = ab.CircuitBuilder()
        b.brillig([("Array", [ab.wexpr(k + 1) for k in range(160)])], [("Simple", 300)], code)
        data = b.to_bytes()
        so, pco, _ = _brillig_run_oracle(acir.decode_circuit(data).opcodes[0].body, ins)
        sc, pcc, _ = _brillig_run_cpp(data, 0, ins, 1)
        assert (so, pco) == (sc, pcc) == (1, 4)


@pytest.mark.parametrize("S", [4, 8, 16, 32])
def test_curve_ops_monolithic_and_split_plans_vs_oracle(S):
    """FixedBaseScalarMul / Pedersen: one micro-op per call at S < 8, partial-sum / addition-tree / finaliser micro-ops at
    S >= 8 (plan.cpp curve_parts / curve_reduce / curve_final) -- both against the oracle, including chained calls, an
    empty Pedersen, limbs that fail validation, the zero scalar and a pre-assigned output."""
    from oracle import grumpkin
    b = ab.CircuitBuilder()
    b.pedersen([(1, 254), (2, 254)], 0, (10, 11))
    b.pedersen([(10, 254)], 5, (12, 13))
    b.pedersen([], 1, (14, 15))
    b.pedersen([(1, 254), (11, 254), (3, 254)], 1023, (16, 17))
    b.logic("AND", (1, 100), (2, 100), 18)
    b.logic("AND", (3, 100), (2, 100), 19)
    b.fixed_base_scalar_mul((18, 128), (19, 128), (20, 21))
    b.arithmetic([(1, 20, 21)], [(1, 16), (ab.P - 1, 24)], 0)
    b.pedersen([(12, 254), (13, 254)], 0, (30, 31))          # same separator and input count as the first call: its constant
    b.pedersen([(30, 254), (2, 254)], 0, (32, 33))           # points H0(IV), H1(n) are computed once per plan (const_curve_point)
    b.fixed_base_scalar_mul((1, 254), (19, 128), (22, 23))   # low limb >= 2^128 in the random rows: BlackBoxFunctionFailed
    data = b.to_bytes()
    with pytest.raises(acvm_b200.AcvmError) as e:      # Pedersen is opt-in: parity with barretenberg's tables is unpinned
        acvm_b200.compile_plan_host(data, [1, 2, 3], S)
    assert e.value.rc == -5 and "pedersen_unpinned" in str(e.value)
    info = _interp_vs_oracle(data, [1, 2, 3], ab.synthetic_inputs(3, n_inputs=3, seed_id=9), 3, S=S, pedersen_unpinned=True)
    assert info["n_curve"] == 8 and (info["n_micro_ops"] > 100) == (S >= 8)
    b = ab.CircuitBuilder()
    b.fixed_base_scalar_mul((1, 128), (2, 128), (4, 3))      # y is pre-assigned: insert_value compares
    b.arithmetic([], [(1, 4), (ab.P - 1, 6)], 1)
    g5 = grumpkin.fixed_base_scalar_mul(5, 0)
    rows = [[0, 0, 0], [5, 0, g5[1]], [1 << 128, 0, 7]]
    _interp_vs_oracle(b.to_bytes(), [1, 2, 3], b"".join(int(v).to_bytes(32, "big") for r in rows for v in r), 3, S=S)


def test_values_that_outlive_the_temporary_pool():
    """Round-1 advisor finding: the round-robin temporaries wrapped over values that were still live (Brillig inputs wait for
    their host segment, the H1 points of a Pedersen call for their chaining round).  Such values now get dedicated slots."""
    # 2100 expression inputs (> the 2048-slot pool) to a host-VM Brillig op; outputs = the first three inputs
    b = ab.CircuitBuilder()
    b.brillig([("Array", [([], [(1, 1)], i + 1) for i in range(2100)])], [("Array", [10, 11, 12])], [dict(op="Stop")])
    data = b.to_bytes()
    for dev in (False, True):
        info, blob = acvm_b200.compile_plan_host(data, [1], 16, device_brillig=dev)
        assert info["n_host_segments"] == (0 if dev else 1)
        st, wm = plan_interp.run_plan(plan_interp.PlanBlob(blob), {1: 5}, circuit=acir.decode_circuit(data))
        assert st[0] == "Solved" and (wm[10], wm[11], wm[12]) == (6, 7, 8)
    # Pedersen with enough inputs for the H1 partial sums to wrap a small pool several times
    b = ab.CircuitBuilder()
    b.pedersen([(1 + (k % 3), 254) for k in range(24)], 2, (10, 11))
    info = _interp_vs_oracle(b.to_bytes(), [1, 2, 3], ab.synthetic_inputs(1, n_inputs=3, seed_id=12), 1, S=8, temp_pool=96,
                             pedersen_unpinned=True)
    assert info["n_micro_ops"] > 24 * 15


def test_straight_line_brillig_is_lowered_to_device_gates():
    """stdlib emits `bytecode: vec![Stop]` constant loads and one-instruction field ops by the thousand
    (stdlib/src/blackbox_fallbacks/uint.rs:51-63,85, hash_to_field.rs:107): no host segment for those."""
    b = ab.CircuitBuilder()
    b.brillig([("Single", ([], [], 1 << 32))], [("Simple", 10)], [dict(op="Stop")])                   # load_constant
    b.brillig([("Single", ab.wexpr(1)), ("Single", ab.wexpr(2))], [("Simple", 11)],
              [dict(op="BinaryFieldOp", destination=0, bop=0, lhs=0, rhs=1)])                         # runs off the end: Finished
    b.brillig([("Single", ab.wexpr(1)), ("Single", ([(3, 1, 2)], [(1, 10)], 7)), ("Array", [ab.wexpr(3), ab.wexpr(11)])],
              [("Simple", 12), ("Array", [13, 14]), ("Simple", 15), ("Simple", 16)], [
        dict(op="BinaryFieldOp", destination=3, bop=2, lhs=0, rhs=1),       # r3 = r0 * r1
        dict(op="Const", destination=4, value=9),
        dict(op="BinaryFieldOp", destination=0, bop=1, lhs=3, rhs=4),       # r0 = r3 - 9
        dict(op="Mov", destination=1, source=2),                            # r1 = pointer to the input array (0)
        dict(op="BinaryFieldOp", destination=2, bop=1, lhs=0, rhs=0),       # r2 = 0
        dict(op="Stop"),
        dict(op="Trap"),
    ])
    b.arithmetic([], [(1, 12), (1, 13), (ab.P - 1, 17)], 0)
    # an output that is already assigned is compared (insert_value): w11 again, equal by construction; then a mismatch
    b.brillig([("Single", ab.wexpr(11))], [("Simple", 11)], [dict(op="Stop")])
    b.brillig([("Single", ab.wexpr(3))], [("Simple", 1)], [dict(op="Stop")])                          # fails unless w3 == w1
    data = b.to_bytes()
    rows = [(5, 6, 9), (0, 0, 0), (7, 8, 7), (ab.P - 1, 2, 3)]
    inp = b"".join(int(v).to_bytes(32, "big") for r in rows for v in r)
    info = _interp_vs_oracle(data, [1, 2, 3], inp, len(rows))
    assert info["n_brillig"] == 5 and info["n_brillig_device"] == 5 and info["n_host_segments"] == 0
    info = _interp_vs_oracle(data, [1, 2, 3], inp, len(rows), device_brillig=False)
    assert info["n_brillig_device"] == 0 and info["n_host_segments"] == 5
    # not straight-line code (a jump, a predicate): stays on the host VM
    b = ab.CircuitBuilder()
    b.brillig([("Single", ab.wexpr(1)), ("Single", ab.wexpr(2))], [("Simple", 10)],
              [dict(op="Jump", location=1), dict(op="BinaryFieldOp", destination=0, bop=0, lhs=0, rhs=1)])
    b.brillig([("Single", ab.wexpr(1))], [("Simple", 11)], [dict(op="Stop")], predicate=ab.wexpr(2))
    info = _interp_vs_oracle(b.to_bytes(), [1, 2], inp[:64] + inp[96:160], 2)
    assert info["n_brillig_device"] == 0 and info["n_host_segments"] == 2


@pytest.mark.parametrize("W", [4, 24, 96])
def test_ring_of_recent_values_plan_vs_oracle(W):
    """Opt-in shared-memory ring (plan.cpp assign_ring): operand fields flagged to read ring entries must name the entry that
    holds the value when the step runs -- the interpreter models the ring, asserts that no entry is read and rewritten in one
    step, and the witness map must still equal the oracle's."""
    data, inputs, _ = ab.synthetic_arith_circuit(900, mode="local", coeffs="noir-like", seed_id=3)
    # (a step of S slots may write S values: with W <= S nothing written in one step survives the next one)
    info = _interp_vs_oracle(data, inputs, ab.synthetic_inputs(2, seed_id=3), 2, 2 if W == 4 else 16, ring_slots=W)
    assert info["ring_slots"] == W and 0 < info["n_ring_reads"] <= info["n_operand_reads"]
    b = ab.CircuitBuilder()
    b.logic("AND", (1, 64), (2, 64), 10)
    b.logic("XOR", (10, 64), (1, 64), 11)
    b.range((11, 64))
    b.arithmetic([(3, 10, 11)], [(1, 1), (ab.P - 1, 12)], 7)
    b.arithmetic([], [(1, 12), (1, 11), (ab.P - 1, 13)], 0)
    b.logic("AND", (13, 254), (12, 254), 14)
    b.arithmetic([(1, 14, 14)], [(ab.P - 1, 12)], 0)      # check, fails for random inputs
    info = _interp_vs_oracle(b.to_bytes(), [1, 2], ab.synthetic_inputs(3, n_inputs=2, seed_id=6), 3, 2, ring_slots=W)
    assert info["n_ring_reads"] > 0


def _int_brillig_circuit():
    """stdlib uint fallbacks (stdlib/src/blackbox_fallbacks/uint.rs:220-330, 510-600): one-instruction BinaryIntOp Brillig
    opcodes with bit_size 127 feeding arithmetic constraints -- plus every other integer op the device lowering covers."""
    b = ab.CircuitBuilder()
    nxt = 10
    ops = [(0, 127), (1, 127), (2, 127), (4, 64), (5, 32), (6, 32), (7, 32), (8, 100), (9, 100), (10, 100), (11, 64), (12, 64),
           (0, 8), (2, 128), (1, 1)]
    for bop, bits in ops:
        b.brillig([("Single", ab.wexpr(1)), ("Single", ab.wexpr(2))], [("Simple", nxt)],
                  [dict(op="BinaryIntOp", destination=0, bop=bop, bit_size=bits, lhs=0, rhs=1)])
        nxt += 1
    # uint.rs sub: (a + 2^width) - b with a constant in register 2, then a field op on the result
    b.brillig([("Single", ab.wexpr(3)), ("Single", ab.wexpr(4)), ("Single", ([], [], 1 << 32))], [("Simple", nxt), ("Simple", nxt + 1)], [
        dict(op="BinaryIntOp", destination=0, bop=0, bit_size=127, lhs=0, rhs=2),
        dict(op="BinaryIntOp", destination=0, bop=1, bit_size=127, lhs=0, rhs=1),
        dict(op="BinaryFieldOp", destination=1, bop=2, lhs=0, rhs=0),
        dict(op="Const", destination=3, value=5),
        dict(op="BinaryIntOp", destination=1, bop=12, bit_size=64, lhs=1, rhs=3),     # (r0^2) >> 5, mod 2^64
    ])
    b.arithmetic([], [(1, nxt), (1, nxt + 1), (ab.P - 1, nxt + 2)], 0)
    return b.to_bytes(), nxt + 3


def _int_brillig_inputs():
    rows = [(5, 3, 9, 4), (3, 5, 4, 9), (ab.P - 1, 2, (1 << 70) + 5, 77), (1 << 127, (1 << 127) - 1, 0, 0), (12345678901234567890, 0, 1, 1),
            (7, 300, 1 << 40, 1 << 33), ((1 << 200) + 17, (1 << 130) + 3, 5, 6)]
    return rows, b"".join(int(v).to_bytes(32, "big") for r in rows for v in r)


def test_integer_brillig_ops_are_lowered_to_the_device():
    data, _ = _int_brillig_circuit()
    rows, inp = _int_brillig_inputs()
    info = _interp_vs_oracle(data, [1, 2, 3, 4], inp, len(rows))
    assert info["n_brillig"] == 16 and info["n_brillig_device"] == 16 and info["n_host_segments"] == 0
    assert info["needs_full_kernel"] == 1
    info = _interp_vs_oracle(data, [1, 2, 3, 4], inp, len(rows), device_brillig=False)
    assert info["n_host_segments"] == 16
    # SignedDiv and bit sizes above 128 stay on the host VM
    b = ab.CircuitBuilder()
    b.brillig([("Single", ab.wexpr(1)), ("Single", ab.wexpr(2))], [("Simple", 10)],
              [dict(op="BinaryIntOp", destination=0, bop=3, bit_size=8, lhs=0, rhs=1)])
    b.brillig([("Single", ab.wexpr(1)), ("Single", ab.wexpr(2))], [("Simple", 11)],
              [dict(op="BinaryIntOp", destination=0, bop=0, bit_size=200, lhs=0, rhs=1)])
    info = _interp_vs_oracle(b.to_bytes(), [1, 2], inp[:64] + inp[128:192], 2)
    assert info["n_brillig_device"] == 0 and info["n_host_segments"] == 2


@pytest.mark.parametrize("packed", [True, False])
def test_packed_hash_pipeline_plan_vs_oracle(packed):
    """SHA256 / Keccak256 / Blake2s over byte inputs are lowered to pack / core / unpack micro-ops (plan.cpp hash_packed): block
    edges, partial chunks, a digest handed to the next call as one packed column (the config-3 chain), a pre-assigned
    output that does not match, and the same circuits with the one-micro-op form."""
    data, inputs, nw = ab.hash_chain_circuit(6)
    rng = random.Random(3)
    inp = b"".join(rng.randrange(256).to_bytes(32, "big") for _ in range(2 * len(inputs)))
    info = _interp_vs_oracle(data, inputs, inp, 2, 16, packed_hashes=packed)
    assert info["n_hash"] == 6 and (info["n_micro_ops"] > 6) == packed
    if packed:   # 6 cores + 6 unpacks + packs: two for the first call (nothing to hand over), one (the fresh half) afterwards
        assert info["n_micro_ops"] == 6 + 6 + 2 + 5
    for name, lengths in (("SHA256", (1, 31, 32, 33, 55, 56, 63, 64, 65, 119, 120, 128, 200)), ("Keccak256", (1, 32, 64, 100, 135, 136, 200)),
                          ("Blake2s", (1, 32, 63, 64, 65, 128, 129))):
        b = ab.CircuitBuilder()
        nxt = 300
        for n in lengths:
            b.hash256(name, [(1 + (k * 7 + n) % 200, 8 if k % 3 else 5) for k in range(n)], list(range(nxt, nxt + 32)))
            nxt += 32
        b.hash256(name, [(w, 8) for w in range(300, 332)] + [(3, 8)], list(range(nxt, nxt + 32)))   # digest of call 0 + one byte
        ins = list(range(1, 201))
        inp = b"".join(rng.randrange(256).to_bytes(32, "big") for _ in range(len(ins)))
        _interp_vs_oracle(b.to_bytes(), ins, inp, 1, 16, packed_hashes=packed)
    # byte-typed inputs holding full-width field values: only the low byte is hashed (fetch_nearest_bytes, generic_ark.rs:305-317)
    b = ab.CircuitBuilder()
    b.hash256("SHA256", [(1 + k % 7, 8 if k % 2 else 3) for k in range(40)], list(range(300, 332)))
    b.hash256("Keccak256", [(w, 8) for w in range(300, 332)] + [(2, 8), (2, 8)], list(range(340, 372)))
    _interp_vs_oracle(b.to_bytes(), list(range(1, 8)), ab.synthetic_inputs(2, n_inputs=7, seed_id=31), 2, 16, packed_hashes=packed)
    # more than 37 chunks: the core's descriptor no longer fits the record and goes to the payload
    b = ab.CircuitBuilder()
    b.hash256("SHA256", [(1 + (k * 11) % 200, 8) for k in range(1250)], list(range(300, 332)))
    ins = list(range(1, 201))
    _interp_vs_oracle(b.to_bytes(), ins, b"".join(rng.randrange(256).to_bytes(32, "big") for _ in range(len(ins))), 1, 16,
                      packed_hashes=packed)
    b = ab.CircuitBuilder()
    b.arithmetic([], [(1, 2), (ab.P - 1, 13)], 0)                 # w13 := w2: output 3 of the hash is pre-assigned
    b.hash256("SHA256", [(1, 8)], list(range(10, 42)))
    good = hashlib.sha256(b"\x05").digest()
    data = b.to_bytes()
    _interp_vs_oracle(data, [1, 2], (5).to_bytes(32, "big") + good[3].to_bytes(32, "big"), 1, 16, packed_hashes=packed)
    # the mismatching instance: same status as the oracle; its witness map is the documented deviation
    info, blob = acvm_b200.compile_plan_host(data, [1, 2], 16, packed_hashes=packed)
    st, wm = plan_interp.run_plan(plan_interp.PlanBlob(blob), {1: 5, 2: (good[3] + 1) % 256}, circuit=acir.decode_circuit(data))
    assert st[0] == "Failure" and acvm_b200.solver.ERR_NAMES[st[1]] == "UnsatisfiedConstrain" and st[2] == 1
    assert wm == {1: 5, 2: (good[3] + 1) % 256, 13: good[3]}


def test_heavy_micro_ops_of_a_step_sit_in_different_warps():
    """Scheduler::emit (plan.cpp): with a tile narrower than a warp (S = 16 -> T = 8, four slots per warp) the hash core and the
    pack / unpack micro-ops beside it take the first slot of different warps; light ops keep consecutive slots."""
    data, inputs, nw = ab.hash_chain_circuit(6)
    HEAVY = {25, 26, 27}   # MK_HASH_PACK / MK_HASH_CORE / MK_HASH_UNPACK (plan.hpp)
    for spread in (True, False):
        info, blob = acvm_b200.compile_plan_host(data, inputs, 16, spread_heavy=spread)
        plan = plan_interp.PlanBlob(blob)
        seen_multi = False
        for s in range(plan.n_steps):
            kinds = [plan.record(s * plan.S + j)[0][0] & 0xFF for j in range(plan.S)]
            pos = [j for j, k in enumerate(kinds) if k in HEAVY]
            assert all(k in HEAVY or k == 0 for k in kinds)
            if len(pos) > 1:
                seen_multi = True
                warps = [j // 4 for j in pos]
                if spread:
                    per_warp = [warps.count(w) for w in range(4)]
                    assert max(per_warp) - min(per_warp) <= 1            # one per warp before any warp gets a second
                    if len(pos) <= 4:
                        assert all(j % 4 == 0 for j in pos)
                else:
                    assert pos == list(range(len(pos)))
        assert seen_multi
    rng = random.Random(5)
    inp = b"".join(rng.randrange(256).to_bytes(32, "big") for _ in range(len(inputs)))
    _interp_vs_oracle(data, inputs, inp, 1, 16, spread_heavy=False)


def _dynamic_memory_selector_circuits():
    """MemoryOp whose `operation` is a witness (memory_op.rs:68-81 evaluates it per instance)."""
    a = ab.CircuitBuilder()          # value = one unassigned witness: read lanes succeed, write lanes are MissingAssignment(w10)
    a.memory_init(0, [4, 5, 6, 7])
    a.memory_op(0, ab.wexpr(3), ab.wexpr(2), ab.wexpr(10))
    a.arithmetic([], [(2, 10), (ab.P - 1, 11)], 3)                 # w11 = 2*w10 + 3
    b = ab.CircuitBuilder()          # value fully assigned: write lanes succeed, read lanes hit the reference's expect() panic
    b.memory_init(0, [4, 5, 6, 7])
    b.memory_op(0, ([], [(1, 3)], 0), ab.wexpr(2), ([], [(5, 1)], 1), predicate=ab.wexpr(8))
    b.memory_op(0, ab.cexpr(0), ab.wexpr(2), ab.wexpr(12))         # observe the write
    return a.to_bytes(), b.to_bytes()


def _dynamic_memory_selector_inputs():
    #        w1  idx sel  w4..w7 (block)     w8 (predicate)
    rows = [(9, 0, 0, 40, 50, 60, 70, 1), (9, 3, 1, 40, 50, 60, 70, 1), (9, 2, 2, 40, 50, 60, 70, 0), (9, 7, 0, 40, 50, 60, 70, 1),
            (9, 7, 1, 40, 50, 60, 70, 1), (9, 1, ab.P - 1, 40, 50, 60, 70, 1), (9, 1 << 70, 1, 40, 50, 60, 70, 1)]
    return rows, b"".join(int(v).to_bytes(32, "big") for r in rows for v in r)


def test_memory_op_with_witness_dependent_selector_plan_vs_oracle():
    rows, inp = _dynamic_memory_selector_inputs()
    for data in _dynamic_memory_selector_circuits():
        _interp_vs_oracle(data, list(range(1, 9)), inp, len(rows))
    # neither readable nor assigned (and the predicated read, which could leave its witness unassigned): refused loudly
    c = ab.CircuitBuilder()
    c.memory_init(0, [4, 5, 6, 7])
    c.memory_op(0, ab.wexpr(3), ab.wexpr(2), ab.wexpr(10), predicate=ab.wexpr(8))
    with pytest.raises(acvm_b200.AcvmError) as e:
        acvm_b200.compile_plan_host(c.to_bytes(), list(range(1, 9)), 16)
    assert e.value.rc == -5 and "selector" in str(e.value)


@pytest.mark.parametrize("seed_id", [4, 5])
def test_mixed_circuits_schedule_variants_vs_oracle(seed_id):
    """BASELINE config-4 shaped circuits (gates, RANGE/AND/XOR, hashes, Pedersen, fixed-base with local operands) through every
    scheduling variant of the plan compiler -- slack scheduling of curve micro-ops, heavy micro-ops spread over warps, packed
    hashes -- against the oracle: any order the scheduler picks must compute the same witness map."""
    data, inputs, nw, counts = ab.mixed_circuit(700, seed_id=seed_id, window=24)
    assert counts["pedersen"] + counts["fixed_base"] >= 2 and counts["hash"] >= 2
    inp = ab.synthetic_inputs(1, n_inputs=len(inputs), seed_id=40 + seed_id)
    for slack, spread, packed in ((True, True, True), (False, True, True), (True, False, False)):
        _interp_vs_oracle(data, inputs, inp, 1, 8, pedersen_unpinned=True, slack_scheduling=slack, spread_heavy=spread, packed_hashes=packed)
    if seed_id == 4:   # the config-2 chain: every H1 sum over a fresh input has slack, the constant sums are shared by all calls
        data, inputs, nw = ab.pedersen_chain_circuit(7, n_fresh=3)
        inp = ab.synthetic_inputs(1, n_inputs=len(inputs), seed_id=50)
        infos = [_interp_vs_oracle(data, inputs, inp, 1, 8, pedersen_unpinned=True, slack_scheduling=slack) for slack in (True, False)]
        assert infos[0]["n_micro_ops"] == infos[1]["n_micro_ops"] and infos[0]["n_steps"] < infos[1]["n_steps"]


def test_interpreter_detects_a_same_step_race():
    """The kernel has one barrier per step, so a step must never hold a reader (or a second writer) of a column that another of
    its slot threads writes.  tests/plan_interp.py asserts that for every step of every plan it runs; this test corrupts a
    schedule to show the assertion fires."""
    b = ab.CircuitBuilder()
    b.arithmetic([], [(1, 1), (ab.P - 1, 3)], 0)      # w3 = w1
    b.arithmetic([], [(1, 3), (ab.P - 1, 4)], 1)      # w4 = w3 + 1: must sit in a later step
    data = b.to_bytes()
    info, blob = acvm_b200.compile_plan_host(data, [1], 4)
    plan = plan_interp.PlanBlob(blob)
    assert plan.n_steps >= 2
    st, wm = plan_interp.run_plan(plan, {1: 5}, circuit=acir.decode_circuit(data))
    assert st == ("Solved",) and wm == {1: 5, 3: 5, 4: 6}
    raw = bytearray(plan.stream)
    S = plan.S
    raw[192:384] = raw[S * 192:(S + 1) * 192]        # the dependent gate moves into step 0, slot 1
    raw[S * 192:(S + 1) * 192] = bytes(192)
    plan.stream = bytes(raw)
    with pytest.raises(AssertionError, match="in the same step"):
        plan_interp.run_plan(plan, {1: 5}, circuit=acir.decode_circuit(data))


def test_medium_mixed_circuit_schedule_is_race_free():
    """5000 config-4 shaped opcodes (33 Pedersen, 27 fixed-base, 83 hash calls among gates and logic ops) through the plan
    interpreter: too long for the Python oracle, but the interpreter's per-step race assertion and a Solved status (every
    gate CHECK of the circuit holds on the computed values) cover the schedule itself."""
    data, inputs, nw, counts = ab.mixed_circuit(5000, seed_id=7, window=64)
    assert counts["pedersen"] > 20 and counts["hash"] > 50
    info, blob = acvm_b200.compile_plan_host(data, inputs, 8, pedersen_unpinned=True)
    plan = plan_interp.PlanBlob(blob)
    iw = inputs_to_dicts(ab.synthetic_inputs(1, n_inputs=len(inputs), seed_id=3), 1, inputs)[0]
    st, wm = plan_interp.run_plan(plan, iw, circuit=acir.decode_circuit(data))
    assert st == ("Solved",) and len(wm) > 7000
