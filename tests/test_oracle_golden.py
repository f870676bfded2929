"""Pins the oracle against every golden vector the reference's own tests hold for this path."""
import hashlib
import os

import pytest

from oracle import acir, field as F, grumpkin, hashes, pwg


def test_fr_hex_vectors(golden):  # acir_field/src/generic_ark.rs:424-438
    k = golden["kats"]["fr_hex"]
    assert F.to_hex(0) == k["0"]
    assert F.to_hex(F.neg(1)) == k["-1"]
    assert F.to_hex(F.neg(2)) == k["-2"]
    assert F.to_hex(F.neg(3)) == k["-3"]


def test_max_num_bits():  # generic_ark.rs:440-443
    assert F.MAX_NUM_BITS == 254 and F.num_bits(F.P - 1) == 254


def test_and_self(golden):  # generic_ark.rs:411-421
    for x in range(0, 10000, 7):
        assert F.and_(x, x, 254) == x


def test_inverse_of_5(golden):  # acvm_js/test/shared/foreign_call.ts:20-27
    assert F.to_hex(F.inverse(5)) == golden["kats"]["inv5"]
    assert F.inverse(0) == 0


def test_sha256_kat(golden):  # brillig_vm/src/black_box.rs:203-209
    assert hashes.sha256(b"hello world").hex() == golden["kats"]["sha256_hello_world"]


def test_keccak_permutation_pinned_by_sha3():
    assert hashes.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    for n in (0, 1, 55, 135, 136, 137, 271, 272, 500):
        m = bytes((i * 7 + n) & 0xFF for i in range(n))
        assert hashes._sha3_256_via_own_permutation(m) == hashlib.sha3_256(m).digest()


def test_fixed_base_kats(golden):  # barretenberg_blackbox_solver/src/wasm/scalar_mul.rs:72-97
    for k in golden["kats"]["fixed_base"]:
        x, y = grumpkin.fixed_base_scalar_mul(k["low"], k["high"])
        assert (F.to_hex(x), F.to_hex(y)) == (k["x"], k["y"])
    assert grumpkin.mul(grumpkin.ORDER, grumpkin.G) is None
    with pytest.raises(grumpkin.BlackBoxFailed):
        grumpkin.fixed_base_scalar_mul(1 << 128, 0)
    with pytest.raises(grumpkin.BlackBoxFailed):
        grumpkin.fixed_base_scalar_mul(grumpkin.ORDER & ((1 << 128) - 1), grumpkin.ORDER >> 128)


def test_all_golden_circuits_decode(golden):  # acir/tests/test_program_serialization.rs (7 byte vectors)
    kinds = {}
    for name, v in golden["rust_serialization"].items():
        c = acir.decode_circuit(bytes(v))
        kinds[name] = [o.kind for o in c.opcodes]
    assert kinds["addition_circuit"] == ["Arithmetic"]
    assert kinds["memory_op_circuit"] == ["MemoryInit", "MemoryOp", "MemoryOp"]
    assert kinds["schnorr_verify_circuit"] == ["BlackBoxFuncCall"]
    assert len(kinds) == 7


def test_witness_map_golden(golden):  # acvm_js/test/shared/witness_compression.ts
    fx = golden["acvm_js_shared"]["witness_compression"]
    wm = acir.decode_witness_map(bytes(fx["expectedCompressedWitnessMap"]))
    assert wm == {int(k): int(v, 16) for k, v in fx["expectedWitnessMap"].items()}


@pytest.mark.parametrize("name", ["fixed_base_scalar_mul", "memory_op"])
def test_end_to_end_fixtures(golden, name):  # acvm_js/test/shared/*.ts expected witness maps
    fx = golden["acvm_js_shared"][name]
    c = acir.decode_circuit(bytes(fx["bytecode"]))
    iw = {int(k): int(v, 16) for k, v in fx["initialWitnessMap"].items()}
    st, wm, err = pwg.solve_circuit(c, iw)
    assert st == "Solved", err
    assert wm == {int(k): int(v, 16) for k, v in fx["expectedWitnessMap"].items()}


def test_addition_fixture(golden):
    fx = golden["acvm_js_shared"]["addition"]
    c = acir.decode_circuit(bytes(fx["bytecode"]))
    st, wm, _ = pwg.solve_circuit(c, {int(k): int(v, 16) for k, v in fx["initialWitnessMap"].items()})
    assert st == "Solved" and wm[fx["resultWitness"]] == int(fx["expectedResult"], 16)


def test_arithmetic_smoke():  # acvm/src/pwg/arithmetic.rs:242-281
    E = acir.Expression
    wm = {2: 2, 3: 1, 4: 1}  # b, c, d
    pwg.solve_arithmetic(wm, E([], [(1, 1), (F.neg(1), 2), (F.neg(1), 3), (F.neg(1), 4)], 0))  # a = b + c + d
    assert wm[1] == 4
    pwg.solve_arithmetic(wm, E([], [(1, 1), (F.neg(1), 5)], 0))  # e = a
    assert wm[5] == 4


def test_unsatisfied_opcode_resolved():  # acvm/tests/solver.rs:490-525
    E = acir.Expression
    ops = [acir.Opcode("Arithmetic", E([], [(1, 1), (F.neg(1), 2)], 0))]
    vm = pwg.ACVM(pwg.StubbedBackend(), ops, {1: 5, 2: 6})
    assert vm.solve() == "Failure"
    assert vm.error.kind == "UnsatisfiedConstrain" and vm.error.opcode_location == 0


def test_quirks():
    E = acir.Expression
    # same unknown twice after evaluate -> two unknowns (arithmetic.rs:188-201)
    with pytest.raises(pwg.ResolutionError) as e:
        pwg.solve_arithmetic({1: 3}, E([(1, 1, 2)], [(1, 2)], 0))
    assert e.value.kind.endswith("TooManyUnknowns")
    # ... but when the known operand is 0 the mul-derived entry is dropped and it solves (arithmetic.rs:217-221)
    wm = {1: 0}
    pwg.solve_arithmetic(wm, E([(1, 1, 2)], [(1, 2)], 5))
    assert wm[2] == F.neg(5)
    # value-dependent drop leaves the witness unassigned
    wm = {1: 0}
    pwg.solve_arithmetic(wm, E([(1, 1, 2)], [], 0))
    assert 2 not in wm
    # >1 surviving mul term -> panic (arithmetic.rs:142)
    with pytest.raises(pwg.ReferencePanic):
        pwg.solve_arithmetic({}, E([(1, 1, 2), (1, 3, 4)], [], 0))


def test_cpp_restatement_matches_python_oracle():
    from acvm_b200 import acir_builder as ab
    from oracle import cref
    cref.build()
    for mode, coeffs in (("local", "dense"), ("global", "noir-like")):
        data, inputs, nw = ab.synthetic_arith_circuit(300, mode=mode, coeffs=coeffs)
        c = acir.decode_circuit(data)
        inp = ab.synthetic_inputs(3)
        res, ow, op = cref.solve_batch(c, inputs, inp, 3, nw, threads=2, want_witness=True)
        for i in range(3):
            iw = {w: int.from_bytes(inp[(i * 8 + k) * 32:(i * 8 + k + 1) * 32], "big") for k, w in enumerate(inputs)}
            st, wm, _ = pwg.solve_circuit(c, iw)
            assert st == "Solved" and res[i, 0] == 0
            assert cref.witness_dict(ow, op, i) == wm
    # the "optimised CPU" variant computes the same witnesses
    data, inputs, nw = ab.synthetic_arith_circuit(500)
    c = acir.decode_circuit(data)
    inp = ab.synthetic_inputs(4)
    res, ow, op = cref.solve_batch(c, inputs, inp, 4, nw, threads=1, want_witness=True)
    res2, ow2 = cref.solve_batch_optimized(c, inputs, inp, 4, nw, threads=2, want_witness=True)
    assert (res2[:, 0] == 0).all() and (ow == ow2).all()
    # failure parity: unsatisfied check at opcode 1
    b = ab.CircuitBuilder()
    b.arithmetic([], [(1, 1), (ab.P - 1, 2)], 0)
    b.arithmetic([], [(1, 2)], ab.P - 7)
    c = acir.decode_circuit(b.to_bytes())
    res, _, _ = cref.solve_batch(c, [1], (8).to_bytes(32, "big") + (7).to_bytes(32, "big"), 2, 3, threads=1)
    assert list(res[0]) == [2, 4, 1, 0] and res[1, 0] == 0


def test_pedersen_kats_are_curve_points(golden):
    for k in golden["kats"]["pedersen"]:
        assert grumpkin.on_curve((int(k["x"], 16), int(k["y"], 16)))


@pytest.mark.xfail(strict=True, reason="Pedersen parity unpinned: the oracle restates the STRUCTURE over its own generators; "
                                       "barretenberg's derivation could not be reproduced (tools/pedersen_generator_search*.py)")
def test_pedersen_kats(golden):
    """The reference's two Pedersen KATs (wasm/pedersen.rs:38-54, acvm_js/test/shared/pedersen.ts:8-16).  Strict xfail: the day
    the oracle reproduces them this test XPASSes, fails the suite, and the opt-in gate (pedersen_unpinned) can be removed."""
    from oracle import pedersen
    for k in golden["kats"]["pedersen"]:
        assert pedersen.commit_native(k["inputs"], k["hash_index"]) == (int(k["x"], 16), int(k["y"], 16))


def test_pedersen_structure():
    from oracle import pedersen
    a = pedersen.commit_native([5, 6], 0)
    assert a != pedersen.commit_native([6, 5], 0) and a != pedersen.commit_native([5, 6], 1)
    assert pedersen.commit_native([], 3) == (0, 0)


def test_ecdsa_reference_kats(golden):  # blackbox_solver/src/lib.rs:216-290
    from oracle import ecdsa
    assert set(golden["kats"]["ecdsa_valid"]) == {"EcdsaSecp256k1", "EcdsaSecp256r1"}
    for name, k in golden["kats"]["ecdsa_valid"].items():
        args = [bytes.fromhex(k[a]) for a in ("hashed_message", "pub_key_x", "pub_key_y", "signature")]
        assert ecdsa.verify(name, *args) is True
        bad = bytearray(args[0]); bad[5] ^= 1
        assert ecdsa.verify(name, bytes(bad), *args[1:]) is False


def test_ecdsa_oracle_against_openssl():
    """Independent pin of the third-party arithmetic (k256 / p256 are not in the tree): OpenSSL, through `cryptography`,
    must agree with oracle/ecdsa.py on signatures made by either side.  OpenSSL accepts high-S signatures, so the
    comparison is on low-S ones; the low-S rule itself is the reference's (lib.rs:133-136)."""
    import random
    crypto = pytest.importorskip("cryptography")
    from cryptography.exceptions import InvalidSignature
    from cryptography.hazmat.primitives import hashes as chashes
    from cryptography.hazmat.primitives.asymmetric import ec as cec, utils as cutils
    from oracle import ecdsa
    rnd = random.Random(11)
    for name, curve in (("EcdsaSecp256k1", cec.SECP256K1()), ("EcdsaSecp256r1", cec.SECP256R1())):
        c = ecdsa.CURVES[name]
        for it in range(6):
            digest = bytes(rnd.randrange(256) for _ in range(32))
            if int.from_bytes(digest, "big") >= c.n:
                continue
            # signed by OpenSSL, verified by the oracle
            key = cec.generate_private_key(curve)
            pub = key.public_key().public_numbers()
            r, s = cutils.decode_dss_signature(key.sign(digest, cec.ECDSA(cutils.Prehashed(chashes.SHA256()))))
            sig = r.to_bytes(32, "big") + s.to_bytes(32, "big")
            px, py = pub.x.to_bytes(32, "big"), pub.y.to_bytes(32, "big")
            assert ecdsa.verify(name, digest, px, py, sig) is (s <= (c.n - 1) // 2)
            low = r.to_bytes(32, "big") + min(s, c.n - s).to_bytes(32, "big")
            assert ecdsa.verify(name, digest, px, py, low) is True
            # signed by the oracle's helper, verified by OpenSSL; and a corrupted copy rejected by both
            d = rnd.randrange(1, c.n)
            P = ecdsa.public_key(name, d)
            r2, s2 = ecdsa.sign(name, d, int.from_bytes(digest, "big"), rnd.randrange(1, c.n))
            opub = cec.EllipticCurvePublicNumbers(P[0], P[1], curve).public_key()
            opub.verify(cutils.encode_dss_signature(r2, s2), digest, cec.ECDSA(cutils.Prehashed(chashes.SHA256())))
            wrong = bytes([digest[0] ^ 0x40]) + digest[1:]
            with pytest.raises(InvalidSignature):
                opub.verify(cutils.encode_dss_signature(r2, s2), wrong, cec.ECDSA(cutils.Prehashed(chashes.SHA256())))
            if int.from_bytes(wrong, "big") < c.n:
                assert ecdsa.verify(name, wrong, P[0].to_bytes(32, "big"), P[1].to_bytes(32, "big"),
                                    r2.to_bytes(32, "big") + s2.to_bytes(32, "big")) is False


def test_permutation_route_reference_literals(golden):  # acvm/src/pwg/directives/sorting.rs:298-372
    import random
    from oracle import sorting
    for k in golden["kats"]["permutation_route"]:
        assert sorting.route(k["inputs"], k["outputs"]) == [bool(b) for b in k["bits"]]
    rnd = random.Random(3)
    for n in list(range(2, 50)) + [64, 100, 257]:     # the reference's own property test (sorting.rs:374-386)
        a = list(range(n))
        b = a[:]
        rnd.shuffle(b)
        c = sorting.route(a, b)
        assert len(c) == sorting.switch_count(n) and sorting.execute_network(c, a) == b
