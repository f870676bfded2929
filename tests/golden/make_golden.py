#!/usr/bin/env python3
"""Extract golden vectors from the reference's own tests into tests/golden/*.json.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py
The JSON files it writes are committed; nothing at test time reads /root/reference.

Sources (all under /root/reference):
  acir/tests/test_program_serialization.rs      7 gzip(bincode) circuit byte vectors
  acvm_js/test/shared/*.ts                       same bytes + initial / expected witness maps
  acir_field/src/generic_ark.rs:424-438          hex of 0,-1,-2,-3
  barretenberg_blackbox_solver/src/wasm/*.rs     Pedersen / fixed-base KATs
  brillig_vm/src/black_box.rs:203-209            SHA-256("hello world")
  blackbox_solver/src/lib.rs:216-290             ECDSA secp256k1 / secp256r1 valid-signature KATs
  acvm/src/pwg/directives/sorting.rs:298-372     permutation-network routing literals
"""
import json, os, re, sys

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def rust_vectors(path):
    src = open(path).read()
    out = {}
    for m in re.finditer(r"fn (\w+)\(\) \{(.*?)\n\}\n", src, re.S):
        name, body = m.group(1), m.group(2)
        v = re.search(r"expected_serialization: Vec<u8> = vec!\[(.*?)\];", body, re.S)
        if v:
            out[name] = [int(x) for x in re.findall(r"\d+", v.group(1))]
    return out


def ts_fixture(path):
    src = open(path).read()
    fx = {}
    for m in re.finditer(r"export const (\w+)\s*(?::\s*\w+)?\s*=\s*Uint8Array\.from\(\[(.*?)\]\)", src, re.S):
        fx[m.group(1)] = [int(x) for x in re.findall(r"\d+", m.group(2))]
    for m in re.finditer(r"export const (\w+)\s*(?::\s*\w+)?\s*=\s*new Map(?:<[^>]*>)?\(\[(.*?)\]\);", src, re.S):
        pairs = re.findall(r"\[\s*(\d+)\s*,\s*\"(0x[0-9a-fA-F]+)\"\s*,?\s*\]", m.group(2))
        fx[m.group(1)] = {int(k): v for k, v in pairs}
    for m in re.finditer(r"export const (\w+)\s*=\s*(\d+);", src):
        fx[m.group(1)] = int(m.group(2))
    for m in re.finditer(r"export const (\w+)\s*=\s*\n?\s*\"(0x[0-9a-fA-F]+)\";", src):
        fx[m.group(1)] = m.group(2)
    return fx


def ecdsa_kats(path):
    """blackbox_solver/src/lib.rs:216-290: the two `verifies_valid_*_signature_with_low_s_value` tests."""
    src = open(path).read()
    out = {}
    for m in re.finditer(r"fn verifies_valid_(k1|r1)_signature_with_low_s_value\(\) \{(.*?)assert!\(valid\)", src, re.S):
        kat = {}
        for a in re.finditer(r"let (\w+)(?::\s*\[u8; \d+\])?\s*=\s*\[(.*?)\];", m.group(2), re.S):
            kat[a.group(1)] = bytes(int(x, 0) for x in re.findall(r"0x[0-9a-fA-F]+|\d+", a.group(2))).hex()
        assert {len(bytes.fromhex(v)) for v in kat.values()} == {32, 64}, kat
        out["EcdsaSecp256" + m.group(1)] = kat
    assert set(out) == {"EcdsaSecp256k1", "EcdsaSecp256r1"}
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit("reference not present; golden files are already committed")
    golden = {"source": "noir-lang/acvm 0.27.0 own tests (see make_golden.py docstring)"}
    golden["rust_serialization"] = rust_vectors(f"{REF}/acir/tests/test_program_serialization.rs")
    shared = {}
    for f in sorted(os.listdir(f"{REF}/acvm_js/test/shared")):
        if f.endswith(".ts"):
            shared[f[:-3]] = ts_fixture(f"{REF}/acvm_js/test/shared/{f}")
    golden["acvm_js_shared"] = shared
    # literal KATs; each line cites where the literal lives in the reference
    golden["kats"] = {
        "fr_hex": {  # acir_field/src/generic_ark.rs:424-438
            "0": "0000000000000000000000000000000000000000000000000000000000000000",
            "-1": "30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000000",
            "-2": "30644e72e131a029b85045b68181585d2833e84879b9709143e1f593efffffff",
            "-3": "30644e72e131a029b85045b68181585d2833e84879b9709143e1f593effffffe",
        },
        "inv5": "135b52945a13d9aa49b9b57c33cd568ba9ae5ce9ca4a2d06e7f3fbd4c6666667",  # acvm_js/test/shared/foreign_call.ts:20-27
        "sha256_hello_world": "b94d27b9934d3e08a52e52d7da7dabfac484efe37a5380ee9088f7ace2efcde9",  # brillig_vm/src/black_box.rs:203-209
        "fixed_base": [  # barretenberg_blackbox_solver/src/wasm/scalar_mul.rs:72-97
            {"low": 1, "high": 0,
             "x": "0000000000000000000000000000000000000000000000000000000000000001",
             "y": "0000000000000002cf135e7506a45d632d270d45f1181294833fc48d823f272c"},
            {"low": 1, "high": 2,
             "x": "0702ab9c7038eeecc179b4f209991bcb68c7cb05bf4c532d804ccac36199c9a9",
             "y": "23f10e9e43a3ae8d75d24154e796aae12ae7af546716e8f81a2564f1b5814130"},
        ],
        "pedersen": [  # barretenberg_blackbox_solver/src/wasm/pedersen.rs:38-54 ; acvm_js/test/shared/pedersen.ts:8-16
            {"inputs": [0, 1], "hash_index": 0,
             "x": "0c5e1ddecd49de44ed5e5798d3f6fb7c71fe3d37f5bee8664cf88a445b5ba0af",
             "y": "230294a041e26fe80b827c2ef5cb8784642bbaa83842da2714d62b1f3c4f9752"},
            {"inputs": [1], "hash_index": 0,
             "x": "09489945604c9686e698cb69d7bd6fc0cdb02e9faae3e1a433f1c342c1a5ecc4",
             "y": "24f50d25508b4dfb1e8a834e39565f646e217b24cb3a475c2e4991d1bb07a9d8"},
        ],
        "ecdsa_valid": ecdsa_kats(f"{REF}/blackbox_solver/src/lib.rs"),  # expected result: true
        "permutation_route": [  # acvm/src/pwg/directives/sorting.rs:298-372 (test_route literals): inputs, outputs, control bits
            {"inputs": [1, 2, 3], "outputs": [1, 2, 3], "bits": [0, 0, 0]},
            {"inputs": [1, 2, 3], "outputs": [1, 3, 2], "bits": [0, 0, 1]},
            {"inputs": [1, 2, 3], "outputs": [3, 2, 1], "bits": [1, 1, 1]},
            {"inputs": [0, 1, 2, 3], "outputs": [2, 3, 0, 1], "bits": [0, 1, 1, 1, 1]},
            {"inputs": [0, 1, 2, 3, 4], "outputs": [0, 3, 4, 2, 1], "bits": [0, 0, 0, 1, 0, 1, 0, 1]},
        ],
        "grumpkin_order": "30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47",  # scalar_mul.rs:42-45
    }
    # sanity: the rust and ts byte vectors must agree where both exist
    pairs = {"addition_circuit": "addition", "fixed_base_scalar_mul_circuit": "fixed_base_scalar_mul",
             "pedersen_circuit": "pedersen", "schnorr_verify_circuit": "schnorr_verify",
             "simple_brillig_foreign_call": "foreign_call", "complex_brillig_foreign_call": "complex_foreign_call",
             "memory_op_circuit": "memory_op"}
    for r, t in pairs.items():
        assert golden["rust_serialization"][r] == shared[t]["bytecode"], (r, t)
    with open(f"{OUT}/reference_vectors.json", "w") as f:
        json.dump(golden, f, indent=0, sort_keys=True)
    print("wrote", f"{OUT}/reference_vectors.json", {k: len(v) for k, v in golden["rust_serialization"].items()})


if __name__ == "__main__":
    main()
