"""CPU oracle for the batched ACIR witness-solve hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``acvm_b200/`` imports this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may use it, and only as the checker.

It is an independent restatement (Python big ints; ``ref_solver.cpp`` for timing) of the
reference algorithm noir-lang/acvm 0.27.0 implements in Rust.  The reference cannot be compiled
in the build container (no cargo/rustc, no wasm runtime, no acvm_backend.wasm), so the oracle
is pinned against every golden vector the reference's own tests hold for this path
(``tests/golden/reference_vectors.json``, extracted by ``tests/golden/make_golden.py``).

Parity status:
  * field, ACIR wire format, arithmetic solver, AND/XOR/RANGE, SHA-256, fixed-base scalar
    mul: pinned by reference KATs / golden byte vectors.
  * Keccak-256: the reference holds no literal KAT; pinned against hashlib.sha3_256 with the
    padding byte switched (same permutation) and the well-known keccak256("") value.
  * Pedersen: see oracle/pedersen.py header (algorithm lives in barretenberg's wasm, which is
    not in the reference tree; only 2 KATs exist).
"""
