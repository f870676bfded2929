"""Pedersen (plookup-structured) commitment for the oracle.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  The reference computes this inside barretenberg's wasm
(`pedersen_plookup_commit_with_hash_index`, barretenberg_blackbox_solver/src/wasm/pedersen.rs:14-35;
acvm_backend.wasm v0.5.0 per build.rs:10 is downloaded at build time and is NOT in the reference
tree, nor is any wasm runtime available here).  The generator points of barretenberg's tables come
from its own hash-to-curve and cannot be derived from anything under /root/reference, so the two
KATs the reference holds (pedersen.rs:38-54, acvm_js/test/shared/pedersen.ts:8-16) cannot be
reproduced.  tests/test_oracle_golden.py::test_pedersen_kats_unpinned records exactly that.

What is restated here is the STRUCTURE of barretenberg's lookup Pedersen:
  commit(inputs, iv)   = affine( hash_single(r, 0) + hash_single(len(inputs), 1) ),
                         r = iv_table[iv].x ; for v in inputs: r = hash_pair(r, v)
  hash_pair(a, b)      = ( hash_single(a, 0) + hash_single(b, 1) ).x
  hash_single(v, par)  = sum_i (slice_i(v) + 1) * G[par][i],  slice_i = 9-bit windows of the canonical value
with this project's OWN, documented generator derivation (keccak256 counter hash-to-curve below).
The GPU kernel implements exactly this function and is tested bit-for-bit against it; the cost
profile (58 table-lookup mixed additions + 1 inversion per chaining round) is that of the reference.
"""
from . import grumpkin
from .field import P
from .hashes import keccak256

BITS_PER_TABLE = 9
NUM_WINDOWS = 29            # 28 * 9 = 252 bits + one 2-bit window = 254
TABLE_SIZE = 1 << BITS_PER_TABLE
IV_TABLE_SIZE = 1024
DOMAIN = b"acvm_b200.pedersen.v1"


def _sqrt(a):
    """Tonelli-Shanks in Fr (p - 1 = 2^28 * odd)."""
    a %= P
    if a == 0:
        return 0
    if pow(a, (P - 1) // 2, P) != 1:
        return None
    s, q = 0, P - 1
    while q % 2 == 0:
        s += 1
        q //= 2
    z = 5
    while pow(z, (P - 1) // 2, P) != P - 1:
        z += 1
    m, c, t, r = s, pow(z, q, P), pow(a, q, P), pow(a, (q + 1) // 2, P)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % P
            i += 1
        b = pow(c, 1 << (m - i - 1), P)
        m, c = i, b * b % P
        t, r = t * c % P, r * b % P
    return r


def derive_generator(index: int):
    """index-th table generator: first counter whose keccak256 gives an x on the curve; y parity from the hash."""
    ctr = 0
    while True:
        h = keccak256(DOMAIN + index.to_bytes(4, "big") + ctr.to_bytes(4, "big"))
        x = int.from_bytes(h, "big") % P
        y = _sqrt((x * x * x - 17) % P)
        if y is not None and y != 0:
            if (y & 1) != (h[0] >> 7):
                y = P - y
            return (x, y)
        ctr += 1


_GENS = None


def generators():
    global _GENS
    if _GENS is None:
        _GENS = [[derive_generator(par * NUM_WINDOWS + i) for i in range(NUM_WINDOWS)] for par in range(2)]
    return _GENS


def hash_single(v: int, parity: int):
    acc = grumpkin.INF
    g = generators()[parity]
    for i in range(NUM_WINDOWS):
        s = (v >> (BITS_PER_TABLE * i)) & (TABLE_SIZE - 1)
        acc = grumpkin.add(acc, grumpkin.mul(s + 1, g[i]))
    return acc


def hash_pair(a: int, b: int) -> int:
    pt = grumpkin.add(hash_single(a, 0), hash_single(b, 1))
    return 0 if pt is grumpkin.INF else pt[0]


def iv_point_x(iv: int) -> int:
    return grumpkin.mul((iv % IV_TABLE_SIZE) + 1, grumpkin.G)[0]


def commit_native(inputs, hash_index: int):
    if len(inputs) == 0:
        return (0, 0)
    r = iv_point_x(hash_index)
    for v in inputs:
        r = hash_pair(r, v % P)
    pt = grumpkin.add(hash_single(r, 0), hash_single(len(inputs), 1))
    return (0, 0) if pt is grumpkin.INF else pt
