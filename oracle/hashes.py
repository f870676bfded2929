"""SHA-256 and Keccak-256 for the oracle.  TEST INFRASTRUCTURE ONLY.

Reference: blackbox_solver/src/lib.rs:47-60,86-91 call RustCrypto sha2 0.10.7 / sha3 0.10.8
(Cargo.lock; crates.io, not vendored).  SHA-256 = FIPS 180-4 (hashlib).  Keccak-256 = original
Keccak padding 0x01..0x80, rate 136 (NOT SHA3's 0x06) -- restated here from the published
permutation; round constants / rotations agree with the in-tree circuit spec
stdlib/src/blackbox_fallbacks/keccak256.rs:14-46.
"""
import hashlib

RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B,
    0x0000000080000001, 0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088,
    0x0000000080008009, 0x000000008000000A, 0x000000008000808B, 0x800000000000008B, 0x8000000000008089,
    0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
    0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008,
]
ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
M64 = (1 << 64) - 1


def _rol(x, n):
    n %= 64
    return ((x << n) | (x >> (64 - n))) & M64 if n else x


def keccak_f(A):
    for rnd in range(24):
        C = [A[x][0] ^ A[x][1] ^ A[x][2] ^ A[x][3] ^ A[x][4] for x in range(5)]
        D = [C[(x - 1) % 5] ^ _rol(C[(x + 1) % 5], 1) for x in range(5)]
        A = [[A[x][y] ^ D[x] for y in range(5)] for x in range(5)]
        B = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                B[y][(2 * x + 3 * y) % 5] = _rol(A[x][y], ROT[x][y])
        A = [[B[x][y] ^ ((~B[(x + 1) % 5][y]) & B[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        A[0][0] ^= RC[rnd]
    return A


def _sponge(msg: bytes, pad_byte: int) -> bytes:
    rate = 136
    m = bytearray(msg)
    m.append(pad_byte)
    while len(m) % rate:
        m.append(0)
    m[-1] |= 0x80
    A = [[0] * 5 for _ in range(5)]
    for off in range(0, len(m), rate):
        blk = m[off:off + rate]
        for i in range(rate // 8):
            A[i % 5][i // 5] ^= int.from_bytes(blk[8 * i:8 * i + 8], "little")
        A = keccak_f(A)
    out = b"".join(A[i % 5][i // 5].to_bytes(8, "little") for i in range(4))
    return out


def keccak256(msg: bytes) -> bytes:
    return _sponge(msg, 0x01)


def _sha3_256_via_own_permutation(msg: bytes) -> bytes:
    """Same sponge with SHA-3 domain byte; used by tests to pin the permutation against hashlib."""
    return _sponge(msg, 0x06)


def sha256(msg: bytes) -> bytes:
    return hashlib.sha256(bytes(msg)).digest()


def blake2s(msg: bytes) -> bytes:
    return hashlib.blake2s(bytes(msg)).digest()
