"""ECDSA verification over secp256k1 / secp256r1 exactly as the reference calls it -- TEST INFRASTRUCTURE ONLY.

Restates `verify_secp256k1_ecdsa_signature` / `verify_secp256r1_ecdsa_signature`
(blackbox_solver/src/lib.rs:101-210), which sit on the third-party crates k256 0.11.6 / p256 0.11.1 (ecdsa 0.14.8,
elliptic-curve 0.12.3; Cargo.lock:828-829,846-847,1369-1370,1566-1567; not vendored under /root/reference).  The crates
implement SEC1 / FIPS 186-4 arithmetic, restated here with Python big integers.  What the *call sites* add, and what
this file must therefore reproduce bit for bit:

  * `Signature::try_from(sig).unwrap()`                 -> r, s must both be in [1, n-1], otherwise the process panics;
  * `EncodedPoint::from_affine_coordinates(x, y, true)` -> COMPRESSED encoding: only the parity of `y` survives;
    `PublicKey::from_encoded_point(..).unwrap()`        -> x >= p or x^3+ax+b a non-residue panics; the point used is
                                                           (x, sqrt(..) with the parity of the given y);
  * `Scalar::from_repr(hashed_msg).unwrap()`            -> the 32-byte hash as an integer must be < n, else panic
    (a slice that is not 32 bytes long panics inside `GenericArray::from_slice`);
  * `s.is_high()` -> false (BIP-0062 low-S rule: s > (n-1)/2 is rejected, no panic);
  * R = u1*G + u2*P; the point at infinity hits `unreachable!`, R.x >= n panics in `Scalar::from_repr(x).unwrap()`;
  * result = (R.x == r) with R.x NOT reduced mod n.

Panics are reported as `ReferencePanic` (the ACVM mirror turns them into the error kind of the same name).

Pins: the two known-answer tests of the reference (blackbox_solver/src/lib.rs:216-290, committed in
tests/golden/reference_vectors.json) and, independently, OpenSSL through the `cryptography` package on random
signatures (tests/test_oracle_golden.py).
"""
from __future__ import annotations


class ReferencePanic(Exception):
    pass


class Curve:
    def __init__(self, name, p, a, b, gx, gy, n):
        self.name, self.p, self.a, self.b, self.gx, self.gy, self.n = name, p, a, b, gx, gy, n


SECP256K1 = Curve(
    "secp256k1",
    p=2**256 - 2**32 - 977, a=0, b=7,
    gx=0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798,
    gy=0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8,
    n=0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141)

SECP256R1 = Curve(
    "secp256r1",
    p=2**256 - 2**224 + 2**192 + 2**96 - 1, a=-3 % (2**256 - 2**224 + 2**192 + 2**96 - 1),
    b=0x5AC635D8AA3A93E7B3EBBD55769886BC651D06B0CC53B0F63BCE3C3E27D2604B,
    gx=0x6B17D1F2E12C4247F8BCE6E563A440F277037D812DEB33A0F4A13945D898C296,
    gy=0x4FE342E2FE1A7F9B8EE7EB4A7C0F9E162BCE33576B315ECECBB6406837BF51F5,
    n=0xFFFFFFFF00000000FFFFFFFFFFFFFFFFBCE6FAADA7179E84F3B9CAC2FC632551)

CURVES = {"EcdsaSecp256k1": SECP256K1, "EcdsaSecp256r1": SECP256R1}

INF = None


def add(c: Curve, P, Q):
    if P is INF:
        return Q
    if Q is INF:
        return P
    (x1, y1), (x2, y2) = P, Q
    if x1 == x2:
        if (y1 + y2) % c.p == 0:
            return INF
        lam = (3 * x1 * x1 + c.a) * pow(2 * y1, -1, c.p) % c.p
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, c.p) % c.p
    x3 = (lam * lam - x1 - x2) % c.p
    return x3, (lam * (x1 - x3) - y1) % c.p


def mul(c: Curve, k: int, P):
    acc = INF
    for bit in bin(k)[2:] if k else "":
        acc = add(c, acc, acc)
        if bit == "1":
            acc = add(c, acc, P)
    return acc


def decompress(c: Curve, x: int, y_is_odd: int):
    """AffinePoint::decompress (k256 arithmetic/affine.rs, p256 via primeorder): None when x >= p or no square root."""
    if x >= c.p:
        return None
    alpha = (x * x * x + c.a * x + c.b) % c.p
    beta = pow(alpha, (c.p + 1) // 4, c.p)   # p = 3 mod 4 for both curves
    if beta * beta % c.p != alpha:
        return None
    if (beta & 1) != y_is_odd:
        beta = (-beta) % c.p
    return x, beta


def verify(curve_name: str, hashed_msg: bytes, pub_key_x: bytes, pub_key_y: bytes, signature: bytes) -> bool:
    """blackbox_solver/src/lib.rs:101-154 (k1) / :156-210 (r1).  Raises ReferencePanic where the reference panics."""
    c = CURVES[curve_name]
    assert len(pub_key_x) == 32 and len(pub_key_y) == 32 and len(signature) == 64   # fixed-size arrays in the reference
    r = int.from_bytes(signature[:32], "big")
    s = int.from_bytes(signature[32:], "big")
    if not (0 < r < c.n and 0 < s < c.n):
        raise ReferencePanic("Signature::try_from(..).unwrap()")
    P = decompress(c, int.from_bytes(pub_key_x, "big"), pub_key_y[31] & 1)
    if P is None:
        raise ReferencePanic("PublicKey::from_encoded_point(..).unwrap()")
    if len(hashed_msg) != 32:
        raise ReferencePanic("GenericArray::from_slice length mismatch")
    z = int.from_bytes(hashed_msg, "big")
    if z >= c.n:
        raise ReferencePanic("Scalar::from_repr(hashed_msg).unwrap()")
    if s > (c.n - 1) // 2:
        return False
    s_inv = pow(s, -1, c.n)
    u1, u2 = z * s_inv % c.n, r * s_inv % c.n
    R = add(c, mul(c, u1, (c.gx, c.gy)), mul(c, u2, P))
    if R is INF:
        raise ReferencePanic("unreachable!(\"Point is uncompressed\")")
    if R[0] >= c.n:
        raise ReferencePanic("Scalar::from_repr(x).unwrap()")
    return R[0] == r


def sign(curve_name: str, d: int, z: int, k: int, low_s: bool = True):
    """Test helper (no counterpart in the reference): textbook ECDSA signature with nonce k."""
    c = CURVES[curve_name]
    R = mul(c, k, (c.gx, c.gy))
    r = R[0] % c.n
    s = pow(k, -1, c.n) * (z + r * d) % c.n
    if low_s and s > (c.n - 1) // 2:
        s = c.n - s
    return r, s


def public_key(curve_name: str, d: int):
    c = CURVES[curve_name]
    return mul(c, d, (c.gx, c.gy))
