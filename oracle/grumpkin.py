"""Grumpkin (y^2 = x^3 - 17 over BN254 Fr) for the oracle.  TEST INFRASTRUCTURE ONLY.

The reference reaches this through barretenberg's wasm export `compute_public_key`
(barretenberg_blackbox_solver/src/wasm/scalar_mul.rs:17-65; barretenberg acvm_backend.wasm
v0.5.0 per build.rs:10 -- not in the reference tree).  The operation is plain scalar * G with
G = (1, sqrt(-16)) as pinned by the KATs in scalar_mul.rs:72-97, so it is restated with textbook
affine arithmetic.
"""
from .field import P

ORDER = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47  # scalar_mul.rs:42-45
B = (-17) % P
G = (1, 0x0000000000000002CF135E7506A45D632D270D45F1181294833FC48D823F272C)
INF = None


def on_curve(pt):
    if pt is INF:
        return True
    x, y = pt
    return (y * y - x * x * x - B) % P == 0


def add(p1, p2):
    if p1 is INF:
        return p2
    if p2 is INF:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return INF
        lam = 3 * x1 * x1 * pow(2 * y1, P - 2, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, P - 2, P) % P
    x3 = (lam * lam - x1 - x2) % P
    return (x3, (lam * (x1 - x3) - y1) % P)


def neg(p):
    return INF if p is INF else (p[0], (-p[1]) % P)


def mul(k, pt):
    acc = INF
    while k:
        if k & 1:
            acc = add(acc, pt)
        pt = add(pt, pt)
        k >>= 1
    return acc


class BlackBoxFailed(Exception):
    def __init__(self, func, reason):
        super().__init__(f"{func}: {reason}")
        self.func, self.reason = func, reason


def fixed_base_scalar_mul(low, high):
    """scalar_mul.rs:17-65: limb range checks, scalar < group order, then scalar*G affine."""
    if low.bit_length() > 128:
        raise BlackBoxFailed("FixedBaseScalarMul", f"Limb {low:064x} is not less than 2^128")
    if high.bit_length() > 128:
        raise BlackBoxFailed("FixedBaseScalarMul", f"Limb {high:064x} is not less than 2^128")
    s = (high << 128) | low
    if s >= ORDER:
        raise BlackBoxFailed("FixedBaseScalarMul", f"{s:x} is not a valid grumpkin scalar")
    pt = mul(s, G)
    if pt is INF:
        # s == 0: barretenberg's encoding of the point at infinity is NOT pinned by any
        # reference test.  The oracle and the kernels agree on (0, 0).
        return (0, 0)
    return pt
