"""Shared ECDSA test material: a circuit around BlackBoxFuncCall::EcdsaSecp256{k1,r1} and input rows that cover every
branch of blackbox_solver/src/lib.rs:101-210 (valid, low-S rejection, wrong message, parity-only use of y, and each
panicking conversion)."""
import random

from acvm_b200 import acir_builder as ab
from oracle import ecdsa

PKX, PKY, SIG, HM = range(1, 33), range(33, 65), range(65, 129), range(129, 161)
OUT = 161
INPUTS = list(range(1, 161))


def circuit(name, n_hm=32, n_pkx=32, preassigned_out=False):
    """ECDSA opcode -> arithmetic on its result -> RecursiveAggregation zeros -> a gate reading one of the zeros."""
    b = ab.CircuitBuilder()
    fi = lambda ws: [(w, 8) for w in ws]
    b.ecdsa(name, fi(list(PKX)[:n_pkx]), fi(PKY), fi(SIG), fi(list(HM)[:n_hm]), OUT)
    b.arithmetic([], [(5, OUT), (ab.P - 1, 162)], 3)                  # w162 = 5 * valid + 3
    b.recursive_aggregation(fi([1, 2]), fi([3]), fi([4]), (5, 254), None, [162 if preassigned_out else 165, 163, 164])
    b.arithmetic([(1, 163, 164)], [(1, 162), (ab.P - 1, 166)], 0)     # w166 = w163 * w164 + w162
    return b.to_bytes()


def cases(name, seed=1, n_random=2):
    """-> list of (label, hashed_msg, pkx, pky, sig, expected) with expected in {True, False, "panic"}."""
    c = ecdsa.CURVES[name]
    rnd = random.Random(seed)
    b32 = lambda v: v.to_bytes(32, "big")
    out = []

    def add(label, z, P, r, s):
        args = (b32(z), b32(P[0]), b32(P[1]), b32(r) + b32(s))
        try:
            exp = ecdsa.verify(name, *args)
        except ecdsa.ReferencePanic:
            exp = "panic"
        out.append((label,) + args + (exp,))

    for _ in range(n_random):
        d, z, k = rnd.randrange(1, c.n), rnd.randrange(c.n), rnd.randrange(1, c.n)
        P = ecdsa.public_key(name, d)
        r, s = ecdsa.sign(name, d, z, k)
        add("valid", z, P, r, s)
        add("high_s", z, P, r, c.n - s)
        add("wrong_msg", (z + 1) % c.n, P, r, s)
        add("other_parity", z, (P[0], P[1] ^ 1), r, s)
        add("wrong_y_same_parity", z, (P[0], (P[1] + 2) % c.p), r, s)
        add("r_zero", z, P, 0, s)
        add("s_zero", z, P, r, 0)
        add("r_is_n", z, P, c.n, s)
        add("hash_is_n", c.n, P, r, s)
        add("x_is_p", z, (c.p, P[1]), r, s)
        x = rnd.randrange(c.p)
        while ecdsa.decompress(c, x, 0) is not None:
            x = rnd.randrange(c.p)
        add("x_not_on_curve", z, (x, P[1]), r, s)
    d = rnd.randrange(1, c.n)
    P = ecdsa.public_key(name, d)
    r, s = rnd.randrange(1, c.n), rnd.randrange(1, (c.n - 1) // 2)
    add("R_at_infinity", (-r * d) % c.n, P, r, s)
    for P, d in (((c.gx, c.gy), 1), ((c.gx, c.p - c.gy), c.n - 1)):      # G + P doubles / cancels in the Shamir table
        z, k = rnd.randrange(c.n), rnd.randrange(1, c.n)
        r, s = ecdsa.sign(name, d, z, k)
        add("P_is_pm_G", z, P, r, s)
    add("zero_hash", 0, ecdsa.public_key(name, 7), *ecdsa.sign(name, 7, 0, 11))
    return out


def input_row(case, rnd=None):
    """160 witnesses, one byte each in its LOW byte (signature/mod.rs:5-18); high bytes are noise when rnd is given."""
    _, hm, pkx, pky, sig, _ = case
    vals = list(pkx) + list(pky) + list(sig) + list(hm)
    if rnd is not None:
        vals = [v + 256 * rnd.randrange(1 << 200) if rnd.random() < 0.3 else v for v in vals]
    return b"".join(int(v).to_bytes(32, "big") for v in vals)
