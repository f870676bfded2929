"""Search for barretenberg v0.5.0's Pedersen generator derivation against the two reference KATs.

TEST/RESEARCH TOOL.  The KATs (the only acceptance test) are
  barretenberg_blackbox_solver/src/wasm/pedersen.rs:38-54   inputs [0, 1], hash_index 0
  acvm_js/test/shared/pedersen.ts:8-16                      inputs [1],    hash_index 0
The structure (merkle_damgard_compress / hash_pair / hash_single over 9-bit windows) is fixed;
what is searched is how `derive_generators` maps a counter to a curve point.
"""
import itertools
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import grumpkin
from oracle.field import P
from oracle.hashes import keccak256, sha256
from oracle.pedersen import _sqrt

KATS = [
    ([0, 1], 0, (0x0c5e1ddecd49de44ed5e5798d3f6fb7c71fe3d37f5bee8664cf88a445b5ba0af,
                 0x230294a041e26fe80b827c2ef5cb8784642bbaa83842da2714d62b1f3c4f9752)),
    ([1], 0, (0x09489945604c9686e698cb69d7bd6fc0cdb02e9faae3e1a433f1c342c1a5ecc4,
              0x24f50d25508b4dfb1e8a834e39565f646e217b24cb3a475c2e4991d1bb07a9d8)),
]
R = (1 << 256) % P


def point_from_x(x, ybit):
    y = _sqrt((x * x * x - 17) % P)
    if y is None:
        return None
    if (y & 1) != ybit:
        y = P - y
    return (x, y)


def gens_from_x(n, ybit_rule):
    """derive_from_x_coordinate(Fq(seed), sign)"""
    out, seed = [], 0
    while len(out) < n:
        seed += 1
        pt = point_from_x(seed, ybit_rule)
        if pt is not None:
            out.append(pt)
    return out


def gens_hash(n, enc, digest_order, clear_top, ybit_src, hashfn):
    out, seed = [], 0
    while len(out) < n:
        seed += 1
        v = seed if enc.startswith("raw") else (seed * R) % P
        if enc.endswith("limb"):
            msg = b"".join(((v >> (64 * j)) & (2**64 - 1)).to_bytes(8, "big") for j in range(4))
        elif enc.endswith("le"):
            msg = v.to_bytes(32, "little")
        else:
            msg = v.to_bytes(32, "big")
        h = hashfn(msg)
        if digest_order == "limb":
            h = b"".join(h[8 * j:8 * j + 8][::-1] for j in range(4))[::-1]
            digest_order_ = "big"
        else:
            digest_order_ = digest_order
        hv = int.from_bytes(h, digest_order_)
        ybit = (hv >> 255) & 1 if ybit_src == "top" else hv & 1
        x = hv & ((1 << 255) - 1) if clear_top else hv
        if x >= P:
            x_red = x % P
        else:
            x_red = x
        pt = point_from_x(x_red, ybit)
        if pt is not None:
            out.append(pt)
    return out


def commit(inputs, iv, gens, ivgen, nwin, layout):
    def table_gen(parity, i):
        if layout == "halves":
            return gens[parity * nwin + i]
        return gens[2 * i + parity]

    def hash_single(v, parity):
        acc = grumpkin.INF
        for i in range(nwin):
            s = (v >> (9 * i)) & 511
            acc = grumpkin.add(acc, grumpkin.mul(s + 1, table_gen(parity, i)))
        return acc

    def hash_pair(a, b):
        return grumpkin.add(hash_single(a, 0), hash_single(b, 1))[0]

    r = grumpkin.mul(iv + 1, ivgen)[0]
    for v in inputs:
        r = hash_pair(r, v % P)
    return grumpkin.add(hash_single(r, 0), hash_single(len(inputs), 1))


def cube_roots():
    # nontrivial cube roots of unity in Fq(grumpkin) = BN254 Fr
    g = 5
    b = pow(g, (P - 1) // 3, P)
    assert b != 1 and pow(b, 3, P) == 1
    return [b, b * b % P]


def commit_endo(inputs, iv, gens, ivgen, beta, split):
    """30-table variant: 15 tables per side, two 9-bit slices per round, second through the endomorphism."""
    def endo(pt):
        return None if pt is None else (pt[0] * beta % P, pt[1])

    def hash_single(v, parity):
        off = 15 * parity
        a0 = a1 = grumpkin.INF
        if split == "interleaved":
            sl = [(v >> (9 * k)) & 511 for k in range(30)]
            for i in range(15):
                a0 = grumpkin.add(a0, grumpkin.mul(sl[2 * i] + 1, gens[off + i]))
                if i < 14:
                    a1 = grumpkin.add(a1, grumpkin.mul(sl[2 * i + 1] + 1, gens[off + i]))
        elif split == "lohi":
            lo, hi = v & ((1 << 126) - 1), v >> 126
            for i in range(14):
                a0 = grumpkin.add(a0, grumpkin.mul(((lo >> (9 * i)) & 511) + 1, gens[off + i]))
            for i in range(15):
                a1 = grumpkin.add(a1, grumpkin.mul(((hi >> (9 * i)) & 511) + 1, gens[off + i]))
        else:  # hilo: a from hi (15 incl small), b from lo
            lo, hi = v & ((1 << 126) - 1), v >> 126
            for i in range(15):
                a0 = grumpkin.add(a0, grumpkin.mul(((hi >> (9 * i)) & 511) + 1, gens[off + i]))
            for i in range(14):
                a1 = grumpkin.add(a1, grumpkin.mul(((lo >> (9 * i)) & 511) + 1, gens[off + i]))
        return grumpkin.add(a0, endo(a1))

    def hash_pair(a, b):
        return grumpkin.add(hash_single(a, 0), hash_single(b, 1))[0]

    r = grumpkin.mul(iv + 1, ivgen)[0]
    for v in inputs:
        r = hash_pair(r, v % P)
    return grumpkin.add(hash_single(r, 0), hash_single(len(inputs), 1))


def main():
    cands = {}
    for ybit in (0, 1):
        cands[("fromx", ybit)] = gens_from_x(64, ybit)
    for enc, order, clear, ysrc, hname in itertools.product(
            ("raw", "mont", "rawlimb", "montlimb", "rawle", "montle"), ("little", "big", "limb"), (True, False), ("top", "low"), ("keccak", "sha256")):
        hf = keccak256 if hname == "keccak" else sha256
        cands[("hash", enc, order, clear, ysrc, hname)] = gens_hash(64, enc, order, clear, ysrc, hf)
    found = False
    for key, gens in cands.items():
        for nwin, layout in ((29, "halves"), (29, "interleaved")):
            for ivname, ivgen in (("g0", gens[0]), ("G", grumpkin.G), ("g58", gens[58])):
                ok = True
                for inputs, iv, want in KATS[1:] + KATS[:1]:
                    got = commit(inputs, iv, gens, ivgen, nwin, layout)
                    if got != want:
                        ok = False
                        break
                if ok:
                    print("MATCH", key, nwin, layout, ivname)
                    found = True
    betas = cube_roots()
    for key, gens in cands.items():
        for beta in betas:
            for split in ("interleaved", "lohi", "hilo"):
                for ivname, ivgen in (("g0", gens[0]), ("G", grumpkin.G), ("g30", gens[30])):
                    inputs, iv, want = KATS[1]
                    if commit_endo(inputs, iv, gens, ivgen, beta, split) == want:
                        print("MATCH-ENDO", key, beta, split, ivname)
                        found = True
    print("found" if found else "no match")


if __name__ == "__main__":
    main()
