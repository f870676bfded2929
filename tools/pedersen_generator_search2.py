"""Second, wider sweep of the generator-derivation search (see pedersen_generator_search.py).

Adds the 'Montgomery leak' variants an old barretenberg could have had (raw limbs used as a field
element without conversion, parity taken on the Montgomery representation), a y > p/2 sign rule,
seeds starting at 0, and runs every variant against KAT 2 ([1], index 0) under both table structures.
"""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import grumpkin
from oracle.field import P
from oracle.hashes import keccak256, sha256, blake2s
from oracle.pedersen import _sqrt
from tools.pedersen_generator_search import KATS, commit, commit_endo, cube_roots

R = (1 << 256) % P
RINV = pow(R, -1, P)
M64 = (1 << 64) - 1


def encode(v, enc):
    if enc == "be32":
        return v.to_bytes(32, "big")
    if enc == "le32":
        return v.to_bytes(32, "little")
    if enc == "limb":   # limb 0 first, each limb big-endian
        return b"".join(((v >> (64 * j)) & M64).to_bytes(8, "big") for j in range(4))
    if enc == "be8":
        return (v & M64).to_bytes(8, "big")
    if enc == "le8":
        return (v & M64).to_bytes(8, "little")
    raise ValueError(enc)


def digest_int(h, order):
    if order == "little":
        return int.from_bytes(h, "little")
    if order == "big":
        return int.from_bytes(h, "big")
    # word64s loaded big-endian, word 0 least significant
    return sum(int.from_bytes(h[8 * j:8 * j + 8], "big") << (64 * j) for j in range(4))


def gens(n, seedform, enc, hname, order, clear, xmode, ysrc, yrule, seed0):
    hf = {"keccak": keccak256, "sha256": sha256, "blake2s": blake2s}[hname]
    out, seed = [], seed0 - 1
    while len(out) < n:
        seed += 1
        v = seed if seedform == "raw" else seed * R % P
        hv = digest_int(hf(encode(v, enc)), order)
        ybit = (hv >> 255) & 1 if ysrc == "top" else hv & 1
        x = hv & ((1 << 255) - 1) if clear else hv
        x = x % P if xmode == "plain" else x * RINV % P
        y = _sqrt((x * x * x - 17) % P)
        if y is None or (x == 0 and y == 0):
            continue
        if yrule == "lsb":
            cur = y & 1
        elif yrule == "lsbmont":
            cur = (y * R % P) & 1
        else:
            cur = 1 if y > (P - 1) // 2 else 0
        if cur != ybit:
            y = P - y
        out.append((x, y))
    return out


def main():
    betas = cube_roots()
    want = KATS[1][2]
    n = 0
    space = itertools.product(("raw", "mont"), ("be32", "le32", "limb", "be8", "le8"), ("keccak", "sha256", "blake2s"),
                              ("little", "big", "wordbe"), (True, False), ("plain", "rinv"), ("top", "low"),
                              ("lsb", "lsbmont", "half"), (1, 0))
    for key in space:
        g = gens(59, *key)
        n += 1
        for layout in ("halves", "interleaved"):
            for ivname, ivgen in (("g0", g[0]), ("G", grumpkin.G), ("g58", g[58])):
                if commit([1], 0, g, ivgen, 29, layout) == want:
                    print("MATCH", key, layout, ivname, flush=True)
        for beta in betas:
            for split in ("interleaved", "lohi", "hilo"):
                for ivname, ivgen in (("g0", g[0]), ("G", grumpkin.G), ("g30", g[30])):
                    if commit_endo([1], 0, g, ivgen, beta, split) == want:
                        print("MATCH-ENDO", key, beta, split, ivname, flush=True)
        if n % 200 == 0:
            print("tried", n, flush=True)
    print("done", n)


if __name__ == "__main__":
    shard, nshard = int(sys.argv[1]), int(sys.argv[2])
    _orig = itertools.product

    def sharded(*a):
        for i, k in enumerate(_orig(*a)):
            if i % nshard == shard:
                yield k
    itertools.product = sharded
    main()
