"""Third sweep: the most likely generator derivation (keccak hash-to-curve of the raw seed, limb-wise big-endian, little-endian
digest, top bit = y parity) and its close variants, under MANY structural variants of the lookup Pedersen: IV generator and
multiple, parity order, final block, window order.  KAT 2 ([1], index 0) is the acceptance test."""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import grumpkin
from oracle.field import P
from tools.pedersen_generator_search import KATS, cube_roots
from tools.pedersen_generator_search2 import gens

want = KATS[1][2]
want0 = KATS[0][2]


def run(g, tag):
    G = grumpkin.G
    betas = cube_roots()
    tables = {}

    def tmul(k, s):
        key = (k, s)
        if key not in tables:
            tables[key] = grumpkin.mul(s, g[k])
        return tables[key]

    def hs(v, parity, nwin, offs, wsize, rev, endo):
        # endo: None | (beta, split) for the 15-table endomorphism variants
        acc = grumpkin.INF
        if endo is None:
            for i in range(nwin):
                s = (v >> (wsize * i)) & ((1 << wsize) - 1)
                k = offs[parity] + (nwin - 1 - i if rev else i)
                acc = grumpkin.add(acc, tmul(k, s + 1))
            return acc
        beta, split = endo
        a0 = a1 = grumpkin.INF
        off = offs[parity]
        if split == "interleaved":
            for i in range(15):
                a0 = grumpkin.add(a0, tmul(off + i, ((v >> (18 * i)) & 511) + 1))
                if i < 14:
                    a1 = grumpkin.add(a1, tmul(off + i, ((v >> (18 * i + 9)) & 511) + 1))
        else:
            lo, hi = v & ((1 << 126) - 1), v >> 126
            first, second = (lo, hi) if split == "lohi" else (hi, lo)
            n0, n1 = (14, 15) if split == "lohi" else (15, 14)
            for i in range(n0):
                a0 = grumpkin.add(a0, tmul(off + i, ((first >> (9 * i)) & 511) + 1))
            for i in range(n1):
                a1 = grumpkin.add(a1, tmul(off + i, ((second >> (9 * i)) & 511) + 1))
        e1 = None if a1 is None else (a1[0] * beta % P, a1[1])
        return grumpkin.add(a0, e1)

    structs = [("A58", 29, (0, 29), 9, False, None), ("A58rev", 29, (0, 29), 9, True, None),
               ("A58swap", 29, (29, 0), 9, False, None), ("A58alt", 29, None, 9, False, None)]
    for beta in betas:
        for split in ("interleaved", "lohi", "hilo"):
            structs.append((f"B30-{split}-{betas.index(beta)}", 15, (0, 15), 9, False, (beta, split)))
            structs.append((f"B30swap-{split}-{betas.index(beta)}", 15, (15, 0), 9, False, (beta, split)))
    ivs = [("G", G)] + [(f"g{k}", g[k]) for k in range(len(g))]
    for name, nwin, offs, wsize, rev, endo in structs:
        if offs is None:
            continue
        h1 = hs(1, 1, nwin, offs, wsize, rev, endo)          # H1(input = 1) and H1(len = 1) coincide for KAT 2
        for ivname, ivg in ivs:
            for ivmul in (1, 0, 2):
                r0 = 0 if ivmul == 0 else grumpkin.mul(ivmul, ivg)[0]
                for pair_order in (0, 1):
                    # hash_pair(r0, 1)
                    if pair_order == 0:
                        pt = grumpkin.add(hs(r0, 0, nwin, offs, wsize, rev, endo), h1)
                    else:
                        pt = grumpkin.add(hs(1, 0, nwin, offs, wsize, rev, endo), hs(r0, 1, nwin, offs, wsize, rev, endo))
                    if pt is None:
                        continue
                    r1 = pt[0]
                    for final in ("len", "nolen", "x_only"):
                        if final == "len":
                            res = grumpkin.add(hs(r1, 0, nwin, offs, wsize, rev, endo), h1)
                        elif final == "nolen":
                            res = hs(r1, 0, nwin, offs, wsize, rev, endo)
                        else:
                            res = pt
                        if res == want:
                            print("MATCH", tag, name, ivname, ivmul, pair_order, final, flush=True)
                            return True
    return False


def main():
    fams = []
    for seedform, enc, order, clear, xmode, ysrc, yrule, seed0 in itertools.product(
            ("raw", "mont"), ("limb", "be32", "le32"), ("little", "big"), (True, False), ("plain",), ("top", "low"), ("lsb", "half"), (1, 0)):
        fams.append((seedform, enc, "keccak", order, clear, xmode, ysrc, yrule, seed0))
    shard, nshard = int(sys.argv[1]), int(sys.argv[2])
    for i, key in enumerate(fams):
        if i % nshard != shard:
            continue
        g = gens(64, *key)
        if run(g, key):
            return
        print("done", i, key, flush=True)


if __name__ == "__main__":
    main()
