"""Shared PermutationSort test material (Directive::PermutationSort, acvm/src/pwg/directives/mod.rs:88-121)."""
import random

from acvm_b200 import acir_builder as ab


def sort_circuit(n, tup, sort_by, n_bits=None, preassign_bit=False):
    """tuples over input witnesses 1..n*tup, PermutationSort, then a gate that uses one control bit"""
    b = ab.CircuitBuilder()
    from oracle import sorting
    nb = sorting.switch_count(n) if n_bits is None else n_bits
    bits = list(range(1000, 1000 + nb))
    if preassign_bit and nb:
        b.arithmetic([], [(1, bits[0])], ab.P - 1)            # bits[0] := 1 before the directive assigns it
    # element i = (w_{i*tup+1} + 3, 2 * w_{i*tup+2}, ...): expressions, not bare witnesses
    inputs = [[([], [(1 + k, 1 + i * tup + k)], 3 * (k == 0)) for k in range(tup)] for i in range(n)]
    b.directive_permutation_sort(inputs, tup, bits, sort_by)
    if nb:
        b.arithmetic([(1, bits[0], 1)], [(1, bits[-1]), (ab.P - 1, 2000)], 5)   # w2000 = bit0 * w1 + bit_last + 5
    return b.to_bytes(), list(range(1, n * tup + 1))


def sort_rows(n, tup, batch):
    """input rows: many equal keys (stable order decides), full-width field values, 16-bit values"""
    rnd = random.Random(n * 10 + tup)
    rows = []
    for i in range(batch):
        if i % 3 == 0:
            rows.append([rnd.randrange(3) for _ in range(n * tup)])
        elif i % 3 == 1:
            rows.append([rnd.randrange(ab.P) for _ in range(n * tup)])
        else:
            rows.append([rnd.randrange(1 << 16) for _ in range(n * tup)])
    return b"".join(int(v).to_bytes(32, "big") for r in rows for v in r)
