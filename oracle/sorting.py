"""Directive::PermutationSort (acvm/src/pwg/directives/mod.rs:88-121) and the switch routing of its permutation network
(acvm/src/pwg/directives/sorting.rs:5-235) -- TEST INFRASTRUCTURE ONLY.

`route(inputs, outputs)` returns the control bits of the recursive (Waksman-style) network that maps `inputs` to
`outputs`.  Many bit assignments realise the same permutation; parity needs the reference's choice, so the loop order
below follows sorting.rs:194-228 decision by decision: start from the last (single) output wire on the lower sub-network,
alternate output->input / input->output hops through sibling wires, restart from the lowest free output switch.
Bits are laid out [input switches (n/2)] [output switches ((n-1)/2)] [upper sub-network] [lower sub-network].

Pinned by the five literal vectors of sorting.rs:298-372 (tests/test_oracle_golden.py) and by executing the network.
"""
from __future__ import annotations


class ReferencePanic(Exception):
    pass


def _require(cond, what):
    if not cond:
        raise ReferencePanic(what)


def route(inputs, outputs):
    _require(len(inputs) == len(outputs), "assert_eq!(inputs.len(), outputs.len())")
    n = len(inputs)
    if n == 0:
        return []
    if n == 1:
        _require(inputs[0] == outputs[0], "assert_eq!(inputs[0], outputs[0])")
        return []
    if n == 2:
        if inputs[0] == outputs[0]:
            _require(inputs[1] == outputs[1], "assert_eq!(inputs[1], outputs[1])")
            return [False]
        _require(inputs[1] == outputs[0] and inputs[0] == outputs[1], "assert_eq! on the crossed pair")
        return [True]
    half = n // 2
    odd = n % 2 == 1
    x_of = {v: i for i, v in enumerate(inputs)}     # x_values: value -> input wire
    y_of = {v: i for i, v in enumerate(outputs)}    # y_values: value -> output wire
    switch_x = [False] * half
    switch_y = [False] * ((n - 1) // 2)
    inner_x = [0] * n
    inner_y = [0] * n
    free = set(range((n - 1) // 2))

    single_x = lambda a: odd and a == n - 1                    # sorting.rs:134-137
    single_y = lambda a: a >= n - 2 + n % 2                    # sorting.rs:139-142
    inner_pos = lambda idx, sw: idx // 2 + half if (sw ^ (idx % 2 == 1)) else idx // 2   # sorting.rs:145-151
    sibling = lambda i: i + 1 - 2 * (i % 2)

    # single wires are routed up front (sorting.rs:57-63)
    inner_y[n - 1] = outputs[n - 1]
    if not odd:
        inner_y[half - 1] = outputs[n - 2]
    else:
        inner_x[n - 1] = inputs[n - 1]

    def set_x(x, sw):
        inner_x[inner_pos(x, sw)] = inputs[x]
        switch_x[x // 2] = sw

    def set_y(y, sw):
        inner_y[inner_pos(y, sw)] = outputs[y]
        switch_y[y // 2] = sw

    def route_out_wire(y, sub):      # sorting.rs:66-86
        if single_y(y):
            _require(sub, "assert!(sub)")
        else:
            set_y(y, sub ^ (y % 2 != 0))
        _require(outputs[y] in x_of, "x_values.remove(..).unwrap()")
        x = x_of.pop(outputs[y])
        if not single_x(x):
            set_x(x, sub ^ (x % 2 != 0))
        return x

    def route_in_wire(x, sub):       # sorting.rs:89-106
        _require(not single_x(x), "assert!(!self.is_single_x(x))")
        set_x(x, sub ^ (x % 2 != 0))
        _require(inputs[x] in y_of, "y_values.remove(..).unwrap()")
        y = y_of.pop(inputs[x])
        if not single_y(y):
            set_y(y, sub ^ (y % 2 != 0))
        return y

    def new_start():                 # sorting.rs:153-160 (take() peeks at the smallest free switch)
        if free:
            s = min(free)
            return s, 2 * s
        return None, 0

    out_idx, start_sub, switch, start = n - 1, True, None, None
    while free:
        if switch is not None:
            free.discard(switch)
        in_idx = route_out_wire(out_idx, start_sub)
        if single_x(in_idx):
            start_sub = not start_sub
            start, out_idx = new_start()
            switch = start
            continue
        out_idx = route_in_wire(sibling(in_idx), not start_sub)
        switch = out_idx // 2
        if start == switch or single_y(out_idx):
            start, out_idx = new_start()
            switch = start
        else:
            out_idx = sibling(out_idx)
    bits = switch_x + switch_y
    bits += route(inner_x[:half], inner_y[:half])
    bits += route(inner_x[half:], inner_y[half:])
    return bits


def switch_count(n):
    """number of switches of the network on n wires (sorting.rs:289-295: sum of ceil(log2(i+1)))"""
    return sum((i).bit_length() for i in range(n))   # ceil(log2(i+1)) == bit_length(i)


def execute_network(config, inputs):
    """sorting.rs:245-287 (the reference's own test helper): apply the switches to `inputs`."""
    n = len(inputs)
    if n <= 1:
        return list(inputs)
    in1, in2 = [], []
    for i in range(n // 2):
        a, b = inputs[2 * i], inputs[2 * i + 1]
        if config[i]:
            a, b = b, a
        in1.append(a)
        in2.append(b)
    if n % 2 == 1:
        in2.append(inputs[-1])
    n2 = n // 2 + (n - 1) // 2
    n3 = n2 + switch_count(n // 2)
    out1 = execute_network(config[n2:n3], in1)
    out2 = execute_network(config[n3:], in2)
    res = []
    for i in range((n - 1) // 2):
        if config[n // 2 + i]:
            res += [out2[i], out1[i]]
        else:
            res += [out1[i], out2[i]]
    if n % 2 == 0:
        res += [out1[-1], out2[-1]]
    else:
        res.append(out2[-1])
    return res


def permutation_sort_bits(elements, sort_by):
    """directives/mod.rs:88-115.  elements: list of tuples of field values (canonical ints).  Returns the control bits."""
    rows = [list(e) + [i] for i, e in enumerate(elements)]
    for i in sort_by:
        for r in rows:
            _require(i < len(r), "index out of bounds in sort_by")

    def key(r):
        return tuple(r[i] for i in sort_by)
    rows.sort(key=key)               # Python's sort is stable, like slice::sort_by
    return route(list(range(len(elements))), [r[-1] for r in rows])
