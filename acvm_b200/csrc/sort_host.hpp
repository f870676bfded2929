// Directive::PermutationSort on the host: stable sort of the tuples + switch routing of the permutation network.
//
// Replaces (reference): acvm/src/pwg/directives/mod.rs:88-121 and acvm/src/pwg/directives/sorting.rs:5-235.
// The routing is sequential, pointer-chasing and recursive per instance (data-dependent walk over the switches), so it
// runs on the host threads like Brillig does: the plan cuts a host segment, the tuple columns are gathered out of HBM
// and the control bits scattered back (runtime.cu).  Many switch settings realise one permutation; the walk below
// makes the reference's choices in the reference's order (restart from the lowest free output switch, alternate
// sub-networks through sibling wires), which tests/test_host_logic.py checks against oracle/sorting.py.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

#include "fr_host.hpp"

namespace acvmb {
namespace psort {

// Values routed are distinct indices < width (the top-level inputs are 0..n-1 and every sub-network sees a subset).
// Returns false where the reference would panic (an output that is not among the inputs).
inline bool route(const std::vector<uint32_t>& in, const std::vector<uint32_t>& out, uint32_t width, std::vector<uint8_t>& bits) {
    const size_t n = in.size();
    if (out.size() != n) return false;
    if (n == 0) return true;
    if (n == 1) return in[0] == out[0];
    if (n == 2) {
        if (in[0] == out[0]) {
            if (in[1] != out[1]) return false;
            bits.push_back(0);
            return true;
        }
        if (in[1] != out[0] || in[0] != out[1]) return false;
        bits.push_back(1);
        return true;
    }
    const size_t half = n / 2, n_sy = (n - 1) / 2;
    const bool odd = n & 1;
    constexpr uint32_t GONE = 0xFFFFFFFFu;
    std::vector<uint32_t> x_of(width, GONE), y_of(width, GONE);
    for (size_t i = 0; i < n; ++i) {
        if (in[i] >= width || out[i] >= width) return false;
        x_of[in[i]] = (uint32_t)i;
        y_of[out[i]] = (uint32_t)i;
    }
    std::vector<uint8_t> sw_x(half, 0), sw_y(n_sy, 0), is_free(n_sy, 1);
    std::vector<uint32_t> inner_x(n, 0), inner_y(n, 0);
    size_t n_free = n_sy, lowest_free = 0;
    auto single_x = [&](size_t a) { return odd && a == n - 1; };
    auto single_y = [&](size_t a) { return a >= n - 2 + (n & 1); };
    auto inner_pos = [&](size_t idx, bool sw) { return (sw != ((idx & 1) == 1)) ? idx / 2 + half : idx / 2; };
    auto sibling = [](size_t i) { return i + 1 - 2 * (i & 1); };
    inner_y[n - 1] = out[n - 1];
    if (!odd) inner_y[half - 1] = out[n - 2];
    else inner_x[n - 1] = in[n - 1];
    auto set_x = [&](size_t x, bool sw) { inner_x[inner_pos(x, sw)] = in[x]; sw_x[x / 2] = sw; };
    auto set_y = [&](size_t y, bool sw) { inner_y[inner_pos(y, sw)] = out[y]; sw_y[y / 2] = sw; };
    bool ok = true;
    auto route_out_wire = [&](size_t y, bool sub) -> size_t {
        if (single_y(y)) { if (!sub) ok = false; }
        else set_y(y, sub != ((y & 1) != 0));
        uint32_t x = x_of[out[y]];
        if (x == GONE) { ok = false; return 0; }
        x_of[out[y]] = GONE;
        if (!single_x(x)) set_x(x, sub != ((x & 1) != 0));
        return x;
    };
    auto route_in_wire = [&](size_t x, bool sub) -> size_t {
        if (single_x(x)) { ok = false; return 0; }
        set_x(x, sub != ((x & 1) != 0));
        uint32_t y = y_of[in[x]];
        if (y == GONE) { ok = false; return 0; }
        y_of[in[x]] = GONE;
        if (!single_y(y)) set_y(y, sub != ((y & 1) != 0));
        return y;
    };
    constexpr size_t NO = (size_t)-1;
    auto new_start = [&](size_t& start, size_t& out_idx) {   // peek at the smallest free switch
        while (lowest_free < n_sy && !is_free[lowest_free]) ++lowest_free;
        if (n_free) { start = lowest_free; out_idx = 2 * lowest_free; }
        else { start = NO; out_idx = 0; }
    };
    size_t out_idx = n - 1, sw = NO, start = NO;
    bool start_sub = true;
    while (n_free) {
        if (sw != NO && sw < n_sy && is_free[sw]) { is_free[sw] = 0; --n_free; }
        size_t in_idx = route_out_wire(out_idx, start_sub);
        if (!ok) return false;
        if (single_x(in_idx)) {
            start_sub = !start_sub;
            new_start(start, out_idx);
            sw = start;
            continue;
        }
        out_idx = route_in_wire(sibling(in_idx), !start_sub);
        if (!ok) return false;
        sw = out_idx / 2;
        if (start == sw || single_y(out_idx)) {
            new_start(start, out_idx);
            sw = start;
        } else {
            out_idx = sibling(out_idx);
        }
    }
    bits.insert(bits.end(), sw_x.begin(), sw_x.end());
    bits.insert(bits.end(), sw_y.begin(), sw_y.end());
    std::vector<uint32_t> a(inner_x.begin(), inner_x.begin() + half), b(inner_y.begin(), inner_y.begin() + half);
    if (!route(a, b, width, bits)) return false;
    a.assign(inner_x.begin() + half, inner_x.end());
    b.assign(inner_y.begin() + half, inner_y.end());
    return route(a, b, width, bits);
}

// elements: n rows of `tuple` canonical field values.  sort_by indexes the row extended by its own position (mod.rs:97-99).
// Returns false where the reference panics (sort_by out of range).
inline bool permutation_sort_bits(const std::vector<U256>& values, uint32_t n, uint32_t tuple, const std::vector<uint32_t>& sort_by,
                                  std::vector<uint8_t>& bits) {
    for (uint32_t k : sort_by)
        if (k > tuple && n) return false;
    std::vector<uint32_t> order(n), base(n);
    for (uint32_t i = 0; i < n; ++i) order[i] = base[i] = i;
    auto less = [&](const U256& a, const U256& b) {
        for (int l = 3; l >= 0; --l)
            if (a.l[l] != b.l[l]) return a.l[l] < b.l[l];
        return false;
    };
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        for (uint32_t k : sort_by) {
            if (k == tuple) {
                if (a != b) return a < b;
                continue;
            }
            const U256 &va = values[(size_t)a * tuple + k], &vb = values[(size_t)b * tuple + k];
            if (less(va, vb)) return true;
            if (less(vb, va)) return false;
        }
        return false;
    });
    return route(base, order, n, bits);
}

}  // namespace psort
}  // namespace acvmb
