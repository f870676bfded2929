// Host-side Grumpkin arithmetic (affine, canonical field values) used ONCE per context to build
// the fixed-base and Pedersen lookup tables that the kernels read from HBM/L2.
// Curve: y^2 = x^3 - 17 over BN254 Fr; generator G = (1, sqrt(-16)) as pinned by the reference KATs
// (barretenberg_blackbox_solver/src/wasm/scalar_mul.rs:72-83).
#pragma once
#include <vector>

#include "fr_host.hpp"

namespace acvmb {
namespace gk {

struct Pt {
    U256 x, y;
    bool inf = true;
};

inline Pt generator() {
    Pt g;
    g.x = hf::from_u64(1);
    g.y.l[0] = 0x833fc48d823f272cULL;
    g.y.l[1] = 0x2d270d45f1181294ULL;
    g.y.l[2] = 0xcf135e7506a45d63ULL;
    g.y.l[3] = 0x0000000000000002ULL;
    g.inf = false;
    return g;
}

inline Pt dbl(const Pt& p) {
    if (p.inf || p.y.is_zero()) return Pt{};
    U256 three = hf::from_u64(3), two = hf::from_u64(2);
    U256 lam = hf::mul(hf::mul(three, hf::mul(p.x, p.x)), hf::inverse(hf::mul(two, p.y)));
    Pt r;
    r.x = hf::sub(hf::sub(hf::mul(lam, lam), p.x), p.x);
    r.y = hf::sub(hf::mul(lam, hf::sub(p.x, r.x)), p.y);
    r.inf = false;
    return r;
}

inline Pt add(const Pt& a, const Pt& b) {
    if (a.inf) return b;
    if (b.inf) return a;
    if (a.x == b.x) {
        if (a.y == b.y) return dbl(a);
        return Pt{};
    }
    U256 lam = hf::mul(hf::sub(b.y, a.y), hf::inverse(hf::sub(b.x, a.x)));
    Pt r;
    r.x = hf::sub(hf::sub(hf::mul(lam, lam), a.x), b.x);
    r.y = hf::sub(hf::mul(lam, hf::sub(a.x, r.x)), a.y);
    r.inf = false;
    return r;
}

inline Pt neg(const Pt& a) {
    Pt r = a;
    if (!r.inf) r.y = hf::neg(r.y);
    return r;
}

inline Pt mul(const U256& k, Pt p) {
    Pt acc;
    for (int i = 0; i < 256; ++i) {
        if ((k.l[i / 64] >> (i % 64)) & 1) acc = add(acc, p);
        p = dbl(p);
    }
    return acc;
}

inline bool on_curve(const Pt& p) {
    if (p.inf) return true;
    U256 lhs = hf::mul(p.y, p.y);
    U256 rhs = hf::sub(hf::mul(hf::mul(p.x, p.x), p.x), hf::from_u64(17));
    return lhs == rhs;
}

// affine point -> 16 x u32 (x then y) in Montgomery form, the layout the kernels read
inline void put_mont(const Pt& p, uint32_t* out) {
    hf::to_limbs32(hf::to_mont(p.x), out);
    hf::to_limbs32(hf::to_mont(p.y), out + 8);
}

// table[w][d-1] = (d * 256^w) * G  for w in 0..31, d in 1..255
inline std::vector<uint32_t> build_fixed_base_table() {
    std::vector<uint32_t> t((size_t)32 * 255 * 16);
    Pt base = generator();
    for (int w = 0; w < 32; ++w) {
        Pt cur = base;
        for (int d = 1; d <= 255; ++d) {
            put_mont(cur, &t[((size_t)w * 255 + (d - 1)) * 16]);
            cur = add(cur, base);
        }
        base = cur;  // 256 * base
    }
    return t;
}

}  // namespace gk
}  // namespace acvmb
