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

// ---- Pedersen (plookup-structured) tables ------------------------------------------------------
// Generator derivation is this project's own (the reference's lives inside barretenberg's wasm and is
// not reproducible from the reference tree -- see DESIGN.md "Pedersen parity"): counter-mode
// keccak256(DOMAIN || index_be32 || ctr_be32) -> x mod p, first x on the curve, y parity from the hash.
constexpr int PED_BITS = 9, PED_WINDOWS = 29, PED_TABLE_SIZE = 1 << PED_BITS, PED_IV_SIZE = 1024;

inline void keccak_f_host(uint64_t st[25]) {
    static const uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL,
        0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL,
        0x0000000080008009ULL, 0x000000008000000aULL, 0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL,
        0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    static const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    auto rol = [](uint64_t x, int n) { return n ? ((x << n) | (x >> (64 - n))) : x; };
    for (int rnd = 0; rnd < 24; ++rnd) {
        uint64_t c[5], b[25];
        for (int x = 0; x < 5; ++x) c[x] = st[x] ^ st[x + 5] ^ st[x + 10] ^ st[x + 15] ^ st[x + 20];
        for (int i = 0; i < 25; ++i) st[i] ^= c[(i % 5 + 4) % 5] ^ rol(c[(i % 5 + 1) % 5], 1);
        for (int y = 0; y < 5; ++y)
            for (int x = 0; x < 5; ++x) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol(st[x + 5 * y], ROT[x + 5 * y]);
        for (int y = 0; y < 5; ++y)
            for (int x = 0; x < 5; ++x) st[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
        st[0] ^= RC[rnd];
    }
}

inline void keccak256_host(const uint8_t* msg, size_t n, uint8_t out[32]) {
    uint64_t st[25] = {0};
    size_t pos = 0;
    for (size_t i = 0; i < n; ++i) {
        st[pos >> 3] ^= (uint64_t)msg[i] << (8 * (pos & 7));
        if (++pos == 136) {
            keccak_f_host(st);
            pos = 0;
        }
    }
    st[pos >> 3] ^= 0x01ULL << (8 * (pos & 7));
    st[16] ^= 0x8000000000000000ULL;
    keccak_f_host(st);
    for (int i = 0; i < 32; ++i) out[i] = (uint8_t)(st[i >> 3] >> (8 * (i & 7)));
}

// Tonelli-Shanks; returns false when `a` is a non-residue
inline bool sqrt_fr(const U256& a, U256& root) {
    if (a.is_zero()) {
        root = a;
        return true;
    }
    U256 pm1;
    hf::sub_raw(pm1, hf::P, hf::from_u64(1));
    U256 half = pm1;
    hf::shr1(half, 0);
    U256 one = hf::from_u64(1);
    if (hf::pow(a, half) != one) return false;
    int s = 0;
    U256 q = pm1;
    while (!(q.l[0] & 1)) {
        hf::shr1(q, 0);
        ++s;
    }
    U256 z = hf::from_u64(5);
    while (hf::pow(z, half) != pm1) z = hf::add(z, one);
    U256 qp1h = q;  // (q + 1) / 2
    hf::add_raw(qp1h, qp1h, one);
    hf::shr1(qp1h, 0);
    int m = s;
    U256 c = hf::pow(z, q), t = hf::pow(a, q), r = hf::pow(a, qp1h);
    while (t != one) {
        int i = 0;
        U256 t2 = t;
        while (t2 != one) {
            t2 = hf::mul(t2, t2);
            ++i;
        }
        U256 b = c;
        for (int k = 0; k < m - i - 1; ++k) b = hf::mul(b, b);
        m = i;
        c = hf::mul(b, b);
        t = hf::mul(t, c);
        r = hf::mul(r, b);
    }
    root = r;
    return true;
}

inline Pt derive_pedersen_generator(uint32_t index) {
    static const char DOMAIN[] = "acvm_b200.pedersen.v1";
    for (uint32_t ctr = 0;; ++ctr) {
        uint8_t msg[sizeof(DOMAIN) - 1 + 8];
        memcpy(msg, DOMAIN, sizeof(DOMAIN) - 1);
        for (int i = 0; i < 4; ++i) {
            msg[sizeof(DOMAIN) - 1 + i] = (uint8_t)(index >> (24 - 8 * i));
            msg[sizeof(DOMAIN) - 1 + 4 + i] = (uint8_t)(ctr >> (24 - 8 * i));
        }
        uint8_t h[32];
        keccak256_host(msg, sizeof(msg), h);
        Pt p;
        p.x = hf::from_be_bytes_reduce(h, 32);
        U256 rhs = hf::sub(hf::mul(hf::mul(p.x, p.x), p.x), hf::from_u64(17));
        U256 y;
        if (!sqrt_fr(rhs, y) || y.is_zero()) continue;
        if ((y.l[0] & 1) != (uint64_t)(h[0] >> 7)) y = hf::neg(y);
        p.y = y;
        p.inf = false;
        return p;
    }
}

// [2*29 tables][512][16 u32]: entry s of table t = (s + 1) * G_t, Montgomery affine; followed by the
// IV table [1024][8 u32]: canonical x coordinate of (k + 1) * G.
inline std::vector<uint32_t> build_pedersen_tables() {
    const size_t n_tables = 2 * PED_WINDOWS;
    std::vector<uint32_t> t(n_tables * PED_TABLE_SIZE * 16 + (size_t)PED_IV_SIZE * 8);
    for (size_t k = 0; k < n_tables; ++k) {
        Pt g = derive_pedersen_generator((uint32_t)k);
        Pt cur = g;
        for (int s = 0; s < PED_TABLE_SIZE; ++s) {
            put_mont(cur, &t[(k * PED_TABLE_SIZE + s) * 16]);
            cur = add(cur, g);
        }
    }
    uint32_t* iv = &t[n_tables * PED_TABLE_SIZE * 16];
    Pt g = generator(), cur = g;
    for (int k = 0; k < PED_IV_SIZE; ++k) {
        hf::to_limbs32(cur.x, iv + (size_t)k * 8);
        cur = add(cur, g);
    }
    return t;
}

}  // namespace gk
}  // namespace acvmb
