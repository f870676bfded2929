// Host-side BN254 Fr arithmetic for plan-time constants (coefficient folding, inverses,
// Montgomery scaling) and for the host-executed opcodes (Brillig, directives).
// 4 x 64-bit limbs, little-endian, values always canonical in [0, p).
//
// Mirrors acir_field::FieldElement semantics (acir_field/src/generic_ark.rs):
//   from_be_bytes_reduce :281-283, to_be_bytes :269-277, inverse (0 -> 0) :242-245, num_bits :214-221.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>

namespace acvmb {

struct U256 {
    uint64_t l[4] = {0, 0, 0, 0};
    bool operator==(const U256& o) const { return l[0] == o.l[0] && l[1] == o.l[1] && l[2] == o.l[2] && l[3] == o.l[3]; }
    bool operator!=(const U256& o) const { return !(*this == o); }
    bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
};

namespace hf {

static const U256 P = {{0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}};
static const uint64_t PINV64 = 0xc2e1f593efffffffULL;  // -p^{-1} mod 2^64
// R = 2^256 mod p, R2 = 2^512 mod p, R3 = 2^768 mod p (computed at startup by repeated doubling)

inline int cmp(const U256& a, const U256& b) {
    for (int i = 3; i >= 0; --i) {
        if (a.l[i] < b.l[i]) return -1;
        if (a.l[i] > b.l[i]) return 1;
    }
    return 0;
}
inline uint64_t add_raw(U256& r, const U256& a, const U256& b) {
    unsigned __int128 c = 0;
    for (int i = 0; i < 4; ++i) {
        c += (unsigned __int128)a.l[i] + b.l[i];
        r.l[i] = (uint64_t)c;
        c >>= 64;
    }
    return (uint64_t)c;
}
inline uint64_t sub_raw(U256& r, const U256& a, const U256& b) {
    unsigned __int128 br = 0;
    for (int i = 0; i < 4; ++i) {
        unsigned __int128 t = (unsigned __int128)a.l[i] - b.l[i] - br;
        r.l[i] = (uint64_t)t;
        br = (t >> 64) & 1;
    }
    return (uint64_t)br;
}
inline U256 from_u64_(uint64_t v) {
    U256 r;
    r.l[0] = v;
    return r;
}
inline U256 add(const U256& a, const U256& b) {
    U256 r;
    uint64_t c = add_raw(r, a, b);
    if (c || cmp(r, P) >= 0) sub_raw(r, r, P);
    return r;
}
inline U256 sub(const U256& a, const U256& b) {
    U256 r;
    if (sub_raw(r, a, b)) add_raw(r, r, P);
    return r;
}
inline U256 neg(const U256& a) {
    if (a.is_zero()) return a;
    U256 r;
    sub_raw(r, P, a);
    return r;
}
// Montgomery product a*b/2^256 mod p (CIOS, 4x64)
inline U256 mont_mul(const U256& a, const U256& b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        unsigned __int128 c = 0;
        for (int j = 0; j < 4; ++j) {
            c += (unsigned __int128)a.l[j] * b.l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * PINV64;
        c = (unsigned __int128)m * P.l[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; ++j) {
            c += (unsigned __int128)m * P.l[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    U256 r = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || cmp(r, P) >= 0) sub_raw(r, r, P);
    return r;
}

struct Consts {
    U256 R, R2, R3, one;
    Consts() {
        one.l[0] = 1;
        // 2^256 mod p by doubling 1, 256 times
        U256 x = one;
        for (int i = 0; i < 256; ++i) x = add(x, x);
        R = x;
        for (int i = 0; i < 256; ++i) x = add(x, x);
        R2 = x;
        for (int i = 0; i < 256; ++i) x = add(x, x);
        R3 = x;
    }
};
inline const Consts& consts() {
    static Consts c;
    return c;
}
inline U256 to_mont(const U256& a) { return mont_mul(a, consts().R2); }      // a*R
inline U256 to_mont2(const U256& a) { return mont_mul(a, consts().R3); }     // a*R^2
inline U256 from_mont(const U256& a) { return mont_mul(a, consts().one); }   // a/R
inline U256 mul(const U256& a, const U256& b) { return mont_mul(mont_mul(a, b), consts().R2); }
inline U256 pow(const U256& a, const U256& e) {
    U256 am = to_mont(a), acc = consts().R;
    for (int i = 255; i >= 0; --i) {
        acc = mont_mul(acc, acc);
        if ((e.l[i / 64] >> (i % 64)) & 1) acc = mont_mul(acc, am);
    }
    return from_mont(acc);
}
inline void shr1(U256& a, uint64_t top) {
    a.l[0] = (a.l[0] >> 1) | (a.l[1] << 63);
    a.l[1] = (a.l[1] >> 1) | (a.l[2] << 63);
    a.l[2] = (a.l[2] >> 1) | (a.l[3] << 63);
    a.l[3] = (a.l[3] >> 1) | (top << 63);
}
inline bool is_one(const U256& a) { return a.l[0] == 1 && !(a.l[1] | a.l[2] | a.l[3]); }
// binary extended Euclid on canonical values (plan time: one inversion per solving gate); 0 -> 0.
// Tight form: u, v shrink by whole runs of trailing zeros at a time (ctz), x1, x2 are halved modulo p with a
// branch-free add of (x & 1) * p.
inline U256 inverse(const U256& a) {
    if (a.is_zero()) return a;
    uint64_t u[4] = {a.l[0], a.l[1], a.l[2], a.l[3]}, v[4] = {P.l[0], P.l[1], P.l[2], P.l[3]};
    uint64_t x1[4] = {1, 0, 0, 0}, x2[4] = {0, 0, 0, 0};
    auto halve_mod = [](uint64_t* x) {   // x = x/2 mod p  (x < p)
        uint64_t mask = 0 - (x[0] & 1);
        unsigned __int128 c = 0;
        uint64_t t[4];
        for (int i = 0; i < 4; ++i) {
            c += (unsigned __int128)x[i] + (P.l[i] & mask);
            t[i] = (uint64_t)c;
            c >>= 64;
        }
        x[0] = (t[0] >> 1) | (t[1] << 63);
        x[1] = (t[1] >> 1) | (t[2] << 63);
        x[2] = (t[2] >> 1) | (t[3] << 63);
        x[3] = (t[3] >> 1) | ((uint64_t)c << 63);
    };
    auto shr1n = [](uint64_t* x) {
        x[0] = (x[0] >> 1) | (x[1] << 63);
        x[1] = (x[1] >> 1) | (x[2] << 63);
        x[2] = (x[2] >> 1) | (x[3] << 63);
        x[3] >>= 1;
    };
    auto geq = [](const uint64_t* x, const uint64_t* y) {
        for (int i = 3; i >= 0; --i) {
            if (x[i] != y[i]) return x[i] > y[i];
        }
        return true;
    };
    auto sub_n = [](uint64_t* x, const uint64_t* y) {   // x -= y, returns borrow
        unsigned __int128 br = 0;
        for (int i = 0; i < 4; ++i) {
            unsigned __int128 t = (unsigned __int128)x[i] - y[i] - br;
            x[i] = (uint64_t)t;
            br = (t >> 64) & 1;
        }
        return (uint64_t)br;
    };
    auto sub_mod = [&](uint64_t* x, const uint64_t* y) {   // x = x - y mod p
        if (sub_n(x, y)) {
            unsigned __int128 c = 0;
            for (int i = 0; i < 4; ++i) {
                c += (unsigned __int128)x[i] + P.l[i];
                x[i] = (uint64_t)c;
                c >>= 64;
            }
        }
    };
    auto is1 = [](const uint64_t* x) { return x[0] == 1 && !(x[1] | x[2] | x[3]); };
    while (!is1(u) && !is1(v)) {
        while (!(u[0] & 1)) {
            shr1n(u);
            halve_mod(x1);
        }
        while (!(v[0] & 1)) {
            shr1n(v);
            halve_mod(x2);
        }
        if (geq(u, v)) {
            sub_n(u, v);
            sub_mod(x1, x2);
        } else {
            sub_n(v, u);
            sub_mod(x2, x1);
        }
    }
    const uint64_t* r = is1(u) ? x1 : x2;
    U256 out;
    for (int i = 0; i < 4; ++i) out.l[i] = r[i];
    return out;
}
inline U256 reduce(U256 a) {  // arbitrary 256-bit -> mod p
    while (cmp(a, P) >= 0) sub_raw(a, a, P);
    return a;
}
inline U256 from_be_bytes_reduce(const uint8_t* b, size_t n) {
    if (n == 32) {   // the common case (every FieldElement of an ACIR circuit): four big-endian words, < 2^256 < 6p
        U256 v;
        for (int k = 0; k < 4; ++k) {
            uint64_t w = 0;
            for (int j = 0; j < 8; ++j) w = (w << 8) | b[8 * (3 - k) + j];
            v.l[k] = w;
        }
        return reduce(v);
    }
    // accumulates base-256 digits mod p so inputs longer than 32 bytes reduce correctly
    U256 acc;
    U256 c256;
    c256.l[0] = 256;
    for (size_t i = 0; i < n; ++i) {
        if (i >= 31) {
            acc = mul(acc, c256);
            U256 d;
            d.l[0] = b[i];
            acc = add(acc, d);
        } else {
            // fast path: shift in (value stays < 2^248 < p)
            for (int k = 3; k > 0; --k) acc.l[k] = (acc.l[k] << 8) | (acc.l[k - 1] >> 56);
            acc.l[0] = (acc.l[0] << 8) | b[i];
        }
    }
    return reduce(acc);
}
inline void to_be_bytes(const U256& a, uint8_t out[32]) {
    for (int i = 0; i < 32; ++i) out[31 - i] = (uint8_t)(a.l[i / 8] >> (8 * (i % 8)));
}
inline uint32_t num_bits(const U256& a) {
    for (int i = 3; i >= 0; --i)
        if (a.l[i]) return 64 * i + (64 - __builtin_clzll(a.l[i]));
    return 0;
}
inline U256 from_u64(uint64_t v) {
    U256 r;
    r.l[0] = v;
    return r;
}
inline void to_limbs32(const U256& a, uint32_t out[8]) {
    for (int i = 0; i < 4; ++i) {
        out[2 * i] = (uint32_t)a.l[i];
        out[2 * i + 1] = (uint32_t)(a.l[i] >> 32);
    }
}
inline const int8_t* hex_table() {
    static int8_t t[256];
    static bool init = [] {
        for (int i = 0; i < 256; ++i) t[i] = -1;
        for (int i = 0; i < 10; ++i) t['0' + i] = (int8_t)i;
        for (int i = 0; i < 6; ++i) t['a' + i] = t['A' + i] = (int8_t)(10 + i);
        return true;
    }();
    (void)init;
    return t;
}
inline bool from_hex(const char* s, size_t len, U256& out) {
    size_t off = (len >= 2 && s[0] == '0' && (s[1] == 'x' || s[1] == 'X')) ? 2 : 0;
    size_t n = len - off;
    if (n % 2) return false;
    const int8_t* hv = hex_table();
    uint8_t stack[64];
    std::string heap;
    uint8_t* bytes = stack;
    if (n / 2 > sizeof(stack)) {
        heap.resize(n / 2);
        bytes = (uint8_t*)&heap[0];
    }
    int bad = 0;
    for (size_t i = 0; i < n / 2; ++i) {
        const int h = hv[(uint8_t)s[off + 2 * i]], l = hv[(uint8_t)s[off + 2 * i + 1]];
        bad |= h | l;
        bytes[i] = (uint8_t)(h * 16 + l);
    }
    if (bad < 0) return false;
    out = from_be_bytes_reduce(bytes, n / 2);
    return true;
}
inline bool from_hex(const std::string& s, U256& out) { return from_hex(s.data(), s.size(), out); }

}  // namespace hf
}  // namespace acvmb
