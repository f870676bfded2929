// ECDSA signature verification over secp256k1 / secp256r1, one lane per call-instance.
//
// Replaces (reference): acvm/src/pwg/blackbox/signature/ecdsa.rs:12-97 -> blackbox_solver/src/lib.rs:67-83 ->
// verify_secp256{k1,r1}_ecdsa_signature (blackbox_solver/src/lib.rs:101-210) on k256 0.11.6 / p256 0.11.1.
// The call-site semantics that matter for parity (which inputs panic, compressed-point public key, low-S rule, R.x
// compared without reduction mod n) are spelled out in oracle/ecdsa.py; this file follows the same order.
//
// Arithmetic: generic 256-bit Montgomery (R = 2^256, 8 x 32-bit limbs, modulus read from a constant table -- the four
// moduli here have the top bit set, so unlike fr.cuh the accumulator carries a ninth word), Jacobian coordinates with
// the curve's `a` as a table constant, Shamir's trick over {G, P, G+P}.  Roughly 5.6 k Montgomery products per
// verification; the op is rare in circuits and is not on the measured hot path, so the code is written for size
// (rolled outer loops, plain C++) and compiles for the host too, where tests/test_host_logic.py checks it against
// oracle/ecdsa.py without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define EC_HD __host__ __device__
#define EC_NOINLINE __noinline__   // one SASS body shared by every kernel of the translation unit (compile time, code size)
#else
#define EC_HD
#define EC_NOINLINE
#endif

namespace ec {

struct Mod256 {
    uint32_t m[8];        // modulus
    uint32_t r2[8];       // R^2 mod m
    uint32_t one[8];      // R mod m
    uint32_t m_minus_2[8];
    uint32_t ninv;        // -m^-1 mod 2^32
};
struct EcCurve {
    Mod256 fp, fn;                    // base field, scalar field
    uint32_t a[8], b[8], gx[8], gy[8];   // Montgomery form mod p
    uint32_t sqrt_exp[8];             // (p + 1) / 4   (p = 3 mod 4 on both curves)
    uint32_t half_n[8];               // (n - 1) / 2
};

#if defined(__CUDACC__)
// seen identically by the host and device passes of nvcc; host code (EC_HD functions called on the CPU) uses the copy below
static __constant__ EcCurve g_ec_curves[2] = {
#include "ecdsa_consts.inc"
};
#endif
#if defined(__CUDA_ARCH__)
#define EC_CURVE(i) (ec::g_ec_curves[i])
#else
static const EcCurve h_ec_curves[2] = {
#include "ecdsa_consts.inc"
};
#define EC_CURVE(i) (ec::h_ec_curves[i])
#endif

struct U8 {
    uint32_t l[8];
};

EC_HD inline bool is_zero(const U8& a) {
    uint32_t o = 0;
    for (int i = 0; i < 8; ++i) o |= a.l[i];
    return o == 0;
}
EC_HD inline bool eq(const U8& a, const U8& b) {
    uint32_t o = 0;
    for (int i = 0; i < 8; ++i) o |= a.l[i] ^ b.l[i];
    return o == 0;
}
// a >= b
EC_HD inline bool geq(const uint32_t* a, const uint32_t* b) {
    for (int i = 7; i >= 0; --i) {
        if (a[i] > b[i]) return true;
        if (a[i] < b[i]) return false;
    }
    return true;
}
EC_HD inline void set(U8& r, const uint32_t* a) {
    for (int i = 0; i < 8; ++i) r.l[i] = a[i];
}

// r = a * b / R mod m   (a, b < m)
EC_HD EC_NOINLINE inline void mmul(U8& r, const U8& a, const U8& b, const Mod256& M) {
    uint32_t t[10];
    for (int i = 0; i < 10; ++i) t[i] = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int i = 0; i < 8; ++i) {
        const uint32_t bi = b.l[i];
        unsigned long long c = 0;
        for (int j = 0; j < 8; ++j) {
            unsigned long long v = (unsigned long long)a.l[j] * bi + t[j] + c;
            t[j] = (uint32_t)v;
            c = v >> 32;
        }
        unsigned long long v = (unsigned long long)t[8] + c;
        t[8] = (uint32_t)v;
        t[9] = (uint32_t)(v >> 32);
        const uint32_t q = t[0] * M.ninv;
        v = (unsigned long long)q * M.m[0] + t[0];
        c = v >> 32;
        for (int j = 1; j < 8; ++j) {
            v = (unsigned long long)q * M.m[j] + t[j] + c;
            t[j - 1] = (uint32_t)v;
            c = v >> 32;
        }
        v = (unsigned long long)t[8] + c;
        t[7] = (uint32_t)v;
        t[8] = t[9] + (uint32_t)(v >> 32);
    }
    if (t[8] || geq(t, M.m)) {
        long long br = 0;
        for (int i = 0; i < 8; ++i) {
            long long d = (long long)t[i] - M.m[i] + br;
            t[i] = (uint32_t)d;
            br = d >> 32;
        }
    }
    for (int i = 0; i < 8; ++i) r.l[i] = t[i];
}
EC_HD inline void madd_mod(U8& r, const U8& a, const U8& b, const Mod256& M) {
    uint32_t t[8];
    unsigned long long c = 0;
    for (int i = 0; i < 8; ++i) {
        c += (unsigned long long)a.l[i] + b.l[i];
        t[i] = (uint32_t)c;
        c >>= 32;
    }
    if (c || geq(t, M.m)) {
        long long br = 0;
        for (int i = 0; i < 8; ++i) {
            long long d = (long long)t[i] - M.m[i] + br;
            t[i] = (uint32_t)d;
            br = d >> 32;
        }
    }
    for (int i = 0; i < 8; ++i) r.l[i] = t[i];
}
EC_HD inline void msub_mod(U8& r, const U8& a, const U8& b, const Mod256& M) {
    uint32_t t[8];
    long long br = 0;
    for (int i = 0; i < 8; ++i) {
        long long d = (long long)a.l[i] - b.l[i] + br;
        t[i] = (uint32_t)d;
        br = d >> 32;
    }
    if (br) {
        unsigned long long c = 0;
        for (int i = 0; i < 8; ++i) {
            c += (unsigned long long)t[i] + M.m[i];
            t[i] = (uint32_t)c;
            c >>= 32;
        }
    }
    for (int i = 0; i < 8; ++i) r.l[i] = t[i];
}
EC_HD inline void to_mont(U8& r, const U8& a, const Mod256& M) {
    U8 r2;
    set(r2, M.r2);
    mmul(r, a, r2, M);
}
EC_HD inline void from_mont(U8& r, const U8& a, const Mod256& M) {
    U8 o;
    for (int i = 0; i < 8; ++i) o.l[i] = i == 0;
    mmul(r, a, o, M);
}
// r = a^e (Montgomery in / out), e a plain 256-bit exponent
EC_HD EC_NOINLINE inline void mpow(U8& r, const U8& a, const uint32_t* e, const Mod256& M) {
    U8 acc;
    set(acc, M.one);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int i = 255; i >= 0; --i) {
        mmul(acc, acc, acc, M);
        if ((e[i >> 5] >> (i & 31)) & 1) mmul(acc, acc, a, M);
    }
    r = acc;
}

struct Jac {
    U8 x, y, z;   // z == 0: point at infinity
};

// dbl-2007-bl with a general `a`
EC_HD EC_NOINLINE inline void jdbl(Jac& p, const EcCurve& C) {
    if (is_zero(p.z)) return;
    const Mod256& F = C.fp;
    U8 xx, yy, yyyy, zz, s, m, t, a;
    mmul(xx, p.x, p.x, F);
    mmul(yy, p.y, p.y, F);
    mmul(yyyy, yy, yy, F);
    mmul(zz, p.z, p.z, F);
    madd_mod(s, p.x, yy, F);
    mmul(s, s, s, F);
    msub_mod(s, s, xx, F);
    msub_mod(s, s, yyyy, F);
    madd_mod(s, s, s, F);
    mmul(m, zz, zz, F);
    set(a, C.a);
    mmul(m, m, a, F);
    madd_mod(t, xx, xx, F);
    madd_mod(t, t, xx, F);
    madd_mod(m, m, t, F);
    // Z3 before Y is overwritten
    madd_mod(t, p.y, p.z, F);
    mmul(t, t, t, F);
    msub_mod(t, t, yy, F);
    msub_mod(p.z, t, zz, F);
    mmul(t, m, m, F);
    msub_mod(t, t, s, F);
    msub_mod(p.x, t, s, F);
    msub_mod(t, s, p.x, F);
    mmul(t, t, m, F);
    madd_mod(yyyy, yyyy, yyyy, F);
    madd_mod(yyyy, yyyy, yyyy, F);
    madd_mod(yyyy, yyyy, yyyy, F);
    msub_mod(p.y, t, yyyy, F);
}

// p += (qx, qy) affine, finite; handles p at infinity, p == q (doubling) and p == -q (infinity)
EC_HD EC_NOINLINE inline void jmadd(Jac& p, const U8& qx, const U8& qy, const EcCurve& C) {
    const Mod256& F = C.fp;
    if (is_zero(p.z)) {
        p.x = qx;
        p.y = qy;
        set(p.z, F.one);
        return;
    }
    U8 z1z1, u2, s2, h, rr, hh, hhh, v, t;
    mmul(z1z1, p.z, p.z, F);
    mmul(u2, qx, z1z1, F);
    mmul(s2, qy, p.z, F);
    mmul(s2, s2, z1z1, F);
    msub_mod(h, u2, p.x, F);
    msub_mod(rr, s2, p.y, F);
    if (is_zero(h)) {
        if (is_zero(rr)) {
            jdbl(p, C);
        } else {
            for (int i = 0; i < 8; ++i) p.z.l[i] = 0;
        }
        return;
    }
    mmul(hh, h, h, F);
    mmul(hhh, hh, h, F);
    mmul(v, p.x, hh, F);
    mmul(p.z, p.z, h, F);
    mmul(t, rr, rr, F);
    msub_mod(t, t, hhh, F);
    msub_mod(t, t, v, F);
    msub_mod(p.x, t, v, F);
    msub_mod(t, v, p.x, F);
    mmul(t, t, rr, F);
    mmul(hhh, hhh, p.y, F);
    msub_mod(p.y, t, hhh, F);
}

EC_HD inline void be32_to_u8(U8& r, const uint8_t* b) {
    for (int i = 0; i < 8; ++i)
        r.l[i] = ((uint32_t)b[28 - 4 * i] << 24) | ((uint32_t)b[29 - 4 * i] << 16) | ((uint32_t)b[30 - 4 * i] << 8) | b[31 - 4 * i];
}

enum : int { EC_FALSE = 0, EC_TRUE = 1, EC_PANIC = 2 };

// hashed_msg, pkx, pky: 32 big-endian bytes each; sig: r || s.  Returns EC_FALSE / EC_TRUE / EC_PANIC.
EC_HD EC_NOINLINE inline int ecdsa_verify(int curve, const uint8_t* hashed_msg, const uint8_t* pkx, const uint8_t* pky, const uint8_t* sig) {
    const EcCurve& C = EC_CURVE(curve);
    const Mod256 &F = C.fp, &N = C.fn;
    U8 r, s, z, x;
    be32_to_u8(r, sig);
    be32_to_u8(s, sig + 32);
    be32_to_u8(z, hashed_msg);
    be32_to_u8(x, pkx);
    // Signature::try_from(..).unwrap(): r, s in [1, n-1]
    if (is_zero(r) || is_zero(s) || geq(r.l, N.m) || geq(s.l, N.m)) return EC_PANIC;
    // PublicKey::from_encoded_point(compressed).unwrap()
    if (geq(x.l, F.m)) return EC_PANIC;
    U8 px, py, alpha, t;
    to_mont(px, x, F);
    mmul(alpha, px, px, F);
    set(t, C.a);
    madd_mod(alpha, alpha, t, F);
    mmul(alpha, alpha, px, F);          // x^3 + a x
    set(t, C.b);
    madd_mod(alpha, alpha, t, F);
    mpow(py, alpha, C.sqrt_exp, F);
    mmul(t, py, py, F);
    if (!eq(t, alpha)) return EC_PANIC;
    from_mont(t, py, F);
    if ((t.l[0] & 1u) != (uint32_t)(pky[31] & 1)) {
        U8 zero;
        for (int i = 0; i < 8; ++i) zero.l[i] = 0;
        msub_mod(py, zero, py, F);
    }
    // Scalar::from_repr(hashed_msg).unwrap()
    if (geq(z.l, N.m)) return EC_PANIC;
    // low-S rule (after every conversion that can panic)
    if (!geq(C.half_n, s.l)) return EC_FALSE;
    U8 sm, sinv, u1, u2;
    to_mont(sm, s, N);
    mpow(sinv, sm, N.m_minus_2, N);      // Montgomery form of s^-1
    mmul(u1, z, sinv, N);                // z plain * sinv Montgomery / R = plain z / s
    mmul(u2, r, sinv, N);
    // table: G, P, G + P
    U8 gx, gy, gpx, gpy;
    set(gx, C.gx);
    set(gy, C.gy);
    Jac gp;
    gp.x = gx;
    gp.y = gy;
    set(gp.z, F.one);
    jmadd(gp, px, py, C);
    const bool gp_inf = is_zero(gp.z);
    if (!gp_inf) {
        U8 zi, zi2;
        mpow(zi, gp.z, F.m_minus_2, F);
        mmul(zi2, zi, zi, F);
        mmul(gpx, gp.x, zi2, F);
        mmul(zi2, zi2, zi, F);
        mmul(gpy, gp.y, zi2, F);
    }
    Jac acc;
    for (int i = 0; i < 8; ++i) acc.x.l[i] = acc.y.l[i] = acc.z.l[i] = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int i = 255; i >= 0; --i) {
        jdbl(acc, C);
        const uint32_t sel = ((u1.l[i >> 5] >> (i & 31)) & 1u) | (((u2.l[i >> 5] >> (i & 31)) & 1u) << 1);
        if (sel == 1) jmadd(acc, gx, gy, C);
        else if (sel == 2) jmadd(acc, px, py, C);
        else if (sel == 3 && !gp_inf) jmadd(acc, gpx, gpy, C);
    }
    if (is_zero(acc.z)) return EC_PANIC;             // unreachable!("Point is uncompressed")
    U8 zi, rx;
    mpow(zi, acc.z, F.m_minus_2, F);
    mmul(zi, zi, zi, F);
    mmul(rx, acc.x, zi, F);
    from_mont(rx, rx, F);
    if (geq(rx.l, N.m)) return EC_PANIC;             // Scalar::from_repr(x).unwrap()
    return eq(rx, r) ? EC_TRUE : EC_FALSE;
}

}  // namespace ec
