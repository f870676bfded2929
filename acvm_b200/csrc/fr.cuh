// K0: BN254 scalar-field (Fr) device library -- 8 x 32-bit limbs, little-endian limb order.
//
// Replaces (reference): acir_field::FieldElement over ark_bn254::Fr = Fp256<MontBackend<FrConfig,4>>
//   mul  acir_field/src/generic_ark.rs:367-373      add :383-394      sub :396-406      neg :360-365
//   inverse :242-245 (0 -> 0)                         (ark-ff 0.4.2, crates.io, not vendored)
//
// Design notes (B200 / sm_100a):
//  * 64-bit `mad.lo/hi.cc.u64` is emulated by ptxas on the 32-bit IMAD pipe, so limbs are 32-bit and
//    every 32x32->64 partial product is written as the PTX pair  mad(c).lo.cc.u32 / madc.hi.cc.u32
//    on the same operands, which ptxas fuses into ONE  IMAD.WIDE.U32(.X)  with carry in/out.
//  * Products are accumulated in two interleaved carry chains ("even"/"odd" columns) so that no
//    partial product ever needs a carry ripple through more than its own chain.
//  * The workhorse is mont_dot<K>:  sum_k a_k*b_k / 2^256 mod p with ONE interleaved Montgomery
//    reduction for all K products (K*64 + 72 wide IMADs instead of K*136).
//  * Witness values live in HBM in CANONICAL form; plan-time constants carry the Montgomery
//    factors (c*R or c*R^2), so no to/from-Montgomery conversion ever touches witness data.
//
// The same header compiles for the host (carry flag emulated) so the limb algorithms are unit
// tested on CPU against Python big ints (tests/test_fr_host.py) before any GPU time is spent.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FR_HD __host__ __device__ __forceinline__
#define FR_D __device__ __forceinline__
#else
#define FR_HD inline
#define FR_D inline
#endif

namespace fr {

constexpr int N = 8;

// p = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
#define FR_P0 0xf0000001u
#define FR_P1 0x43e1f593u
#define FR_P2 0x79b97091u
#define FR_P3 0x2833e848u
#define FR_P4 0x8181585du
#define FR_P5 0xb85045b6u
#define FR_P6 0xe131a029u
#define FR_P7 0x30644e72u
#define FR_M0 0xefffffffu  // -p^{-1} mod 2^32

template <int I> struct PLimb;
template <> struct PLimb<0> { static constexpr uint32_t v = FR_P0; };
template <> struct PLimb<1> { static constexpr uint32_t v = FR_P1; };
template <> struct PLimb<2> { static constexpr uint32_t v = FR_P2; };
template <> struct PLimb<3> { static constexpr uint32_t v = FR_P3; };
template <> struct PLimb<4> { static constexpr uint32_t v = FR_P4; };
template <> struct PLimb<5> { static constexpr uint32_t v = FR_P5; };
template <> struct PLimb<6> { static constexpr uint32_t v = FR_P6; };
template <> struct PLimb<7> { static constexpr uint32_t v = FR_P7; };

FR_HD uint32_t p_limb(int i) {
    switch (i) {
        case 0: return FR_P0; case 1: return FR_P1; case 2: return FR_P2; case 3: return FR_P3;
        case 4: return FR_P4; case 5: return FR_P5; case 6: return FR_P6; default: return FR_P7;
    }
}

// ---------------------------------------------------------------------------------------------
// carry-chain primitives.  Device: one PTX instruction each, the carry lives in CC.CF.
// Host: identical semantics with the flag in a thread_local (test build only).
// ---------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define FR_PRIM __device__ __forceinline__
FR_PRIM void mul_lo(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
FR_PRIM void mul_hi(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
FR_PRIM void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) { asm volatile("{ .reg .u64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t; }" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b)); }
FR_PRIM void mad_lo_cc(uint32_t& r, uint32_t a, uint32_t b, uint32_t c) { asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); }
FR_PRIM void madc_lo_cc(uint32_t& r, uint32_t a, uint32_t b, uint32_t c) { asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); }
FR_PRIM void madc_hi_cc(uint32_t& r, uint32_t a, uint32_t b, uint32_t c) { asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); }
FR_PRIM void madc_hi(uint32_t& r, uint32_t a, uint32_t b, uint32_t c) { asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); }
FR_PRIM void add_cc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
FR_PRIM void addc_cc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
FR_PRIM void addc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
FR_PRIM void sub_cc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
FR_PRIM void subc_cc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
FR_PRIM void subc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
#else
#define FR_PRIM inline
namespace hostcc { static thread_local uint32_t cf = 0; }
FR_PRIM void mul_lo(uint32_t& r, uint32_t a, uint32_t b) { r = (uint32_t)((uint64_t)a * b); }
FR_PRIM void mul_hi(uint32_t& r, uint32_t a, uint32_t b) { r = (uint32_t)(((uint64_t)a * b) >> 32); }
FR_PRIM void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a * b; lo = (uint32_t)t; hi = (uint32_t)(t >> 32); }
FR_PRIM void mad_lo_cc(uint32_t& r, uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)((uint64_t)a * b) + c; r = (uint32_t)t; hostcc::cf = (uint32_t)(t >> 32); }
FR_PRIM void madc_lo_cc(uint32_t& r, uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)((uint64_t)a * b) + c + hostcc::cf; r = (uint32_t)t; hostcc::cf = (uint32_t)(t >> 32); }
FR_PRIM void madc_hi_cc(uint32_t& r, uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (((uint64_t)a * b) >> 32) + c + hostcc::cf; r = (uint32_t)t; hostcc::cf = (uint32_t)(t >> 32); }
FR_PRIM void madc_hi(uint32_t& r, uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (((uint64_t)a * b) >> 32) + c + hostcc::cf; r = (uint32_t)t; }
FR_PRIM void add_cc(uint32_t& r, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; r = (uint32_t)t; hostcc::cf = (uint32_t)(t >> 32); }
FR_PRIM void addc_cc(uint32_t& r, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + hostcc::cf; r = (uint32_t)t; hostcc::cf = (uint32_t)(t >> 32); }
FR_PRIM void addc(uint32_t& r, uint32_t a, uint32_t b) { r = a + b + hostcc::cf; }
FR_PRIM void sub_cc(uint32_t& r, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; r = (uint32_t)t; hostcc::cf = (uint32_t)(t >> 63); }
FR_PRIM void subc_cc(uint32_t& r, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - hostcc::cf; r = (uint32_t)t; hostcc::cf = (uint32_t)(t >> 63); }
FR_PRIM void subc(uint32_t& r, uint32_t a, uint32_t b) { r = a - b - hostcc::cf; }
#endif

struct Fe {
    uint32_t l[N];
};

// ---------------------------------------------------------------------------------------------
// row helpers (all loops fully unrolled: limb indices are compile-time register names)
// ---------------------------------------------------------------------------------------------
// acc[j],acc[j+1] = a[j]*bi  for j = 0,2,..,6
FR_PRIM void mul_n(uint32_t* acc, const uint32_t* a, uint32_t bi) {
#pragma unroll
    for (int j = 0; j < N; j += 2) {
        mul_lo(acc[j], a[j], bi);
        mul_hi(acc[j + 1], a[j], bi);
    }
}

// acc[j],acc[j+1] += a[j]*bi for j = 0,2,..,6 as ONE carry chain; leaves the carry-out in CF
FR_PRIM void cmad_n(uint32_t* acc, const uint32_t* a, uint32_t bi) {
    mad_lo_cc(acc[0], a[0], bi, acc[0]);
    madc_hi_cc(acc[1], a[0], bi, acc[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
        madc_lo_cc(acc[j], a[j], bi, acc[j]);
        madc_hi_cc(acc[j + 1], a[j], bi, acc[j + 1]);
    }
}

// Same chain, split across the two integer pipes: the four wide products are carry-free IMAD.WIDE.U32
// (full rate on the FMA-heavy pipe) and the eight limb additions run as an IADD3/IADD3.X chain on the ALU
// pipe.  Measured on B200: IMAD.WIDE.U32.X (carry in/out) issues at HALF the rate of the carry-free form; the
// hypothesis that moving chains to the ALU pipe would help was tested and REJECTED (see FR_ALU_SPLIT below).
// SPLIT level (template parameter of dot_row / mont_dot_fn): 0 = every chain on the FMA pipe, 1 = even
// a-chains on the ALU pipe, 2 = + even modulus chains, 3 = + odd modulus chains, 4 = + odd a-chains.
// MEASURED (B200, profiles/r1_pipe_split_sweep.txt): every level > 0 is SLOWER (67.6 -> 46.8 G Fr-mul/s from level
// 0 to 4): carry-consuming IADD3.X chains are no cheaper than IMAD.WIDE.U32.X, so the default stays 0 and the
// levels remain only as a tuning/measurement hook.
#ifndef FR_ALU_SPLIT
#define FR_ALU_SPLIT 0
#endif
FR_PRIM void cmad_n_alu(uint32_t* acc, const uint32_t* a, uint32_t bi) {
    uint32_t lo[N / 2], hi[N / 2];
#pragma unroll
    for (int j = 0; j < N / 2; ++j) mul_wide(lo[j], hi[j], a[2 * j], bi);
    add_cc(acc[0], acc[0], lo[0]);
    addc_cc(acc[1], acc[1], hi[0]);
#pragma unroll
    for (int j = 1; j < N / 2; ++j) {
        addc_cc(acc[2 * j], acc[2 * j], lo[j]);
        addc_cc(acc[2 * j + 1], acc[2 * j + 1], hi[j]);
    }
}

// same, with the modulus as the multiplicand (immediates); OFF selects p[j+OFF]
template <int OFF>
FR_PRIM void cmad_p(uint32_t* acc, uint32_t mi) {
    mad_lo_cc(acc[0], p_limb(OFF), mi, acc[0]);
    madc_hi_cc(acc[1], p_limb(OFF), mi, acc[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
        if (j + OFF < N) {
            madc_lo_cc(acc[j], p_limb(j + OFF), mi, acc[j]);
            madc_hi_cc(acc[j + 1], p_limb(j + OFF), mi, acc[j + 1]);
        }
    }
}

template <int OFF>
FR_PRIM void cmad_p_alu(uint32_t* acc, uint32_t mi) {
    uint32_t lo[N / 2], hi[N / 2];
#pragma unroll
    for (int j = 0; j < N / 2; ++j) mul_wide(lo[j], hi[j], p_limb(2 * j + OFF), mi);
    add_cc(acc[0], acc[0], lo[0]);
    addc_cc(acc[1], acc[1], hi[0]);
#pragma unroll
    for (int j = 1; j < N / 2; ++j) {
        addc_cc(acc[2 * j], acc[2 * j], lo[j]);
        addc_cc(acc[2 * j + 1], acc[2 * j + 1], hi[j]);
    }
}

// odd := (odd >> 64) + a_odd*bi, consuming the incoming carry (column shift of the CIOS step)
FR_PRIM void madc_n_rshift(uint32_t* odd, const uint32_t* a1, uint32_t bi) {
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) {
        madc_lo_cc(odd[j], a1[j], bi, odd[j + 2]);
        madc_hi_cc(odd[j + 1], a1[j], bi, odd[j + 3]);
    }
    madc_lo_cc(odd[N - 2], a1[N - 2], bi, 0);
    madc_hi(odd[N - 1], a1[N - 2], bi, 0);
}

// odd := (odd >> 64) + a_odd*bi on the ALU pipe (carry-free wide products + IADD3.X chain)
FR_PRIM void madc_n_rshift_alu(uint32_t* odd, const uint32_t* a1, uint32_t bi) {
    uint32_t lo[N / 2], hi[N / 2];
#pragma unroll
    for (int j = 0; j < N / 2; ++j) mul_wide(lo[j], hi[j], a1[2 * j], bi);
#pragma unroll
    for (int j = 0; j < N / 2 - 1; ++j) {
        addc_cc(odd[2 * j], odd[2 * j + 2], lo[j]);
        addc_cc(odd[2 * j + 1], odd[2 * j + 3], hi[j]);
    }
    addc_cc(odd[N - 2], lo[N / 2 - 1], 0);
    addc(odd[N - 1], hi[N / 2 - 1], 0);
}

// One CIOS row for a K-term dot product: acc += sum_k a_k * b_k[i]; then one reduction row.
// `even` holds columns 0..7, `odd` columns 1..8; the two swap roles every row.
template <int K, int SPLIT>
FR_PRIM void dot_row(uint32_t* even, uint32_t* odd, const Fe* const* a, const uint32_t* bi, bool first, bool has_init = false) {
    if (first && has_init) {
        // `even` already holds the initial accumulator (< 2^256): even += a_even*b0 as a chain, its carry joins column 8
        mul_n(odd, a[0]->l + 1, bi[0]);
        if (SPLIT >= 1) cmad_n_alu(even, a[0]->l, bi[0]); else cmad_n(even, a[0]->l, bi[0]);
        addc(odd[N - 1], odd[N - 1], 0);
    } else if (first) {
        mul_n(odd, a[0]->l + 1, bi[0]);
        mul_n(even, a[0]->l, bi[0]);
    } else {
        add_cc(even[0], even[0], odd[1]);
        if (SPLIT >= 4) madc_n_rshift_alu(odd, a[0]->l + 1, bi[0]); else madc_n_rshift(odd, a[0]->l + 1, bi[0]);
        if (SPLIT >= 1) cmad_n_alu(even, a[0]->l, bi[0]); else cmad_n(even, a[0]->l, bi[0]);
        addc(odd[N - 1], odd[N - 1], 0);
    }
#pragma unroll
    for (int k = 1; k < K; ++k) {
        if (SPLIT >= 4) cmad_n_alu(odd, a[k]->l + 1, bi[k]); else cmad_n(odd, a[k]->l + 1, bi[k]);  // a_k[1],a_k[3],a_k[5],a_k[7]
        if (SPLIT >= 1) cmad_n_alu(even, a[k]->l, bi[k]); else cmad_n(even, a[k]->l, bi[k]);
        addc(odd[N - 1], odd[N - 1], 0);
    }
    uint32_t mi = even[0] * FR_M0;
    if (SPLIT >= 3) cmad_p_alu<1>(odd, mi); else cmad_p<1>(odd, mi);
    if (SPLIT >= 2) cmad_p_alu<0>(even, mi); else cmad_p<0>(even, mi);
    addc(odd[N - 1], odd[N - 1], 0);
}

// r = (init + sum_k a_k*b_k) * 2^-256 mod p, result in [0, 2p) provided sum_k a_k*b_k < 4.5 p^2; `init`
// (8 limbs, < 2^256, may be null) is a plan constant in Montgomery form, i.e. the gate's additive constant folded into the
// reduction: (c*R + sum)/R = c + sum/R, which saves the separate modular addition and its conditional subtraction.
// `bl(k, i)` returns limb i of b_k: the b operands may live in registers OR be fetched limb by limb
// from shared memory (the plan-time coefficients), which keeps them out of the register file.
template <int K, typename BL, int SPLIT = FR_ALU_SPLIT>
FR_PRIM void mont_dot_fn(Fe& r, const Fe* const* a, BL bl, const uint32_t* init = nullptr) {
    uint32_t even[N], odd[N];
    uint32_t bi[K];
    if (init) {
#pragma unroll
        for (int i = 0; i < N; ++i) even[i] = init[i];
    }
#pragma unroll
    for (int i = 0; i < N; i += 2) {
#pragma unroll
        for (int k = 0; k < K; ++k) bi[k] = bl(k, i);
        dot_row<K, SPLIT>(even, odd, a, bi, i == 0, init != nullptr);
#pragma unroll
        for (int k = 0; k < K; ++k) bi[k] = bl(k, i + 1);
        dot_row<K, SPLIT>(odd, even, a, bi, false);
    }
    // merge: result = even + (odd >> 32)
    add_cc(r.l[0], even[0], odd[1]);
#pragma unroll
    for (int i = 1; i < N - 1; ++i) addc_cc(r.l[i], even[i], odd[i + 1]);
    addc(r.l[N - 1], even[N - 1], 0);
}

struct PtrLimbs {
    const Fe* const* b;
    FR_PRIM uint32_t operator()(int k, int i) const { return b[k]->l[i]; }
};

template <int K>
FR_PRIM void mont_dot_raw(Fe& r, const Fe* const* a, const Fe* const* b) {
    mont_dot_fn<K>(r, a, PtrLimbs{b});
}

// r = a - p if a >= p else a   (a < 2p)
FR_PRIM void cond_sub_p(Fe& a) {
    uint32_t t[N];
    sub_cc(t[0], a.l[0], FR_P0);
    subc_cc(t[1], a.l[1], FR_P1);
    subc_cc(t[2], a.l[2], FR_P2);
    subc_cc(t[3], a.l[3], FR_P3);
    subc_cc(t[4], a.l[4], FR_P4);
    subc_cc(t[5], a.l[5], FR_P5);
    subc_cc(t[6], a.l[6], FR_P6);
    subc_cc(t[7], a.l[7], FR_P7);
    uint32_t borrow;
    subc(borrow, 0, 0);  // 0xffffffff if a < p
#pragma unroll
    for (int i = 0; i < N; ++i) a.l[i] = borrow ? a.l[i] : t[i];
}

// r = a + b (no reduction; caller guarantees a + b < 2^256)
FR_PRIM void add_raw(Fe& r, const Fe& a, const Fe& b) {
    add_cc(r.l[0], a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < N - 1; ++i) addc_cc(r.l[i], a.l[i], b.l[i]);
    addc(r.l[N - 1], a.l[N - 1], b.l[N - 1]);
}

FR_PRIM void add_mod(Fe& r, const Fe& a, const Fe& b) {  // a,b < p
    add_raw(r, a, b);
    cond_sub_p(r);
}

FR_PRIM void sub_mod(Fe& r, const Fe& a, const Fe& b) {  // a,b < p
    uint32_t t[N];
    sub_cc(t[0], a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < N; ++i) subc_cc(t[i], a.l[i], b.l[i]);
    uint32_t borrow;
    subc(borrow, 0, 0);
    // add p back under the borrow mask
    add_cc(r.l[0], t[0], FR_P0 & borrow);
    addc_cc(r.l[1], t[1], FR_P1 & borrow);
    addc_cc(r.l[2], t[2], FR_P2 & borrow);
    addc_cc(r.l[3], t[3], FR_P3 & borrow);
    addc_cc(r.l[4], t[4], FR_P4 & borrow);
    addc_cc(r.l[5], t[5], FR_P5 & borrow);
    addc_cc(r.l[6], t[6], FR_P6 & borrow);
    addc(r.l[7], t[7], FR_P7 & borrow);
}

FR_PRIM bool is_zero(const Fe& a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) o |= a.l[i];
    return o == 0;
}

FR_PRIM bool eq(const Fe& a, const Fe& b) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) o |= a.l[i] ^ b.l[i];
    return o == 0;
}

// a >= p ?
FR_PRIM bool geq_p(const Fe& a) {
    uint32_t t;
    sub_cc(t, a.l[0], FR_P0);
    subc_cc(t, a.l[1], FR_P1);
    subc_cc(t, a.l[2], FR_P2);
    subc_cc(t, a.l[3], FR_P3);
    subc_cc(t, a.l[4], FR_P4);
    subc_cc(t, a.l[5], FR_P5);
    subc_cc(t, a.l[6], FR_P6);
    subc_cc(t, a.l[7], FR_P7);
    uint32_t borrow;
    subc(borrow, 0, 0);
    return borrow == 0;
}

// full reduction of an arbitrary 256-bit value (used by AND/XOR with num_bits >= 254): <= 5 subtractions
FR_PRIM void reduce_256(Fe& a) {
#pragma unroll 1
    for (int it = 0; it < 6; ++it) {
        if (!geq_p(a)) break;
        uint32_t t[N];
        sub_cc(t[0], a.l[0], FR_P0);
        subc_cc(t[1], a.l[1], FR_P1);
        subc_cc(t[2], a.l[2], FR_P2);
        subc_cc(t[3], a.l[3], FR_P3);
        subc_cc(t[4], a.l[4], FR_P4);
        subc_cc(t[5], a.l[5], FR_P5);
        subc_cc(t[6], a.l[6], FR_P6);
        subc(t[7], a.l[7], FR_P7);
#pragma unroll
        for (int i = 0; i < N; ++i) a.l[i] = t[i];
    }
}

// plain Montgomery product, fully reduced: r = a*b/R mod p  (a*b < 4.5p^2)
template <int SPLIT = FR_ALU_SPLIT>
FR_PRIM void mont_mul_s(Fe& r, const Fe& a, const Fe& b) {
    const Fe* aa[1] = {&a};
    const Fe* bb[1] = {&b};
    mont_dot_fn<1, PtrLimbs, SPLIT>(r, aa, PtrLimbs{bb});
    cond_sub_p(r);
}
FR_PRIM void mont_mul(Fe& r, const Fe& a, const Fe& b) { mont_mul_s<FR_ALU_SPLIT>(r, a, b); }

// ---------------------------------------------------------------------------------------------
// Modular inverse by the binary extended Euclid, shaped for SIMT: every iteration is the SAME straight-line code for
// all lanes (conditional swap so that u > v, u -= v, x1 -= x2, then strip ALL trailing zeros of u at once: u >>= k and
// x1 := x1 / 2^k mod p through one Montgomery-style correction x1 + ((-x1 / p) mod 2^k) * p, exact for k <= 32).
// ~190 iterations of ~110 instructions instead of the ~380 Montgomery products of a Fermat ladder; lanes only differ
// in their trip count.  Plain C++ (no carry asm) so the host test build runs the same code.
// a: any representation (the value is inverted as an integer mod p); r = a^-1 mod p, canonical.  a == 0 (mod p) gives 0,
// like FieldElement::inverse (acir_field/src/generic_ark.rs:242-245) -- and, more to the point, terminates: lanes that
// already failed and the padding lanes of a last tile run every micro-op on arbitrary column contents.
// ---------------------------------------------------------------------------------------------
FR_PRIM uint32_t bea_funnel_r(uint32_t lo, uint32_t hi, uint32_t k) {   // (hi:lo) >> k, 0 <= k <= 32
#if defined(__CUDA_ARCH__)
    return __funnelshift_rc(lo, hi, k);
#else
    return k == 0 ? lo : (k >= 32 ? hi : ((lo >> k) | (hi << (32 - k))));
#endif
}
FR_PRIM uint32_t bea_ctz32(uint32_t v) {   // v != 0
#if defined(__CUDA_ARCH__)
    return (uint32_t)(__ffs((int)v) - 1);
#else
    return (uint32_t)__builtin_ctz(v);
#endif
}
// u even, u != 0: u >>= ctz(u), x := x / 2^ctz(u) mod p   (k <= 32 bits per pass; a second pass needs u[0] == 0)
FR_PRIM void bea_strip(uint32_t* u, uint32_t* x) {
    do {
        const uint32_t k = u[0] ? bea_ctz32(u[0]) : 32u;
#pragma unroll
        for (int i = 0; i < N - 1; ++i) u[i] = bea_funnel_r(u[i], u[i + 1], k);
        u[N - 1] = bea_funnel_r(u[N - 1], 0u, k);
        // x + m*p is divisible by 2^k for m = (-x / p) mod 2^k; the sum has 9 limbs and is < 2^k * p
        const uint32_t m = (x[0] * FR_M0) & (k == 32u ? 0xFFFFFFFFu : ((1u << k) - 1u));
        uint32_t t[N + 1];
#pragma unroll
        for (int i = 0; i < N; ++i) t[i] = x[i];
        cmad_p<0>(t, m);          // columns (0,1) (2,3) (4,5) (6,7) += m * p0, p2, p4, p6
        addc(t[N], 0u, 0u);
        cmad_p<1>(t + 1, m);      // columns (1,2) (3,4) (5,6) (7,8) += m * p1, p3, p5, p7
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = bea_funnel_r(t[i], t[i + 1], k);
    } while (!(u[0] & 1u));
}
FR_PRIM void inv_bea(Fe& r, const Fe& a) {
    uint32_t u[N], v[N], x1[N], x2[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { u[i] = a.l[i]; v[i] = p_limb(i); x1[i] = i == 0; x2[i] = 0; }
    {
        uint32_t any = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) any |= u[i];
        if (any == 0) {
#pragma unroll
            for (int i = 0; i < N; ++i) r.l[i] = 0;
            return;
        }
    }
    if (!(u[0] & 1u)) bea_strip(u, x1);
    while (true) {
        // d = u - v; when it borrows the roles swap: (u, x1) <-> (v, x2) and d = v - u
        uint32_t d[N], lt;
        sub_cc(d[0], u[0], v[0]);
#pragma unroll
        for (int i = 1; i < N; ++i) subc_cc(d[i], u[i], v[i]);
        subc(lt, 0u, 0u);   // 0xffffffff if u < v
        if (lt) {
            sub_cc(d[0], v[0], u[0]);
#pragma unroll
            for (int i = 1; i < N - 1; ++i) subc_cc(d[i], v[i], u[i]);
            subc(d[N - 1], v[N - 1], u[N - 1]);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                v[i] = u[i];
                const uint32_t t = x1[i];
                x1[i] = x2[i];
                x2[i] = t;
            }
        }
        uint32_t rest = v[0] ^ 1u;
#pragma unroll
        for (int i = 1; i < N; ++i) rest |= v[i];
        if (rest == 0) break;   // min(u, v) == 1: x2 * a == 1
        uint32_t dz = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) { u[i] = d[i]; dz |= d[i]; }
        if (dz == 0) {          // u == v != 1: gcd(a, p) = p, i.e. a == 0 (mod p) -> 0
#pragma unroll
            for (int i = 0; i < N; ++i) x2[i] = 0;
            break;
        }
        // x1 = x1 - x2 mod p
        uint32_t bm;
        sub_cc(x1[0], x1[0], x2[0]);
#pragma unroll
        for (int i = 1; i < N; ++i) subc_cc(x1[i], x1[i], x2[i]);
        subc(bm, 0u, 0u);
        add_cc(x1[0], x1[0], FR_P0 & bm);
        addc_cc(x1[1], x1[1], FR_P1 & bm);
        addc_cc(x1[2], x1[2], FR_P2 & bm);
        addc_cc(x1[3], x1[3], FR_P3 & bm);
        addc_cc(x1[4], x1[4], FR_P4 & bm);
        addc_cc(x1[5], x1[5], FR_P5 & bm);
        addc_cc(x1[6], x1[6], FR_P6 & bm);
        addc(x1[7], x1[7], FR_P7 & bm);
        bea_strip(u, x1);   // u - v of two odd numbers is even and non-zero
    }
#pragma unroll
    for (int i = 0; i < N; ++i) r.l[i] = x2[i];
}

// number of significant bits of a canonical value (acir_field/src/generic_ark.rs:214-221)
FR_PRIM uint32_t num_bits(const Fe& a) {
    uint32_t nb = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        if (a.l[i]) {
#if defined(__CUDA_ARCH__)
            nb = 32 * i + (32 - __clz(a.l[i]));
#else
            nb = 32 * i + (32 - __builtin_clz(a.l[i]));
#endif
        }
    }
    return nb;
}

}  // namespace fr
