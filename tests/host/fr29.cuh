// K0b: carry-free Montgomery arithmetic for the arithmetic-gate hot loop -- 9 limbs of 29 bits, R = 2^261.
//
// Why a second representation: on B200 the wide multiply-add WITH a carry flag (IMAD.WIDE.U32.X, what a 32-bit-limb
// CIOS needs on almost every partial product) issues at HALF the rate of the carry-free IMAD.WIDE.U32 (measured:
// 9.2 vs 17.8 T/s, profiles/r1_pipe_split_sweep.txt).  With 29-bit limbs a 64-bit column accumulator can absorb every
// partial product of a 4-term dot product (<= 36 products < 2^58 each per column) without any carry, so ALL 171 wide
// multiplies of a Montgomery product are the full-rate form; carries are resolved once per row (one 64-bit shift+add)
// and once at the end.  R = 2^261 >> p also makes every lazy bound trivial: inputs up to 2^257 still give results
// below 1.1 p.
//
// STATUS: measured and NOT used by the gate kernel (profiles/r1_multiplier_experiments.txt).  Although every multiply
// is the full-rate instruction (162 IMAD.WIDE.U32 + 11 IMAD per product in SASS), the representation needs ~180 extra
// ALU instructions per product (64-bit column shifts/adds, limb conversions), and the kernel is bound by per-warp
// dependent-issue latency at the occupancy HBM capacity allows (4 CTAs/SM), not by the FMA pipe: the gate kernel got
// 1.5x SLOWER (21.4 vs 14.1 ms).  Kept as a tested building block and as the record of the experiment.
//
// Replaces the same reference arithmetic as fr.cuh (acir_field/src/generic_ark.rs:360-406 over ark-ff Fp256).
// fr.cuh (8 x 32, R = 2^256) remains the representation of the curve / general-gate code and of the tables.
// The header is plain C++ (no inline asm) and compiles for the host, where tests/test_host_logic.py checks it
// against Python big ints.
#pragma once
#include <stdint.h>

#include "../../acvm_b200/csrc/fr.cuh"

namespace fr29 {

constexpr int L = 9;
constexpr uint32_t MASK = (1u << 29) - 1;
constexpr uint32_t PINV = 0x0fffffffu;   // -p^{-1} mod 2^29

#if defined(__CUDA_ARCH__)
#define FR29_FN __device__ __forceinline__
#else
#define FR29_FN inline
#endif

FR29_FN uint32_t p_limb29(int i) {
    switch (i) {
        case 0: return 0x10000001u; case 1: return 0x1f0fac9fu; case 2: return 0x0e5c2450u; case 3: return 0x07d090f3u;
        case 4: return 0x1585d283u; case 5: return 0x02db40c0u; case 6: return 0x00a6e141u; case 7: return 0x0e5c2634u;
        default: return 0x0030644eu;
    }
}

struct Fe9 {
    uint32_t l[L];
};

FR29_FN uint32_t funnel_r(uint32_t lo, uint32_t hi, int s) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, s);
#else
    return s ? (uint32_t)((((uint64_t)hi << 32) | lo) >> s) : lo;
#endif
}

// 8 x 32-bit words (value < 2^256) -> 9 x 29-bit limbs (normalised)
FR29_FN void to9(Fe9& r, const fr::Fe& a) {
#pragma unroll
    for (int i = 0; i < L; ++i) {
        const int bit = 29 * i, w = bit >> 5, s = bit & 31;
        uint32_t lo = a.l[w], hi = (w + 1 < 8) ? a.l[w + 1] : 0u;
        r.l[i] = funnel_r(lo, hi, s) & MASK;
    }
    // limb 8 starts at bit 232 = word 7 bit 8: 24 bits remain, no mask issue
}

// normalised 29-bit limbs (value < 2^256) -> 8 x 32-bit words
FR29_FN void from9(fr::Fe& r, const uint32_t* l) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int bit = 32 * k, i = bit / 29, s = bit % 29;   // word k starts inside limb i at offset s
        uint32_t v = l[i] >> s;                               // 29 - s bits
        if (i + 1 < L) v |= l[i + 1] << (29 - s);             // next 29 bits
        if (29 - s + 29 < 32 && i + 2 < L) v |= l[i + 2] << (58 - s);
        r.l[k] = v;
    }
}

// out = sum_k a_k * b_k * 2^-261 mod p, normalised limbs, value < p + (sum a_k b_k) / 2^261  (see DESIGN.md bounds).
// a_k: register-resident limbs (<= 2^30 each); bl(k, j): limb j of b_k (< 2^29), typically read from shared memory.
template <int K, typename BL>
FR29_FN void mont_dot9(uint32_t* out, const Fe9* const* a, BL bl) {
    unsigned long long t[2 * L + 1];
#pragma unroll
    for (int i = 0; i < 2 * L + 1; ++i) t[i] = 0;
#pragma unroll
    for (int j = 0; j < L; ++j) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint32_t b = bl(k, j);
#pragma unroll
            for (int i = 0; i < L; ++i) t[j + i] += (unsigned long long)a[k]->l[i] * b;
        }
        const uint32_t m = ((uint32_t)t[j] * PINV) & MASK;
#pragma unroll
        for (int i = 0; i < L; ++i) t[j + i] += (unsigned long long)m * p_limb29(i);
        t[j + 1] += t[j] >> 29;     // column j is now 0 mod 2^29
    }
#pragma unroll
    for (int i = 0; i < L; ++i) {
        out[i] = (uint32_t)t[L + i] & MASK;
        t[L + i + 1] += t[L + i] >> 29;
    }
}

struct PtrLimbs9 {
    const Fe9* const* b;
    FR29_FN uint32_t operator()(int k, int j) const { return b[k]->l[j]; }
};

}  // namespace fr29
