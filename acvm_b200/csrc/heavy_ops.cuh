// Heavy blackbox micro-ops for the FULL step-VM kernel variant: SHA-256, Keccak-256, Grumpkin
// fixed-base scalar multiplication and Pedersen (plookup) hashing.
//
// Replaces (reference):
//   hash glue     acvm/src/pwg/blackbox/hash.rs:28-103  (get_hash_input / write_digest_to_outputs)
//   sha256        blackbox_solver/src/lib.rs:47-50   -> sha2 0.10.7   (FIPS 180-4)
//   keccak256     blackbox_solver/src/lib.rs:57-60   -> sha3 0.10.8   (Keccak, pad 0x01..0x80, rate 136)
//   fixed base    acvm/src/pwg/blackbox/fixed_base_scalar_mul.rs:11-27 -> barretenberg_blackbox_solver/src/wasm/scalar_mul.rs:17-65
//   pedersen      acvm/src/pwg/blackbox/pedersen.rs:11-28 -> barretenberg_blackbox_solver/src/wasm/pedersen.rs:14-35
//
// One lane executes one call-instance.  Message bytes are gathered straight from the witness
// columns (each message byte is its own 32-byte witness, hash.rs:51-66) and the 32 digest bytes
// are scattered to 32 output columns.
#pragma once
#include "ecdsa.cuh"
#include "fr.cuh"
#include "plan.hpp"

namespace acvmb {

using fr::Fe;

// tables built on the host at context creation (runtime.cu) and shared by every launch
struct CurveTables {
    const uint32_t* fixed_base;   // [32 windows][255][16]  affine (x,y) of (d * 256^w) * G, Montgomery form
    const uint32_t* pedersen;     // see pedersen section
};
static __constant__ CurveTables g_curve_tables;

template <int T>
__device__ __forceinline__ void hv_load(Fe& v, const uint4* cb, uint32_t w) {
    const uint4* p = cb + (size_t)w * (2 * T);
    uint4 lo = p[0], hi = p[T];
    v.l[0] = lo.x; v.l[1] = lo.y; v.l[2] = lo.z; v.l[3] = lo.w;
    v.l[4] = hi.x; v.l[5] = hi.y; v.l[6] = hi.z; v.l[7] = hi.w;
}
template <int T>
__device__ __forceinline__ void hv_store(uint4* cb, uint32_t w, const Fe& v) {
    uint4* p = cb + (size_t)w * (2 * T);
    p[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    p[T] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ void hv_fail(unsigned long long* fail, uint32_t opcode, uint32_t kind, uint32_t aux) {
    unsigned long long key = ((unsigned long long)opcode << 32) | ((unsigned long long)(kind & 0xF) << 28) | (aux & 0x0FFFFFFFu);
    atomicMin(fail, key);
}

template <int T>
__device__ __forceinline__ void insert_value_dev(const OpRec* r, bool check, uint32_t slot, const Fe& v, uint4* cb, unsigned long long* fail);

// write digest byte i to output witness i (insert_value semantics when the output is pre-assigned).
// The output ids live in the payload (global memory): they are fetched eight at a time so that the eight dependent
// L2 round trips overlap instead of adding up (a rolled one-by-one loop made this the longest part of a hash micro-op).
template <int T>
__device__ __forceinline__ void write_digest(const uint8_t* digest, const uint32_t* outs, uint32_t check_mask, uint4* cb,
                                             unsigned long long* fail, uint32_t opcode) {
#pragma unroll 1
    for (int i0 = 0; i0 < 32; i0 += 8) {
        uint32_t o[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) o[g] = outs[i0 + g];
        if (((check_mask >> i0) & 0xFFu) == 0) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                uint4* p = cb + (size_t)o[g] * (2 * T);
                p[0] = make_uint4(digest[i0 + g], 0u, 0u, 0u);
                p[T] = make_uint4(0u, 0u, 0u, 0u);
            }
            continue;
        }
#pragma unroll 1
        for (int g = 0; g < 8; ++g) {
            const int i = i0 + g;
            Fe v;
#pragma unroll
            for (int k = 0; k < 8; ++k) v.l[k] = 0;
            v.l[0] = digest[i];
            if ((check_mask >> i) & 1) {
                Fe old;
                hv_load<T>(old, cb, o[g]);
                if (!fr::eq(old, v)) {  // insert_value replaces the old value before it reports the mismatch
                    hv_fail(fail, opcode, EK_UNSATISFIED_CONSTRAIN, 0);
                    hv_store<T>(cb, o[g], v);
                }
            } else {
                hv_store<T>(cb, o[g], v);
            }
        }
    }
}

// Message gathering shared by the hash kernels (get_hash_input, hash.rs:51-66): input k contributes its low
// ceil(num_bits/8) bytes, little-endian.  Each input is a separate witness column, i.e. a separate dependent global
// load; they are issued in groups of 8 before any byte is consumed so the L2 latencies overlap instead of adding up.
// `take` bounds the number of bytes pushed (Keccak256VariableLength).
template <int T, typename Push>
__device__ __forceinline__ void gather_message(const uint4* cb, const uint32_t* ins, uint32_t n_in, unsigned long long take, Push push) {
#pragma unroll 1
    for (uint32_t k0 = 0; k0 < n_in; k0 += 8) {
        uint32_t w[8], nb[8], low[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const uint32_t k = k0 + g < n_in ? k0 + g : n_in - 1;
            w[g] = ins[2 * k];
            uint32_t b = (ins[2 * k + 1] + 7) >> 3;
            nb[g] = b > 32 ? 32 : b;
        }
#pragma unroll
        for (int g = 0; g < 8; ++g) low[g] = reinterpret_cast<const uint32_t*>(cb + (size_t)w[g] * (2 * T))[0];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            if (k0 + g >= n_in) break;
            if (nb[g] <= 4) {
                for (uint32_t j = 0; j < nb[g] && take; ++j, --take) push((low[g] >> (8 * j)) & 0xFF);
            } else {
                Fe v;
                hv_load<T>(v, cb, w[g]);
#pragma unroll 1
                for (uint32_t j = 0; j < nb[g] && take; ++j, --take) push((v.l[j >> 2] >> (8 * (j & 3))) & 0xFF);
            }
        }
    }
}

// Fast path of the hash micro-ops (payload flag pl[3]): every input is ONE byte, so message byte j is the low byte of
// witness ins[2j].  fetch_words packs bytes [base, base + 4*NW) into NW words, byte j at bits 8*(j & 3) (little-endian),
// zeros past the end of the message.  Sixteen descriptor loads, then sixteen column loads, are in flight together, and
// every index is a compile-time constant, so the block stays in registers (the byte-at-a-time path indexes its block
// dynamically, i.e. through local memory, and cost ~67 instructions per message byte).
template <int T, int NW>
__device__ __forceinline__ void fetch_words(uint32_t* wds, const uint4* cb, const uint32_t* ins, uint32_t base, uint32_t n_in) {
#pragma unroll
    for (int q0 = 0; q0 < NW; q0 += 4) {
        uint32_t id[16], lo[16];
#pragma unroll
        for (int g = 0; g < 16; ++g) {
            const uint32_t k = base + 4 * q0 + g;
            id[g] = (4 * q0 + g < 4 * NW && k < n_in) ? ins[2 * k] : 0xFFFFFFFFu;
        }
#pragma unroll
        for (int g = 0; g < 16; ++g)
            lo[g] = id[g] != 0xFFFFFFFFu ? (reinterpret_cast<const uint32_t*>(cb + (size_t)id[g] * (2 * T))[0] & 0xFFu) : 0u;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (q0 + q < NW) wds[q0 + q] = lo[4 * q] | (lo[4 * q + 1] << 8) | (lo[4 * q + 2] << 16) | (lo[4 * q + 3] << 24);
    }
}

// ---------------------------------------------------------------------------------------------
// SHA-256
// ---------------------------------------------------------------------------------------------
static __constant__ uint32_t SHA_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

__device__ __forceinline__ uint32_t rotr32(uint32_t x, int n) { return __funnelshift_r(x, x, n); }

static __device__ __noinline__ void sha256_compress(uint32_t* h, const uint32_t* blk) {
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = blk[i];
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        uint32_t wi;
        if (i < 16) {
            wi = w[i];
        } else {
            uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
            uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
            uint32_t s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
            wi = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
            w[i & 15] = wi;
        }
        uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + SHA_K[i] + wi;
        uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

// the compression of a block whose K[i] + W[i] are plan-time constants (the padding-only last block of a message of k * 64
// bytes, plan.cpp hash_packed): no message schedule, 16 x 128-bit loads of a table every call of that length shares
static __device__ __noinline__ void sha256_compress_kw(uint32_t* h, const uint32_t* kw) {
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    const uint4* t = reinterpret_cast<const uint4*>(kw);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const uint4 v = __ldg(t + q);
        const uint32_t k4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
            uint32_t ch = (e & f) ^ (~e & g);
            uint32_t t1 = hh + S1 + ch + k4[j];
            uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
            uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
            uint32_t t2 = S0 + mj;
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

// payload: [n_in][check_mask][var_size witness | NONE][0][ (witness, num_bits) * n_in ][ 32 output witnesses ]
// SHA-256 of a message whose bytes are one witness each (payload flag pl[3]); see fetch_words
template <int T>
__device__ __noinline__ void exec_sha256_bytes(const OpRec* r, uint4* cb, unsigned long long* fail, const uint32_t* pl) {
    const uint32_t n_in = pl[0], check_mask = pl[1];
    const uint32_t* ins = pl + 4;
    const uint32_t* outs = ins + 2 * n_in;
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    uint32_t blk[16];
    uint32_t base = 0;
#pragma unroll 1
    for (; n_in - base >= 64; base += 64) {
        fetch_words<T, 16>(blk, cb, ins, base, n_in);
#pragma unroll
        for (int i = 0; i < 16; ++i) blk[i] = __byte_perm(blk[i], 0, 0x0123);
        sha256_compress(h, blk);
    }
    fetch_words<T, 16>(blk, cb, ins, base, n_in);
    const uint32_t rem = n_in - base;   // 0..63 message bytes in the last block(s)
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if ((uint32_t)i == (rem >> 2)) blk[i] |= 0x80u << (8 * (rem & 3));
        blk[i] = __byte_perm(blk[i], 0, 0x0123);
    }
    if (rem >= 56) {
        sha256_compress(h, blk);
#pragma unroll
        for (int i = 0; i < 16; ++i) blk[i] = 0;
    }
    const unsigned long long bits = (unsigned long long)n_in * 8;
    blk[14] = (uint32_t)(bits >> 32);
    blk[15] = (uint32_t)bits;
    sha256_compress(h, blk);
    uint8_t dg[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) dg[i] = (uint8_t)(h[i >> 2] >> (24 - 8 * (i & 3)));
    write_digest<T>(dg, outs, check_mask, cb, fail, r->w[1]);
}

template <int T>
__device__ __noinline__ void exec_sha256(const OpRec* r, uint4* cb, unsigned long long* fail, const uint32_t* payload) {
    const uint32_t* pl = payload + r->w[7];
    const uint32_t n_in = pl[0], check_mask = pl[1];
    const uint32_t* ins = pl + 4;
    const uint32_t* outs = ins + 2 * n_in;
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    uint32_t blk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) blk[i] = 0;
    uint32_t pos = 0;
    unsigned long long total = 0;
    auto push = [&](uint32_t byte) {
        blk[pos >> 2] |= byte << (24 - 8 * (pos & 3));
        if (++pos == 64) {
            sha256_compress(h, blk);
#pragma unroll
            for (int i = 0; i < 16; ++i) blk[i] = 0;
            pos = 0;
        }
    };
    for (uint32_t k = 0; k < n_in; ++k) {
        uint32_t nbytes = (ins[2 * k + 1] + 7) >> 3;   // fetch_nearest_bytes: low ceil(bits/8) bytes, little-endian
        total += nbytes > 32 ? 32 : nbytes;
    }
    if (n_in) gather_message<T>(cb, ins, n_in, ~0ULL, push);
    push(0x80);
    while (pos != 56) push(0);
    unsigned long long bits = total * 8;
    blk[14] = (uint32_t)(bits >> 32);
    blk[15] = (uint32_t)bits;
    sha256_compress(h, blk);
    uint8_t digest[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) digest[i] = (uint8_t)(h[i >> 2] >> (24 - 8 * (i & 3)));
    write_digest<T>(digest, outs, check_mask, cb, fail, r->w[1]);
}

// ---------------------------------------------------------------------------------------------
// BLAKE2s-256, unkeyed (RFC 7693) -- Blake2s opcode and HashToField128Security (blackbox_solver/src/lib.rs:52-55,62-65)
// ---------------------------------------------------------------------------------------------
static __constant__ uint8_t BLAKE2S_SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};

static __device__ __noinline__ void blake2s_compress(uint32_t* h, const uint32_t* m, unsigned long long t, bool last) {
    const uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
    uint32_t v[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = h[i]; v[i + 8] = IV[i]; }
    v[12] ^= (uint32_t)t;
    v[13] ^= (uint32_t)(t >> 32);
    if (last) v[14] = ~v[14];
#define B2S_G(a, b, c, d, x, y)                                       \
    v[a] = v[a] + v[b] + (x); v[d] = rotr32(v[d] ^ v[a], 16);          \
    v[c] = v[c] + v[d];       v[b] = rotr32(v[b] ^ v[c], 12);          \
    v[a] = v[a] + v[b] + (y); v[d] = rotr32(v[d] ^ v[a], 8);           \
    v[c] = v[c] + v[d];       v[b] = rotr32(v[b] ^ v[c], 7);
#pragma unroll 1
    for (int r = 0; r < 10; ++r) {
        const uint8_t* s = BLAKE2S_SIGMA[r];
        B2S_G(0, 4, 8, 12, m[s[0]], m[s[1]])
        B2S_G(1, 5, 9, 13, m[s[2]], m[s[3]])
        B2S_G(2, 6, 10, 14, m[s[4]], m[s[5]])
        B2S_G(3, 7, 11, 15, m[s[6]], m[s[7]])
        B2S_G(0, 5, 10, 15, m[s[8]], m[s[9]])
        B2S_G(1, 6, 11, 12, m[s[10]], m[s[11]])
        B2S_G(2, 7, 8, 13, m[s[12]], m[s[13]])
        B2S_G(3, 4, 9, 14, m[s[14]], m[s[15]])
    }
#undef B2S_G
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[i + 8];
}

// Blake2s / HashToField128Security over a message whose bytes are one witness each (payload flag pl[3]); see fetch_words.
// Whole little-endian blocks; the last one (possibly empty or full) carries the final flag.
template <int T>
__device__ __noinline__ void exec_blake2s_bytes(const OpRec* r, bool to_field, uint4* cb, unsigned long long* fail, const uint32_t* pl) {
    const uint32_t n_in = pl[0], check_mask = pl[1];
    const uint32_t* ins = pl + 4;
    const uint32_t* outs = ins + 2 * n_in;
    uint32_t h[8] = {0x6A09E667u ^ 0x01010020u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
    uint32_t blk[16];
    uint32_t base = 0;
#pragma unroll 1
    for (; n_in - base > 64; base += 64) {
        fetch_words<T, 16>(blk, cb, ins, base, n_in);
        blake2s_compress(h, blk, (unsigned long long)base + 64, false);
    }
    fetch_words<T, 16>(blk, cb, ins, base, n_in);
    blake2s_compress(h, blk, n_in, true);
    if (to_field) {
        Fe f;   // from_be_bytes_reduce(digest): digest byte 0 is the most significant
#pragma unroll
        for (int k = 0; k < 8; ++k) f.l[7 - k] = __byte_perm(h[k], 0, 0x0123);
        fr::reduce_256(f);
        insert_value_dev<T>(r, check_mask & 1, outs[0], f, cb, fail);
    } else {
        uint8_t dg[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) dg[i] = (uint8_t)(h[i >> 2] >> (8 * (i & 3)));
        write_digest<T>(dg, outs, check_mask, cb, fail, r->w[1]);
    }
}

static __device__ __noinline__ void keccak_f1600(unsigned long long* st);

// ---------------------------------------------------------------------------------------------
// Packed hash pipeline (plan.cpp hash_packed): the message arrives as columns of 32 packed bytes (byte i of a chunk at limb
// i >> 2, bits 8 * (i & 3)), the digest leaves as one such column.  Only exec_hash_core sits on the dependency chain of a
// hash chain: two 128-bit loads per 32 message bytes instead of a descriptor load + a column load per byte, one column
// store instead of 64; gathering the message bytes and scattering the digest bytes run on other slot threads.
// ---------------------------------------------------------------------------------------------
template <int T>
__device__ __noinline__ void exec_hash_pack(const OpRec* r, uint4* cb, const uint32_t* payload) {
    const uint32_t* pl = payload + r->w[7];
    const uint32_t n = pl[0];
    uint32_t id[32], lo[32];
#pragma unroll
    for (int g = 0; g < 32; ++g) id[g] = (uint32_t)g < n ? pl[1 + g] : 0xFFFFFFFFu;
#pragma unroll
    for (int g = 0; g < 32; ++g)
        lo[g] = id[g] != 0xFFFFFFFFu ? (reinterpret_cast<const uint32_t*>(cb + (size_t)id[g] * (2 * T))[0] & 0xFFu) : 0u;
    Fe v;
#pragma unroll
    for (int q = 0; q < 8; ++q) v.l[q] = lo[4 * q] | (lo[4 * q + 1] << 8) | (lo[4 * q + 2] << 16) | (lo[4 * q + 3] << 24);
    hv_store<T>(cb, r->w[2], v);
}

// words [8c, 8c + 8) of the message = column chunk[c]
template <int T, int NCHUNK>
__device__ __forceinline__ void load_chunks(uint32_t* wds, const uint4* cb, const uint32_t* chunk, uint32_t first, uint32_t n_chunks) {
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
        Fe v;
        if (first + c < n_chunks) {
            hv_load<T>(v, cb, chunk[first + c]);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) v.l[k] = 0;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) wds[8 * c + k] = v.l[k];
    }
}

template <int T>
__device__ __noinline__ void exec_hash_core(const OpRec* r, uint4* cb, const uint32_t* payload) {
    const uint32_t* pl = r->w[6] == 1 ? &r->c[0][0] : payload + r->w[7];   // descriptor in the record (shared memory) or in the payload
    const uint32_t func = pl[0], n_in = pl[1], n_chunks = pl[2];
    const uint32_t* chunk = pl + 3;
    Fe dg;
    if (func == 0) {   // SHA-256: 64-byte blocks = two chunks
        uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
        uint32_t blk[16];
        uint32_t base = 0;
#pragma unroll 1
        for (; n_in - base >= 64; base += 64) {
            load_chunks<T, 2>(blk, cb, chunk, base >> 5, n_chunks);
#pragma unroll
            for (int i = 0; i < 16; ++i) blk[i] = __byte_perm(blk[i], 0, 0x0123);
            sha256_compress(h, blk);
        }
        if (n_in == base && r->w[5] != 0xFFFFFFFFu) {   // whole blocks only: the padding block's K + W table comes with the plan
            sha256_compress_kw(h, payload + r->w[5]);
        } else {
            load_chunks<T, 2>(blk, cb, chunk, base >> 5, n_chunks);
            const uint32_t rem = n_in - base;   // 0..63 message bytes in the last block(s); bytes past the message are zero
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if ((uint32_t)i == (rem >> 2)) blk[i] |= 0x80u << (8 * (rem & 3));
                blk[i] = __byte_perm(blk[i], 0, 0x0123);
            }
            if (rem >= 56) {
                sha256_compress(h, blk);
#pragma unroll
                for (int i = 0; i < 16; ++i) blk[i] = 0;
            }
            const unsigned long long bits = (unsigned long long)n_in * 8;
            blk[14] = (uint32_t)(bits >> 32);
            blk[15] = (uint32_t)bits;
            sha256_compress(h, blk);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) dg.l[k] = __byte_perm(h[k], 0, 0x0123);   // digest byte 4k at the low end of limb k
    } else if (func == 1) {   // Keccak-256, one block (n_in <= 135): 34 words from five chunks
        unsigned long long st[25];
#pragma unroll
        for (int i = 0; i < 25; ++i) st[i] = 0;
        uint32_t wds[40];
        load_chunks<T, 5>(wds, cb, chunk, 0, n_chunks);
#pragma unroll
        for (int i = 0; i < 17; ++i) {
            unsigned long long lane = ((unsigned long long)wds[2 * i + 1] << 32) | wds[2 * i];
            if ((uint32_t)i == (n_in >> 3)) lane ^= 0x01ULL << (8 * (n_in & 7));
            st[i] ^= lane;
        }
        st[16] ^= 0x8000000000000000ULL;
        keccak_f1600(st);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            dg.l[2 * k] = (uint32_t)st[k];
            dg.l[2 * k + 1] = (uint32_t)(st[k] >> 32);
        }
    } else {   // Blake2s: 64-byte little-endian blocks, the last one (possibly full) carries the final flag
        uint32_t h[8] = {0x6A09E667u ^ 0x01010020u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
        uint32_t blk[16];
        uint32_t base = 0;
#pragma unroll 1
        for (; n_in - base > 64; base += 64) {
            load_chunks<T, 2>(blk, cb, chunk, base >> 5, n_chunks);
            blake2s_compress(h, blk, (unsigned long long)base + 64, false);
        }
        load_chunks<T, 2>(blk, cb, chunk, base >> 5, n_chunks);
        blake2s_compress(h, blk, n_in, true);
#pragma unroll
        for (int k = 0; k < 8; ++k) dg.l[k] = h[k];
    }
    hv_store<T>(cb, r->w[2], dg);
}

template <int T>
__device__ __noinline__ void exec_hash_unpack(const OpRec* r, uint4* cb, unsigned long long* fail, const uint32_t* payload) {
    const uint32_t* pl = payload + r->w[7];
    Fe v;
    hv_load<T>(v, cb, r->w[3]);
    uint8_t dg[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) dg[i] = (uint8_t)(v.l[i >> 2] >> (8 * (i & 3)));
    write_digest<T>(dg, pl + 1, pl[0], cb, fail, r->w[1]);
}

// to_field: one output = digest reduced mod p (HashToField128Security); else 32 byte outputs (Blake2s)
template <int T>
__device__ __noinline__ void exec_blake2s(const OpRec* r, bool to_field, uint4* cb, unsigned long long* fail, const uint32_t* payload) {
    const uint32_t* pl = payload + r->w[7];
    const uint32_t n_in = pl[0], check_mask = pl[1];
    const uint32_t* ins = pl + 4;
    const uint32_t* outs = ins + 2 * n_in;
    uint32_t h[8] = {0x6A09E667u ^ 0x01010020u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
    uint32_t blk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) blk[i] = 0;
    uint32_t pos = 0;
    unsigned long long total = 0;
    auto push = [&](uint32_t byte) {
        if (pos == 64) {   // a full block is only compressed once more input follows (the last block carries the final flag)
            blake2s_compress(h, blk, total, false);
#pragma unroll
            for (int i = 0; i < 16; ++i) blk[i] = 0;
            pos = 0;
        }
        blk[pos >> 2] |= byte << (8 * (pos & 3));
        ++pos;
        ++total;
    };
    if (n_in) gather_message<T>(cb, ins, n_in, ~0ULL, push);
    blake2s_compress(h, blk, total, true);
    if (to_field) {
        Fe f;   // from_be_bytes_reduce(digest): digest byte 0 is the most significant
#pragma unroll
        for (int k = 0; k < 8; ++k) f.l[7 - k] = __byte_perm(h[k], 0, 0x0123);
        fr::reduce_256(f);
        insert_value_dev<T>(r, check_mask & 1, outs[0], f, cb, fail);
    } else {
        uint8_t digest[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) digest[i] = (uint8_t)(h[i >> 2] >> (8 * (i & 3)));
        write_digest<T>(digest, outs, check_mask, cb, fail, r->w[1]);
    }
}

// ---------------------------------------------------------------------------------------------
// Keccak-256 (original Keccak padding 0x01 .. 0x80, rate 136 bytes)
// ---------------------------------------------------------------------------------------------
static __constant__ unsigned long long KECCAK_RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL,
    0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL,
    0x0000000080008009ULL, 0x000000008000000aULL, 0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL,
    0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};

// 64-bit rotate as two 32-bit funnel shifts (the generic shift/or form costs six instructions on the 32-bit datapath); n is a
// compile-time constant wherever this is called from an unrolled loop
__device__ __forceinline__ unsigned long long rol64(unsigned long long x, int n) {
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    if (n >= 32) {
        const uint32_t t = lo;
        lo = hi;
        hi = t;
        n -= 32;
    }
    if (n == 0) return ((unsigned long long)hi << 32) | lo;
    const uint32_t rh = __funnelshift_l(lo, hi, n), rl = __funnelshift_l(hi, lo, n);
    return ((unsigned long long)rh << 32) | rl;
}
// (Tried and rejected: the same rotate as two wide multiplies by 2^n from constant memory plus two IMAD adds, which moves the
// rotates of Keccak-f[1600] from the ALU pipe to the multiplier pipe and balances the two -- 134 ALU + 134 IMAD instructions
// per round instead of 192 + 32.  Measured on the config-3 hash chain: 36.9 ms against 31.4 ms with the funnel shifts.)
#ifndef KECCAK_ROL
#define KECCAK_ROL rol64
#endif
// state index = x + 5*y
static __device__ __noinline__ void keccak_f1600(unsigned long long* st) {
    unsigned long long a[25];
#pragma unroll
    for (int i = 0; i < 25; ++i) a[i] = st[i];
#pragma unroll 1
    for (int rnd = 0; rnd < 24; ++rnd) {
        unsigned long long c[5], d[5], b[25];
#pragma unroll
        for (int x = 0; x < 5; ++x) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
#pragma unroll
        for (int x = 0; x < 5; ++x) d[x] = c[(x + 4) % 5] ^ KECCAK_ROL(c[(x + 1) % 5], 1);
#pragma unroll
        for (int i = 0; i < 25; ++i) a[i] ^= d[i % 5];
        // rho + pi : B[y, 2x+3y] = rot(A[x,y], r[x,y])
        constexpr int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
#pragma unroll
        for (int y = 0; y < 5; ++y)
#pragma unroll
            for (int x = 0; x < 5; ++x) b[y + 5 * ((2 * x + 3 * y) % 5)] = KECCAK_ROL(a[x + 5 * y], ROT[x + 5 * y]);
#pragma unroll
        for (int y = 0; y < 5; ++y)
#pragma unroll
            for (int x = 0; x < 5; ++x) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
        a[0] ^= KECCAK_RC[rnd];
    }
#pragma unroll
    for (int i = 0; i < 25; ++i) st[i] = a[i];
}

// Keccak-256 of a message whose bytes are one witness each (payload flag pl[3]), no length cut; see fetch_words
template <int T>
__device__ __noinline__ void exec_keccak256_bytes(const OpRec* r, uint4* cb, unsigned long long* fail, const uint32_t* pl) {
    const uint32_t n_in = pl[0], check_mask = pl[1];
    const uint32_t* ins = pl + 4;
    const uint32_t* outs = ins + 2 * n_in;
    unsigned long long st[25];
#pragma unroll
    for (int i = 0; i < 25; ++i) st[i] = 0;
    uint32_t base = 0, wds[34];
#pragma unroll 1
    for (; n_in - base >= 136; base += 136) {
        fetch_words<T, 34>(wds, cb, ins, base, n_in);
#pragma unroll
        for (int i = 0; i < 17; ++i) st[i] ^= ((unsigned long long)wds[2 * i + 1] << 32) | wds[2 * i];
        keccak_f1600(st);
    }
    fetch_words<T, 34>(wds, cb, ins, base, n_in);
    const uint32_t rem = n_in - base;   // 0..135
#pragma unroll
    for (int i = 0; i < 17; ++i) {
        unsigned long long lane = ((unsigned long long)wds[2 * i + 1] << 32) | wds[2 * i];
        if ((uint32_t)i == (rem >> 3)) lane ^= 0x01ULL << (8 * (rem & 7));
        st[i] ^= lane;
    }
    st[16] ^= 0x8000000000000000ULL;
    keccak_f1600(st);
    uint8_t dg[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) dg[i] = (uint8_t)(st[i >> 3] >> (8 * (i & 7)));
    write_digest<T>(dg, outs, check_mask, cb, fail, r->w[1]);
}

template <int T>
__device__ __noinline__ void exec_keccak256(const OpRec* r, uint4* cb, unsigned long long* fail, const uint32_t* payload) {
    const uint32_t* pl = payload + r->w[7];
    const uint32_t n_in = pl[0], check_mask = pl[1], var_w = pl[2];
    const uint32_t* ins = pl + 4;
    const uint32_t* outs = ins + 2 * n_in;
    // Keccak256VariableLength: keep only the first to_u128(var_message_size) bytes (hash.rs:67-85)
    unsigned long long take = ~0ULL;
    if (var_w != 0xFFFFFFFFu) {
        Fe sz;
        hv_load<T>(sz, cb, var_w);
        unsigned long long total = 0;
        for (uint32_t k = 0; k < n_in; ++k) {
            uint32_t nb = (ins[2 * k + 1] + 7) >> 3;
            total += nb > 32 ? 32 : nb;
        }
        unsigned long long lo = ((unsigned long long)sz.l[1] << 32) | sz.l[0];
        if ((sz.l[2] | sz.l[3]) || lo > total) {
            hv_fail(fail, r->w[1], EK_BLACKBOX_FAILED, BB_Keccak256);
            return;
        }
        take = lo;
    }
    unsigned long long st[25];
#pragma unroll
    for (int i = 0; i < 25; ++i) st[i] = 0;
    uint32_t pos = 0;
    auto push = [&](uint32_t byte) {
        st[pos >> 3] ^= (unsigned long long)byte << (8 * (pos & 7));
        if (++pos == 136) {
            keccak_f1600(st);
            pos = 0;
        }
    };
    if (n_in) gather_message<T>(cb, ins, n_in, take, push);
    // pad10*1 with the Keccak domain byte 0x01: the two pad bits may share one byte (0x81)
    st[pos >> 3] ^= 0x01ULL << (8 * (pos & 7));
    st[16] ^= 0x8000000000000000ULL;   // byte 135
    keccak_f1600(st);
    uint8_t digest[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) digest[i] = (uint8_t)(st[i >> 3] >> (8 * (i & 7)));
    write_digest<T>(digest, outs, check_mask, cb, fail, r->w[1]);
}

// ---------------------------------------------------------------------------------------------
// Grumpkin: y^2 = x^3 - 17 over BN254 Fr (the VM's own field).  Jacobian accumulator + affine table
// points, all in Montgomery form.
// ---------------------------------------------------------------------------------------------
struct Jac {
    Fe X, Y, Z;
    bool inf;
};

__device__ __forceinline__ void fe_mul(Fe& r, const Fe& a, const Fe& b) { fr::mont_mul(r, a, b); }
__device__ __forceinline__ void fe_sqr(Fe& r, const Fe& a) { fr::mont_mul(r, a, a); }
__device__ __forceinline__ void fe_dbl(Fe& r, const Fe& a) { fr::add_mod(r, a, a); }

// mixed addition acc += (px, py)  ("madd-2007-bl": 7M + 4S); the plan-level invariants of fixed-base /
// Pedersen windows exclude acc == +-P except through the explicit checks below.
static __device__ __noinline__ void jac_madd(Jac& acc, const Fe& px, const Fe& py) {
    if (acc.inf) {
        acc.X = px;
        acc.Y = py;
        // Z = R mod p (Montgomery one)
        acc.Z.l[0] = 0x4ffffffbu; acc.Z.l[1] = 0xac96341cu; acc.Z.l[2] = 0x9f60cd29u; acc.Z.l[3] = 0x36fc7695u;
        acc.Z.l[4] = 0x7879462eu; acc.Z.l[5] = 0x666ea36fu; acc.Z.l[6] = 0x9a07df2fu; acc.Z.l[7] = 0x0e0a77c1u;
        acc.inf = false;
        return;
    }
    Fe z1z1, u2, s2, h, hh, i, j, rr, v, t;
    fe_sqr(z1z1, acc.Z);
    fe_mul(u2, px, z1z1);
    fe_mul(s2, py, acc.Z);
    fe_mul(s2, s2, z1z1);
    fr::sub_mod(h, u2, acc.X);
    fr::sub_mod(rr, s2, acc.Y);
    if (fr::is_zero(h)) {
        if (fr::is_zero(rr)) {
            // doubling (a = 0): "dbl-2009-l"
            Fe A, B, C, D, E, F;
            fe_sqr(A, acc.X);
            fe_sqr(B, acc.Y);
            fe_sqr(C, B);
            fr::add_mod(D, acc.X, B);
            fe_sqr(D, D);
            fr::sub_mod(D, D, A);
            fr::sub_mod(D, D, C);
            fe_dbl(D, D);
            fe_dbl(E, A);
            fr::add_mod(E, E, A);
            fe_sqr(F, E);
            Fe X3, Y3, Z3;
            fe_dbl(t, D);
            fr::sub_mod(X3, F, t);
            fr::sub_mod(t, D, X3);
            fe_mul(Y3, E, t);
            fe_dbl(C, C); fe_dbl(C, C); fe_dbl(C, C);
            fr::sub_mod(Y3, Y3, C);
            fe_mul(Z3, acc.Y, acc.Z);
            fe_dbl(Z3, Z3);
            acc.X = X3; acc.Y = Y3; acc.Z = Z3;
        } else {
            acc.inf = true;
        }
        return;
    }
    fe_sqr(hh, h);
    fe_dbl(i, hh);
    fe_dbl(i, i);             // I = 4*HH
    fe_mul(j, h, i);          // J = H*I
    fe_dbl(rr, rr);           // r = 2*(S2-Y1)
    fe_mul(v, acc.X, i);      // V = X1*I
    Fe X3, Y3, Z3;
    fe_sqr(X3, rr);
    fr::sub_mod(X3, X3, j);
    fe_dbl(t, v);
    fr::sub_mod(X3, X3, t);   // X3 = r^2 - J - 2V
    fr::sub_mod(t, v, X3);
    fe_mul(Y3, rr, t);
    fe_mul(t, acc.Y, j);
    fe_dbl(t, t);
    fr::sub_mod(Y3, Y3, t);   // Y3 = r*(V-X3) - 2*Y1*J
    fr::add_mod(Z3, acc.Z, h);
    fe_sqr(Z3, Z3);
    fr::sub_mod(Z3, Z3, z1z1);
    fr::sub_mod(Z3, Z3, hh);  // Z3 = (Z1+H)^2 - Z1Z1 - HH
    acc.X = X3; acc.Y = Y3; acc.Z = Z3;
}

// Montgomery form of 1/a for a Montgomery-form input a != 0: binary extended Euclid on the integer a*R (fr::inv_bea,
// ~190 uniform iterations) and one product with R^3 to land back in Montgomery form: (aR)^-1 * R^3 / R = a^-1 * R.
// (Was a Fermat ladder: 253 squarings + 127 products, the single largest cost of every curve micro-op.)
static __device__ __noinline__ void fe_inv(Fe& r, const Fe& a) {
    Fe w, r3;
    fr::inv_bea(w, a);
    r3.l[0] = 0xb4bf0040u; r3.l[1] = 0x5e94d8e1u; r3.l[2] = 0x1cfbb6b8u; r3.l[3] = 0x2a489cbeu;
    r3.l[4] = 0xa19fcfedu; r3.l[5] = 0x893cc664u; r3.l[6] = 0x7fcc657cu; r3.l[7] = 0x0cf8594bu;   // R^3 mod p
    fe_mul(r, w, r3);
}

// Jacobian (Montgomery) -> affine canonical
static __device__ __noinline__ void jac_to_affine_canonical(Fe& x, Fe& y, const Jac& p) {
    if (p.inf) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { x.l[k] = 0; y.l[k] = 0; }
        return;
    }
    Fe zi, zi2, zi3, one;
    fe_inv(zi, p.Z);
    fe_sqr(zi2, zi);
    fe_mul(zi3, zi2, zi);
    fe_mul(x, p.X, zi2);
    fe_mul(y, p.Y, zi3);
#pragma unroll
    for (int k = 0; k < 8; ++k) one.l[k] = 0;
    one.l[0] = 1;
    fe_mul(x, x, one);   // from Montgomery
    fe_mul(y, y, one);
}

template <int T>
__device__ __forceinline__ void write_point(const OpRec* r, uint32_t flags, const Fe& x, const Fe& y, uint32_t out_x, uint32_t out_y,
                                            uint4* cb, unsigned long long* fail) {
    // insert_value(outputs.0), insert_value(outputs.1)   (pwg/mod.rs:338-357)
    if (flags & GF_OUT_CHECK) {
        Fe old;
        hv_load<T>(old, cb, out_x);
        if (!fr::eq(old, x)) {
            hv_fail(fail, r->w[1], EK_UNSATISFIED_CONSTRAIN, 0);
            hv_store<T>(cb, out_x, x);
        }
    } else {
        hv_store<T>(cb, out_x, x);
    }
    if (flags & GF_OUT2_CHECK) {
        Fe old;
        hv_load<T>(old, cb, out_y);
        if (!fr::eq(old, y)) {
            hv_fail(fail, r->w[1], EK_UNSATISFIED_CONSTRAIN, 0);
            hv_store<T>(cb, out_y, y);
        }
    } else {
        hv_store<T>(cb, out_y, y);
    }
}

// FixedBaseScalarMul: x = low, y = high, out = w[2] (x coordinate), w[5] = y coordinate slot.
// 8-bit fixed windows over the precomputed table: 32 mixed additions + one inversion.
template <int T>
__device__ __noinline__ void exec_fixed_base(const OpRec* r, uint32_t flags, uint4* cb, unsigned long long* fail) {
    Fe lo, hi;
    hv_load<T>(lo, cb, r->w[3]);
    hv_load<T>(hi, cb, r->w[4]);
    // limbs must be < 2^128 (scalar_mul.rs:25-35)
    if ((lo.l[4] | lo.l[5] | lo.l[6] | lo.l[7]) || (hi.l[4] | hi.l[5] | hi.l[6] | hi.l[7])) {
        hv_fail(fail, r->w[1], EK_BLACKBOX_FAILED, BB_FixedBaseScalarMul);
        return;
    }
    uint32_t s[8] = {lo.l[0], lo.l[1], lo.l[2], lo.l[3], hi.l[0], hi.l[1], hi.l[2], hi.l[3]};
    // scalar < Grumpkin group order n (scalar_mul.rs:40-51)
    const uint32_t n[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    bool lt = false;
#pragma unroll
    for (int i = 7; i >= 0; --i) {
        if (s[i] != n[i]) { lt = s[i] < n[i]; break; }
    }
    if (!lt) {
        hv_fail(fail, r->w[1], EK_BLACKBOX_FAILED, BB_FixedBaseScalarMul);
        return;
    }
    Jac acc;
    acc.inf = true;
#pragma unroll 1
    for (int w = 0; w < 32; ++w) {
        uint32_t d = (s[w >> 2] >> (8 * (w & 3))) & 0xFF;
        if (d == 0) continue;
        const uint4* tp = reinterpret_cast<const uint4*>(g_curve_tables.fixed_base + ((size_t)w * 255 + (d - 1)) * 16);
        uint4 a = tp[0], b = tp[1], c = tp[2], e = tp[3];
        Fe px, py;
        px.l[0] = a.x; px.l[1] = a.y; px.l[2] = a.z; px.l[3] = a.w; px.l[4] = b.x; px.l[5] = b.y; px.l[6] = b.z; px.l[7] = b.w;
        py.l[0] = c.x; py.l[1] = c.y; py.l[2] = c.z; py.l[3] = c.w; py.l[4] = e.x; py.l[5] = e.y; py.l[6] = e.z; py.l[7] = e.w;
        jac_madd(acc, px, py);
    }
    Fe x, y;
    jac_to_affine_canonical(x, y, acc);   // scalar 0 -> (0, 0): encoding of infinity is unpinned in the reference
    write_point<T>(r, flags, x, y, r->w[2], r->w[5], cb, fail);
}

// ---------------------------------------------------------------------------------------------
// Pedersen (plookup-structured): see oracle/pedersen.py for the exact function and DESIGN.md for the
// parity status.  payload: [n_in][domain_separator][witness * n_in]; out = w[2] (x), w[5] (y).
// ---------------------------------------------------------------------------------------------
constexpr int PED_WINDOWS = 29, PED_TABLE_SIZE = 512, PED_IV_SIZE = 1024;

static __device__ __noinline__ void ped_hash_single(Jac& acc, const Fe& v, int parity) {
    uint32_t l[9];
#pragma unroll
    for (int i = 0; i < 8; ++i) l[i] = v.l[i];
    l[8] = 0;
#pragma unroll 1
    for (int i = 0; i < PED_WINDOWS; ++i) {
        const int o = 9 * i, limb = o >> 5, sh = o & 31;
        uint32_t s = __funnelshift_r(l[limb], l[limb + 1], sh) & 0x1FF;
        const uint4* tp = reinterpret_cast<const uint4*>(g_curve_tables.pedersen +
                                                         (((size_t)(parity * PED_WINDOWS + i)) * PED_TABLE_SIZE + s) * 16);
        uint4 a = tp[0], b = tp[1], c = tp[2], e = tp[3];
        Fe px, py;
        px.l[0] = a.x; px.l[1] = a.y; px.l[2] = a.z; px.l[3] = a.w; px.l[4] = b.x; px.l[5] = b.y; px.l[6] = b.z; px.l[7] = b.w;
        py.l[0] = c.x; py.l[1] = c.y; py.l[2] = c.z; py.l[3] = c.w; py.l[4] = e.x; py.l[5] = e.y; py.l[6] = e.z; py.l[7] = e.w;
        jac_madd(acc, px, py);
    }
}

template <int T>
__device__ __noinline__ void exec_pedersen(const OpRec* r, uint32_t flags, uint4* cb, unsigned long long* fail, const uint32_t* payload) {
    const uint32_t* pl = payload + r->w[7];
    const uint32_t n_in = pl[0], iv = pl[1];
    Fe x, y;
    if (n_in == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { x.l[k] = 0; y.l[k] = 0; }
        write_point<T>(r, flags, x, y, r->w[2], r->w[5], cb, fail);
        return;
    }
    Fe rr;
    {
        const uint32_t* ivp = g_curve_tables.pedersen + (size_t)2 * PED_WINDOWS * PED_TABLE_SIZE * 16 + (size_t)(iv % PED_IV_SIZE) * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) rr.l[k] = ivp[k];
    }
    Jac acc;
#pragma unroll 1
    for (uint32_t k = 0; k < n_in; ++k) {
        Fe v;
        hv_load<T>(v, cb, pl[2 + k]);
        acc.inf = true;
        ped_hash_single(acc, rr, 0);
        ped_hash_single(acc, v, 1);
        jac_to_affine_canonical(rr, y, acc);     // chaining value = x coordinate (0 at infinity)
    }
    acc.inf = true;
    ped_hash_single(acc, rr, 0);
    Fe cnt;
#pragma unroll
    for (int k = 0; k < 8; ++k) cnt.l[k] = 0;
    cnt.l[0] = n_in;
    ped_hash_single(acc, cnt, 1);
    jac_to_affine_canonical(x, y, acc);
    write_point<T>(r, flags, x, y, r->w[2], r->w[5], cb, fail);
}

// ---------------------------------------------------------------------------------------------
// Curve micro-ops for plan-level parallelism.  A fixed-base multiplication or a Pedersen chaining round is a sum of
// table points; one lane doing all of it serially (exec_fixed_base / exec_pedersen above) leaves the other S-1 slot
// threads of the instance idle for ~1 ms.  The plan compiler (plan.cpp: curve_sum_*) instead emits
//   MK_CURVE_PART  partial sum of a few table points selected by windows of one scalar  -> Jacobian point in 3 slots
//   MK_JAC_ADD     point + point                                                       -> Jacobian point in 3 slots
//   MK_JAC_FINAL   point (+ point) -> affine canonical (x, y) with insert_value semantics (one inversion)
// and the list scheduler runs the independent ones in the same step.  Jacobian coordinates travel through temporary
// columns in Montgomery form; Z = 0 encodes the point at infinity.
// MK_CURVE_PART parameters (r->c[0]): [0] scalar source: 0 = slot w[3], 1 = Pedersen IV table entry c[0][5], 2 = immediate
// c[0][5];  [1] table: 0 = fixed-base (8-bit windows, digit 0 skipped), 1 = Pedersen (9-bit windows);  [2] first window
// of the scalar;  [3] number of windows;  [4] table window offset (fixed-base high limb: 16; Pedersen parity 1: 29).
// ---------------------------------------------------------------------------------------------
static __device__ __noinline__ void jac_dbl(Jac& acc) {   // "dbl-2009-l", a = 0
    if (acc.inf) return;
    if (fr::is_zero(acc.Y)) { acc.inf = true; return; }
    Fe A, B, C, D, E, F, t, X3, Y3, Z3;
    fe_sqr(A, acc.X);
    fe_sqr(B, acc.Y);
    fe_sqr(C, B);
    fr::add_mod(D, acc.X, B);
    fe_sqr(D, D);
    fr::sub_mod(D, D, A);
    fr::sub_mod(D, D, C);
    fe_dbl(D, D);
    fe_dbl(E, A);
    fr::add_mod(E, E, A);
    fe_sqr(F, E);
    fe_dbl(t, D);
    fr::sub_mod(X3, F, t);
    fr::sub_mod(t, D, X3);
    fe_mul(Y3, E, t);
    fe_dbl(C, C); fe_dbl(C, C); fe_dbl(C, C);
    fr::sub_mod(Y3, Y3, C);
    fe_mul(Z3, acc.Y, acc.Z);
    fe_dbl(Z3, Z3);
    acc.X = X3; acc.Y = Y3; acc.Z = Z3;
}

// a += b, both Jacobian ("add-2007-bl": 11M + 5S), every special case handled
static __device__ __noinline__ void jac_add_full(Jac& a, const Jac& b) {
    if (b.inf) return;
    if (a.inf) { a = b; return; }
    Fe z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v, t;
    fe_sqr(z1z1, a.Z);
    fe_sqr(z2z2, b.Z);
    fe_mul(u1, a.X, z2z2);
    fe_mul(u2, b.X, z1z1);
    fe_mul(s1, a.Y, b.Z);
    fe_mul(s1, s1, z2z2);
    fe_mul(s2, b.Y, a.Z);
    fe_mul(s2, s2, z1z1);
    fr::sub_mod(h, u2, u1);
    fr::sub_mod(rr, s2, s1);
    if (fr::is_zero(h)) {
        if (fr::is_zero(rr)) jac_dbl(a); else a.inf = true;
        return;
    }
    fe_dbl(i, h);
    fe_sqr(i, i);             // I = (2H)^2
    fe_mul(j, h, i);          // J = H*I
    fe_dbl(rr, rr);           // r = 2*(S2-S1)
    fe_mul(v, u1, i);         // V = U1*I
    Fe X3, Y3, Z3;
    fe_sqr(X3, rr);
    fr::sub_mod(X3, X3, j);
    fe_dbl(t, v);
    fr::sub_mod(X3, X3, t);
    fr::sub_mod(t, v, X3);
    fe_mul(Y3, rr, t);
    fe_mul(t, s1, j);
    fe_dbl(t, t);
    fr::sub_mod(Y3, Y3, t);
    fr::add_mod(Z3, a.Z, b.Z);
    fe_sqr(Z3, Z3);
    fr::sub_mod(Z3, Z3, z1z1);
    fr::sub_mod(Z3, Z3, z2z2);
    fe_mul(Z3, Z3, h);
    a.X = X3; a.Y = Y3; a.Z = Z3;
}

template <int T>
__device__ __forceinline__ void load_jac(Jac& p, const uint4* cb, uint32_t base) {
    hv_load<T>(p.X, cb, base);
    hv_load<T>(p.Y, cb, base + 1);
    hv_load<T>(p.Z, cb, base + 2);
    p.inf = fr::is_zero(p.Z);
}
template <int T>
__device__ __forceinline__ void store_jac(uint4* cb, uint32_t base, const Jac& p) {
    if (p.inf) {
        Fe z;
#pragma unroll
        for (int k = 0; k < 8; ++k) z.l[k] = 0;
        hv_store<T>(cb, base, z);
        hv_store<T>(cb, base + 1, z);
        hv_store<T>(cb, base + 2, z);
    } else {
        hv_store<T>(cb, base, p.X);
        hv_store<T>(cb, base + 1, p.Y);
        hv_store<T>(cb, base + 2, p.Z);
    }
}
__device__ __forceinline__ void load_table_point(Fe& px, Fe& py, const uint32_t* table, size_t row) {
    const uint4* tp = reinterpret_cast<const uint4*>(table + row * 16);
    uint4 a = tp[0], b = tp[1], c = tp[2], e = tp[3];
    px.l[0] = a.x; px.l[1] = a.y; px.l[2] = a.z; px.l[3] = a.w; px.l[4] = b.x; px.l[5] = b.y; px.l[6] = b.z; px.l[7] = b.w;
    py.l[0] = c.x; py.l[1] = c.y; py.l[2] = c.z; py.l[3] = c.w; py.l[4] = e.x; py.l[5] = e.y; py.l[6] = e.z; py.l[7] = e.w;
}

template <int T>
__device__ __noinline__ void exec_curve_part(const OpRec* r, uint4* cb) {
    const uint32_t mode = r->c[0][0], table = r->c[0][1], first = r->c[0][2], count = r->c[0][3], toff = r->c[0][4], imm = r->c[0][5];
    uint32_t l[9];
    if (mode == 0) {
        Fe v;
        hv_load<T>(v, cb, r->w[3]);
#pragma unroll
        for (int k = 0; k < 8; ++k) l[k] = v.l[k];
    } else if (mode == 1) {
        const uint32_t* ivp = g_curve_tables.pedersen + (size_t)2 * PED_WINDOWS * PED_TABLE_SIZE * 16 + (size_t)(imm % PED_IV_SIZE) * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) l[k] = ivp[k];
    } else {
#pragma unroll
        for (int k = 1; k < 8; ++k) l[k] = 0;
        l[0] = imm;
    }
    l[8] = 0;
    Jac acc;
    acc.inf = true;
#pragma unroll 1
    for (uint32_t w = first; w < first + count; ++w) {
        Fe px, py;
        if (table == 0) {
            const uint32_t d = (l[w >> 2] >> (8 * (w & 3))) & 0xFF;
            if (d == 0) continue;
            load_table_point(px, py, g_curve_tables.fixed_base, (size_t)(toff + w) * 255 + (d - 1));
        } else {
            const uint32_t o = 9 * w, limb = o >> 5, sh = o & 31;
            const uint32_t s9 = __funnelshift_r(l[limb], l[limb + 1], sh) & 0x1FF;
            load_table_point(px, py, g_curve_tables.pedersen, (size_t)(toff + w) * PED_TABLE_SIZE + s9);
        }
        jac_madd(acc, px, py);
    }
    store_jac<T>(cb, r->w[2], acc);
}

template <int T>
__device__ __noinline__ void exec_jac_add(const OpRec* r, uint4* cb) {
    Jac a, b;
    load_jac<T>(a, cb, r->w[3]);
    load_jac<T>(b, cb, r->w[4]);
    jac_add_full(a, b);
    store_jac<T>(cb, r->w[2], a);
}

// w[3] (+ w[4]) -> affine; out x = w[2], out y = w[5] (NONE: x only, the Pedersen chaining value).
// c[0][0] != 0: FixedBaseScalarMul input validation on w[6] = low, w[7] = high first (scalar_mul.rs:25-51).
template <int T>
__device__ __noinline__ void exec_jac_final(const OpRec* r, uint32_t flags, uint4* cb, unsigned long long* fail) {
    if (r->c[0][0]) {
        Fe lo, hi;
        hv_load<T>(lo, cb, r->w[6]);
        hv_load<T>(hi, cb, r->w[7]);
        bool bad = (lo.l[4] | lo.l[5] | lo.l[6] | lo.l[7]) || (hi.l[4] | hi.l[5] | hi.l[6] | hi.l[7]);
        if (!bad) {
            const uint32_t s[8] = {lo.l[0], lo.l[1], lo.l[2], lo.l[3], hi.l[0], hi.l[1], hi.l[2], hi.l[3]};
            const uint32_t n[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
            bool lt = false;
#pragma unroll
            for (int i = 7; i >= 0; --i) {
                if (s[i] != n[i]) { lt = s[i] < n[i]; break; }
            }
            bad = !lt;
        }
        if (bad) {
            hv_fail(fail, r->w[1], EK_BLACKBOX_FAILED, BB_FixedBaseScalarMul);
            return;
        }
    }
    Jac a;
    load_jac<T>(a, cb, r->w[3]);
    if (r->w[4] != 0xFFFFFFFFu) {
        Jac b;
        load_jac<T>(b, cb, r->w[4]);
        jac_add_full(a, b);
    }
    Fe x, y;
    jac_to_affine_canonical(x, y, a);
    if (r->w[5] == 0xFFFFFFFFu) hv_store<T>(cb, r->w[2], x);
    else write_point<T>(r, flags, x, y, r->w[2], r->w[5], cb, fail);
}

// ---------------------------------------------------------------------------------------------
// Value-dependent arithmetic gate: the reference's evaluate() + solve() run per lane
// (acvm/src/pwg/arithmetic.rs:27-127,212-239).  `mu` is this lane's "assigned by opcode" table for the witnesses
// whose assignment depends on instance values (entry = opcode index that assigned it, NONE = unassigned).
// payload layout: see plan.cpp general_gate().
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pl_fe(Fe& v, const uint32_t* p) {
#pragma unroll
    for (int k = 0; k < 8; ++k) v.l[k] = p[k];
}

template <int T>
__device__ __noinline__ void exec_general(const OpRec* r, uint4* cb, unsigned long long* fail, const uint32_t* payload, uint32_t* mu) {
    const uint32_t* pl = payload + r->w[7];
    const uint32_t n_mul = pl[0], n_lin = pl[1];
    Fe q;
    pl_fe(q, pl + 2);
    const uint32_t* p = pl + 10;
    uint32_t n_entries = 0, n_mulrem = 0, ew = 0xFFFFFFFFu, emu = 0xFFFFFFFFu;
    Fe ecoef;
#pragma unroll
    for (int k = 0; k < 8; ++k) ecoef.l[k] = 0;
    auto is_known = [&](uint32_t m) { return m == 0xFFFFFFFFu || mu[(size_t)m * T] != 0xFFFFFFFFu; };
#pragma unroll 1
    for (uint32_t i = 0; i < n_mul; ++i, p += 20) {
        Fe cR, cR2;
        pl_fe(cR, p);
        pl_fe(cR2, p + 8);
        const uint32_t w1 = p[16], m1 = p[17], w2 = p[18], m2 = p[19];
        const bool k1 = is_known(m1), k2 = is_known(m2);
        if (k1 && k2) {          // MulTerm::Solved: q_c += c * w1 * w2
            Fe a, b, t;
            hv_load<T>(a, cb, w1);
            hv_load<T>(b, cb, w2);
            fr::mont_mul(t, cR2, a);
            fr::mont_mul(t, t, b);
            fr::add_mod(q, q, t);
        } else if (!k1 && !k2) { // both unknown: kept as a mul term unless c == 0
            if (!fr::is_zero(cR)) ++n_mulrem;
        } else {                 // one unknown: (c * w_known, w_unknown) unless the product is zero
            Fe a, v;
            hv_load<T>(a, cb, k1 ? w1 : w2);
            fr::mont_mul(v, cR, a);
            if (!fr::is_zero(v)) {
                ++n_entries;
                ecoef = v;
                ew = k1 ? w2 : w1;
                emu = k1 ? m2 : m1;
            }
        }
    }
#pragma unroll 1
    for (uint32_t i = 0; i < n_lin; ++i, p += 18) {
        Fe cR;
        pl_fe(cR, p);
        const uint32_t w = p[16], m = p[17];
        if (is_known(m)) {
            Fe a, t;
            hv_load<T>(a, cb, w);
            fr::mont_mul(t, cR, a);
            fr::add_mod(q, q, t);
        } else {
            Fe c;
            pl_fe(c, p + 8);
            if (!fr::is_zero(c)) {
                ++n_entries;
                ecoef = c;
                ew = w;
                emu = m;
            }
        }
    }
    if (n_mulrem > 1) {          // arithmetic.rs:142 panics
        hv_fail(fail, r->w[1], EK_REFERENCE_PANIC, 0);
    } else if (n_mulrem == 1 || n_entries > 1) {
        hv_fail(fail, r->w[1], EK_TOO_MANY_UNKNOWNS, 0);
    } else if (n_entries == 0) {
        if (!fr::is_zero(q)) hv_fail(fail, r->w[1], EK_UNSATISFIED_CONSTRAIN, 0);
    } else {
        // w := -(q / coeff)     (arithmetic.rs:103-125; Div = multiply by the inverse, generic_ark.rs:375-380)
        Fe R2, one, cm, inv, qm, val;
        R2.l[0] = 0xae216da7u; R2.l[1] = 0x1bb8e645u; R2.l[2] = 0xe35c59e3u; R2.l[3] = 0x53fe3ab1u;
        R2.l[4] = 0x53bb8085u; R2.l[5] = 0x8c49833du; R2.l[6] = 0x7f4e44a5u; R2.l[7] = 0x0216d0b1u;
#pragma unroll
        for (int k = 0; k < 8; ++k) one.l[k] = 0;
        one.l[0] = 1;
        fr::mont_mul(cm, ecoef, R2);     // coeff * R
        fe_inv(inv, cm);                 // coeff^-1 * R
        fr::mont_mul(qm, q, R2);         // q * R
        fr::mont_mul(val, qm, inv);      // q/coeff * R
        fr::mont_mul(val, val, one);     // q/coeff
        if (!fr::is_zero(val)) {
            Fe zero;
#pragma unroll
            for (int k = 0; k < 8; ++k) zero.l[k] = 0;
            fr::sub_mod(val, zero, val);
        }
        hv_store<T>(cb, ew, val);
        mu[(size_t)emu * T] = r->w[1];
    }
}

template <int T>
__device__ __noinline__ void exec_require(const OpRec* r, unsigned long long* fail, const uint32_t* payload, const uint32_t* mu) {
    const uint32_t* pl = payload + r->w[7];
    const uint32_t n = pl[0];
    for (uint32_t i = 0; i < n; ++i) {
        if (mu[(size_t)pl[2 + 2 * i] * T] == 0xFFFFFFFFu) {   // blackbox/mod.rs:55-62: first missing input
            hv_fail(fail, r->w[1], EK_MISSING_ASSIGNMENT, pl[1 + 2 * i]);
            return;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Directives and memory blocks
// ---------------------------------------------------------------------------------------------
template <int T>
__device__ __forceinline__ void insert_value_dev(const OpRec* r, bool check, uint32_t slot, const Fe& v, uint4* cb, unsigned long long* fail) {
    if (check) {   // insert_value on an assigned witness: replace, then report a mismatch (pwg/mod.rs:338-357)
        Fe old;
        hv_load<T>(old, cb, slot);
        if (!fr::eq(old, v)) {
            hv_fail(fail, r->w[1], EK_UNSATISFIED_CONSTRAIN, 0);
            hv_store<T>(cb, slot, v);
        }
    } else {
        hv_store<T>(cb, slot, v);
    }
}

// Directive::ToLeRadix (directives/mod.rs:60-87): little-endian base-`radix` digits of the canonical value; missing high
// digits are 0; more digits than output witnesses => UnsatisfiedConstrain; zero decomposes to the single digit [0].
template <int T>
__device__ __noinline__ void exec_to_le_radix(const OpRec* r, uint4* cb, unsigned long long* fail, const uint32_t* payload) {
    const uint32_t* pl = payload + r->w[7];
    const uint32_t n_b = pl[0], radix = pl[1];
    const uint32_t* mask = pl + 2;
    const uint32_t* outs = mask + (n_b + 31) / 32;
    Fe v;
    hv_load<T>(v, cb, r->w[3]);
    uint32_t l[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) l[k] = v.l[k];
    if (n_b == 0) {   // [0] has length 1 > 0, and any non-zero value has >= 1 digit
        hv_fail(fail, r->w[1], EK_UNSATISFIED_CONSTRAIN, 0);
        return;
    }
    const bool pow2 = (radix & (radix - 1)) == 0;
    const int sh = 31 - __clz(radix);
#pragma unroll 1
    for (uint32_t i = 0; i < n_b; ++i) {
        uint32_t digit;
        if (pow2) {
            digit = l[0] & (radix - 1);
#pragma unroll
            for (int k = 0; k < 7; ++k) l[k] = __funnelshift_r(l[k], l[k + 1], sh);
            l[7] >>= sh;
        } else {
            unsigned long long rem = 0;
#pragma unroll
            for (int k = 7; k >= 0; --k) {
                unsigned long long cur = (rem << 32) | l[k];
                l[k] = (uint32_t)(cur / radix);
                rem = cur % radix;
            }
            digit = (uint32_t)rem;
        }
        Fe d;
#pragma unroll
        for (int k = 0; k < 8; ++k) d.l[k] = 0;
        d.l[0] = digit & 0xFF;
        insert_value_dev<T>(r, (mask[i >> 5] >> (i & 31)) & 1, outs[i], d, cb, fail);
    }
    if (l[0] | l[1] | l[2] | l[3] | l[4] | l[5] | l[6] | l[7]) hv_fail(fail, r->w[1], EK_UNSATISFIED_CONSTRAIN, 0);
}

// Directive::Quotient (directives/mod.rs:28-59): euclidean division of the canonical 256-bit integers; predicate == 0 or
// b == 0 gives q = r = 0.
template <int T>
__device__ __noinline__ void exec_quotient(const OpRec* r, uint32_t flags, uint4* cb, unsigned long long* fail) {
    Fe a, b, q, rem;
    hv_load<T>(a, cb, r->w[3]);
    hv_load<T>(b, cb, r->w[4]);
    bool pred = true;
    if (r->w[5] != 0xFFFFFFFFu) {
        Fe p;
        hv_load<T>(p, cb, r->w[5]);
        pred = !fr::is_zero(p);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) { q.l[k] = 0; rem.l[k] = 0; }
    if (pred && !fr::is_zero(b)) {
        uint32_t al[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) al[k] = a.l[k];
        const int top = (int)fr::num_bits(a) - 1;
#pragma unroll 1
        for (int i = top; i >= 0; --i) {
            // rem = (rem << 1) | bit_i(a)
#pragma unroll
            for (int k = 7; k > 0; --k) rem.l[k] = __funnelshift_l(rem.l[k - 1], rem.l[k], 1);
            rem.l[0] = (rem.l[0] << 1) | ((al[i >> 5] >> (i & 31)) & 1);
            // if rem >= b: rem -= b, set quotient bit
            uint32_t t[8], borrow;
            fr::sub_cc(t[0], rem.l[0], b.l[0]);
#pragma unroll
            for (int k = 1; k < 8; ++k) fr::subc_cc(t[k], rem.l[k], b.l[k]);
            fr::subc(borrow, 0, 0);
            if (borrow == 0) {
#pragma unroll
                for (int k = 0; k < 8; ++k) rem.l[k] = t[k];
                uint32_t ql[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) ql[k] = q.l[k];
                ql[i >> 5] |= 1u << (i & 31);
#pragma unroll
                for (int k = 0; k < 8; ++k) q.l[k] = ql[k];
            }
        }
    }
    insert_value_dev<T>(r, flags & GF_OUT_CHECK, r->w[2], q, cb, fail);
    insert_value_dev<T>(r, flags & GF_OUT2_CHECK, r->w[6], rem, cb, fail);
}

// BinaryIntOp of a Brillig opcode that the plan lowered to the device (plan.cpp brillig_symbolic), on canonical values:
// evaluate_binary_bigint_op (brillig_vm/src/arithmetic.rs:23-81) for 1 <= bit_size <= 128, every op but SignedDiv.  Operands
// are whole field values (the reference reduces mod 2^bit_size only where it says so); results are < 2^128 < p.
template <int T>
__device__ __noinline__ void exec_int_op(const OpRec* r, uint32_t flags, uint4* cb, unsigned long long* fail) {
    Fe a, b, res;
    hv_load<T>(a, cb, r->w[3]);
    hv_load<T>(b, cb, r->w[4]);
    const uint32_t op = r->w[7] & 0xFF, bs = r->w[7] >> 8;
    auto mask = [&](Fe& v) {   // v mod 2^bs
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (bs <= 32u * k) v.l[k] = 0;
            else if (bs < 32u * (k + 1)) v.l[k] &= (1u << (bs - 32u * k)) - 1u;
        }
    };
    auto geq = [&](const Fe& x, const Fe& y) {   // x >= y
        uint32_t t, borrow;
        fr::sub_cc(t, x.l[0], y.l[0]);
#pragma unroll
        for (int k = 1; k < 8; ++k) fr::subc_cc(t, x.l[k], y.l[k]);
        fr::subc(borrow, 0, 0);
        return borrow == 0;
    };
    bool panic = false;
#pragma unroll
    for (int k = 0; k < 8; ++k) res.l[k] = 0;
    switch (op) {
        case 0:   // Add: (a + b) % 2^bs  (a, b < p: no carry out of 256 bits)
            fr::add_raw(res, a, b);
            mask(res);
            break;
        case 1: {   // Sub: (2^bs + a - b) % 2^bs ; BigUint underflow panics
            Fe t;
#pragma unroll
            for (int k = 0; k < 8; ++k) t.l[k] = ((uint32_t)k == (bs >> 5)) ? (1u << (bs & 31)) : 0u;
            fr::add_raw(t, t, a);   // bs <= 128 and a < 2^254: fits
            if (!geq(t, b)) panic = true;
            fr::sub_cc(res.l[0], t.l[0], b.l[0]);
#pragma unroll
            for (int k = 1; k < 7; ++k) fr::subc_cc(res.l[k], t.l[k], b.l[k]);
            fr::subc(res.l[7], t.l[7], b.l[7]);
            mask(res);
            break;
        }
        case 2: {   // Mul: low bs <= 128 bits of a * b
            uint32_t acc[4] = {0, 0, 0, 0};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint32_t carry = 0;
#pragma unroll
                for (int j = 0; i + j < 4; ++j) {
                    const unsigned long long t = (unsigned long long)a.l[j] * b.l[i] + acc[i + j] + carry;
                    acc[i + j] = (uint32_t)t;
                    carry = (uint32_t)(t >> 32);
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) res.l[k] = acc[k];
            mask(res);
            break;
        }
        case 4: {   // UnsignedDiv: (a % 2^bs) / (b % 2^bs) ; division by zero panics
            mask(a);
            mask(b);
            if (fr::is_zero(b)) {
                panic = true;
                break;
            }
            Fe rem;
#pragma unroll
            for (int k = 0; k < 8; ++k) rem.l[k] = 0;
            const int top = (int)fr::num_bits(a) - 1;
#pragma unroll 1
            for (int i = top; i >= 0; --i) {
#pragma unroll
                for (int k = 7; k > 0; --k) rem.l[k] = __funnelshift_l(rem.l[k - 1], rem.l[k], 1);
                uint32_t limb = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k == (i >> 5)) limb = a.l[k];
                rem.l[0] = (rem.l[0] << 1) | ((limb >> (i & 31)) & 1);
                if (geq(rem, b)) {
                    fr::sub_cc(rem.l[0], rem.l[0], b.l[0]);
#pragma unroll
                    for (int k = 1; k < 7; ++k) fr::subc_cc(rem.l[k], rem.l[k], b.l[k]);
                    fr::subc(rem.l[7], rem.l[7], b.l[7]);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k == (i >> 5)) res.l[k] |= 1u << (i & 31);
                }
            }
            break;
        }
        case 5:   // Equals / LessThan / LessThanEquals on the values mod 2^bs
        case 6:
        case 7: {
            mask(a);
            mask(b);
            const bool ge = geq(a, b), le = geq(b, a);
            res.l[0] = (op == 5) ? (ge && le) : (op == 6 ? !ge : le);
            break;
        }
        case 8:
        case 9:
        case 10:
#pragma unroll
            for (int k = 0; k < 8; ++k) res.l[k] = op == 8 ? (a.l[k] & b.l[k]) : (op == 9 ? (a.l[k] | b.l[k]) : (a.l[k] ^ b.l[k]));
            mask(res);
            break;
        case 11:
        case 12: {   // Shl / Shr: the shift amount must fit u128; (a << s) % 2^bs resp. (a >> s) % 2^bs
            if (b.l[4] | b.l[5] | b.l[6] | b.l[7]) {
                panic = true;
                break;
            }
            const uint32_t s = (b.l[1] | b.l[2] | b.l[3] || b.l[0] >= 256u) ? 256u : b.l[0];
            if (s < 256u) {
                const uint32_t ws = s >> 5, bsft = s & 31;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    uint32_t lo = 0, hi = 0;   // limbs that land in position k
                    if (op == 11) {           // left: res[k] = a[k-ws] << bsft | a[k-ws-1] >> (32-bsft)
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if ((uint32_t)j + ws == (uint32_t)k) hi = a.l[j];
                            if ((uint32_t)j + ws + 1 == (uint32_t)k) lo = a.l[j];
                        }
                        res.l[k] = bsft ? ((hi << bsft) | (lo >> (32 - bsft))) : hi;
                    } else {                  // right: res[k] = a[k+ws] >> bsft | a[k+ws+1] << (32-bsft)
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if ((uint32_t)k + ws == (uint32_t)j) lo = a.l[j];
                            if ((uint32_t)k + ws + 1 == (uint32_t)j) hi = a.l[j];
                        }
                        res.l[k] = bsft ? ((lo >> bsft) | (hi << (32 - bsft))) : lo;
                    }
                }
            }
            mask(res);
            break;
        }
        default:
            panic = true;
            break;
    }
    if (panic) {
        hv_fail(fail, r->w[1], EK_REFERENCE_PANIC, 0);
        return;
    }
    insert_value_dev<T>(r, flags & GF_OUT_CHECK, r->w[2], res, cb, fail);
}

// index.try_to_u64().unwrap() as u32  (memory_op.rs:70-72): > 64 bits panics in the reference
__device__ __forceinline__ bool mem_index(const Fe& idx, uint32_t& out) {
    out = idx.l[0];
    return (idx.l[2] | idx.l[3] | idx.l[4] | idx.l[5] | idx.l[6] | idx.l[7]) == 0;
}

template <int T>
__device__ __noinline__ void exec_mem(const OpRec* r, uint32_t kind, uint4* cb, unsigned long long* fail, const uint32_t* payload) {
    const uint32_t base = payload[r->w[7]], len = payload[r->w[7] + 1];
    Fe idx;
    hv_load<T>(idx, cb, r->w[3]);
    uint32_t mi;
    if (!mem_index(idx, mi)) {
        hv_fail(fail, r->w[1], EK_REFERENCE_PANIC, 0);
        return;
    }
    bool pred = true;
    if (r->w[5] != 0xFFFFFFFFu) {
        Fe p;
        hv_load<T>(p, cb, r->w[5]);
        pred = !fr::is_zero(p);
    }
    if (r->w[6] != 0xFFFFFFFFu) {   // witness-dependent read/write selector (memory_op.rs:68,81): 0 reads, anything else writes
        Fe sel;
        hv_load<T>(sel, cb, r->w[6]);
        const bool wants_read = fr::is_zero(sel);
        if (kind == MK_MEM_READ && !wants_read) {       // get_value(value) of the unassigned witness the plan reads into
            hv_fail(fail, r->w[1], EK_MISSING_ASSIGNMENT, r->c[0][0]);
            return;
        }
        if (kind == MK_MEM_WRITE && wants_read) {       // value.to_witness() is None for an assigned value: the reference panics
            hv_fail(fail, r->w[1], EK_REFERENCE_PANIC, 0);
            return;
        }
    }
    if (kind == MK_MEM_READ) {
        Fe v;
#pragma unroll
        for (int k = 0; k < 8; ++k) v.l[k] = 0;
        if (pred) {   // a zero predicate skips the read and zeroes the output (memory_op.rs:98-104)
            if (mi >= len) {
                hv_fail(fail, r->w[1], EK_INDEX_OUT_OF_BOUNDS, mi);
                return;
            }
            hv_load<T>(v, cb, base + mi);
        }
        hv_store<T>(cb, r->w[2], v);
    } else if (pred) {
        if (mi >= len) {
            hv_fail(fail, r->w[1], EK_INDEX_OUT_OF_BOUNDS, mi);
            return;
        }
        Fe v;
        hv_load<T>(v, cb, r->w[4]);
        hv_store<T>(cb, base + mi, v);
    }
}

// ---------------------------------------------------------------------------------------------
// ECDSA secp256k1 / secp256r1 (signature/ecdsa.rs:12-97): every input is one byte = the low byte of its witness
// (signature/mod.rs:5-18, to_be_bytes().last()); out := 1 / 0; reference panics -> EK_REFERENCE_PANIC.
// ---------------------------------------------------------------------------------------------
static __device__ __noinline__ int ecdsa_verify_dev(int curve, const uint8_t* bytes) {
    return ec::ecdsa_verify(curve, bytes + 128, bytes, bytes + 32, bytes + 64);
}
template <int T>
__device__ __noinline__ void exec_ecdsa(const OpRec* r, uint32_t flags, uint4* cb, unsigned long long* fail, const uint32_t* payload) {
    const uint32_t* pl = payload + r->w[7];
    uint8_t bytes[160];   // pkx | pky | sig | hashed message
#pragma unroll 1
    for (int k0 = 0; k0 < 160; k0 += 8) {
        uint32_t low[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) low[g] = reinterpret_cast<const uint32_t*>(cb + (size_t)pl[1 + k0 + g] * (2 * T))[0];
#pragma unroll
        for (int g = 0; g < 8; ++g) bytes[k0 + g] = (uint8_t)low[g];
    }
    const int res = ecdsa_verify_dev((int)pl[0], bytes);
    if (res == ec::EC_PANIC) {
        hv_fail(fail, r->w[1], EK_REFERENCE_PANIC, 0);
        return;
    }
    Fe v;
#pragma unroll
    for (int k = 0; k < 8; ++k) v.l[k] = 0;
    v.l[0] = res == ec::EC_TRUE;
    insert_value_dev<T>(r, flags & GF_OUT_CHECK, r->w[2], v, cb, fail);
}

template <int T>
__device__ __forceinline__ void exec_heavy(const OpRec* r, uint32_t kind, uint32_t flags, uint4* cb, unsigned long long* fail,
                                           const uint32_t* payload, uint32_t* mu) {
    switch (kind) {
        // payload word 3 of a hash op: every message input is one byte -> packed-word variants (fetch_words)
        case MK_SHA256:
            if (payload[r->w[7] + 3]) exec_sha256_bytes<T>(r, cb, fail, payload + r->w[7]); else exec_sha256<T>(r, cb, fail, payload);
            break;
        case MK_KECCAK256:
            if (payload[r->w[7] + 3]) exec_keccak256_bytes<T>(r, cb, fail, payload + r->w[7]); else exec_keccak256<T>(r, cb, fail, payload);
            break;
        case MK_BLAKE2S:
        case MK_HASH_TO_FIELD:
            if (payload[r->w[7] + 3]) exec_blake2s_bytes<T>(r, kind == MK_HASH_TO_FIELD, cb, fail, payload + r->w[7]);
            else exec_blake2s<T>(r, kind == MK_HASH_TO_FIELD, cb, fail, payload);
            break;
        case MK_FIXED_BASE:
            exec_fixed_base<T>(r, flags, cb, fail);
            break;
        case MK_PEDERSEN:
            exec_pedersen<T>(r, flags, cb, fail, payload);
            break;
        case MK_ECDSA:
            exec_ecdsa<T>(r, flags, cb, fail, payload);
            break;
        case MK_CURVE_PART:
            exec_curve_part<T>(r, cb);
            break;
        case MK_JAC_ADD:
            exec_jac_add<T>(r, cb);
            break;
        case MK_JAC_FINAL:
            exec_jac_final<T>(r, flags, cb, fail);
            break;
        case MK_GATE_GENERAL:
            exec_general<T>(r, cb, fail, payload, mu);
            break;
        case MK_REQUIRE:
            exec_require<T>(r, fail, payload, mu);
            break;
        case MK_COPY: {
            Fe v;
            hv_load<T>(v, cb, r->w[3]);
            hv_store<T>(cb, r->w[2], v);
            break;
        }
        case MK_TO_LE_RADIX:
            exec_to_le_radix<T>(r, cb, fail, payload);
            break;
        case MK_QUOTIENT:
            exec_quotient<T>(r, flags, cb, fail);
            break;
        case MK_INT_OP:
            exec_int_op<T>(r, flags, cb, fail);
            break;
        case MK_HASH_PACK:
            exec_hash_pack<T>(r, cb, payload);
            break;
        case MK_HASH_CORE:
            exec_hash_core<T>(r, cb, payload);
            break;
        case MK_HASH_UNPACK:
            exec_hash_unpack<T>(r, cb, fail, payload);
            break;
        case MK_MEM_READ:
        case MK_MEM_WRITE:
            exec_mem<T>(r, kind, cb, fail, payload);
            break;
        default:
            break;
    }
}

}  // namespace acvmb
