// Host Brillig VM: "Brillig opcodes execute on the host brillig_vm with results DMA'd back into the device
// WitnessMap" (north star).  Restates brillig_vm 0.27.0:
//   VM::process_opcode            brillig_vm/src/lib.rs:153-318
//   registers (grow on write)     brillig_vm/src/registers.rs:13-42
//   memory (grow on write)        brillig_vm/src/memory.rs:17-48
//   field / bigint binary ops     brillig_vm/src/arithmetic.rs:7-96
//   blackbox ops                  brillig_vm/src/black_box.rs:42-165 (Sha256, Keccak256, FixedBaseScalarMul here)
// Values are canonical U256 field elements.  A condition under which the reference would panic!() (register past
// 2^16, pointer that does not fit usize, out-of-range memory read, bigint underflow, division by zero ...) ends the
// run with Status::Panic, which the caller reports as ACVMB_E_REFERENCE_PANIC.
#pragma once
#include "ecdsa.cuh"
#include <string>
#include <vector>

#include "acir.hpp"
#include "curve_host.hpp"
#include "fr_host.hpp"

namespace acvmb {
namespace bvm {

enum class Status { Finished, Failure, ForeignCallWait, Panic };

struct Result {
    Status status = Status::Finished;
    std::string message;
    std::vector<uint64_t> call_stack;   // Failure: call stack + failing pc (lib.rs:125-132)
    std::vector<std::vector<U256>> fc_inputs;   // ForeignCallWait: resolved inputs (lib.rs:198-202)
};

constexpr size_t MAX_REGISTERS = 1u << 16;
constexpr size_t MAX_MEMORY = 1u << 26;     // sanity bound for this implementation (the reference is bounded by RAM)
constexpr uint64_t MAX_STEPS = 1ull << 32;

struct PanicEx {
    const char* what;
};

inline const uint32_t* sha256_k() {
    static const uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
        0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
        0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
        0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
        0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
        0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    return K;
}

// K[i] + W[i] of the LAST block of a message whose length is a multiple of 64 bytes: that block is padding only
// (0x80, zeros, the bit length), so its whole message schedule is a plan-time constant (heavy_ops.cuh sha256_compress_kw)
inline void sha256_pad_block_kw(uint64_t n_bytes, uint32_t kw[64]) {
    const uint32_t* K = sha256_k();
    auto rotr = [](uint32_t x, int r) { return (x >> r) | (x << (32 - r)); };
    uint32_t w[64] = {0};
    const uint64_t bits = n_bytes * 8;
    w[0] = 0x80000000u;
    w[14] = (uint32_t)(bits >> 32);
    w[15] = (uint32_t)bits;
    for (int i = 16; i < 64; ++i) {
        uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    for (int i = 0; i < 64; ++i) kw[i] = K[i] + w[i];
}

inline void sha256_host(const uint8_t* msg, size_t n, uint8_t out[32]) {
    const uint32_t* K = sha256_k();
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    std::vector<uint8_t> m(msg, msg + n);
    m.push_back(0x80);
    while (m.size() % 64 != 56) m.push_back(0);
    uint64_t bits = (uint64_t)n * 8;
    for (int i = 7; i >= 0; --i) m.push_back((uint8_t)(bits >> (8 * i)));
    auto rotr = [](uint32_t x, int r) { return (x >> r) | (x << (32 - r)); };
    for (size_t off = 0; off < m.size(); off += 64) {
        uint32_t w[64];
        for (int i = 0; i < 16; ++i)
            w[i] = ((uint32_t)m[off + 4 * i] << 24) | ((uint32_t)m[off + 4 * i + 1] << 16) | ((uint32_t)m[off + 4 * i + 2] << 8) | m[off + 4 * i + 3];
        for (int i = 16; i < 64; ++i) {
            uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; ++i) {
            uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
            uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    for (int i = 0; i < 32; ++i) out[i] = (uint8_t)(h[i >> 2] >> (24 - 8 * (i & 3)));
}

// BLAKE2s-256, unkeyed (RFC 7693) -- blackbox_solver/src/lib.rs:52-55 -> blake2 0.10.6
inline void blake2s_host(const uint8_t* msg, size_t n, uint8_t out[32]) {
    static const uint32_t IV[8] = {0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19};
    static const uint8_t SIGMA[10][16] = {
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
        {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
        {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
        {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
        {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
    uint32_t h[8];
    for (int i = 0; i < 8; ++i) h[i] = IV[i];
    h[0] ^= 0x01010020u;   // digest length 32, no key, fanout = depth = 1
    auto rotr = [](uint32_t x, int r) { return (x >> r) | (x << (32 - r)); };
    size_t off = 0;
    do {
        const size_t take = n - off > 64 ? 64 : n - off;
        const bool last = off + take == n;
        uint8_t blk[64] = {0};
        for (size_t i = 0; i < take; ++i) blk[i] = msg[off + i];
        off += take;
        uint32_t m[16], v[16];
        for (int i = 0; i < 16; ++i)
            m[i] = (uint32_t)blk[4 * i] | ((uint32_t)blk[4 * i + 1] << 8) | ((uint32_t)blk[4 * i + 2] << 16) | ((uint32_t)blk[4 * i + 3] << 24);
        for (int i = 0; i < 8; ++i) { v[i] = h[i]; v[8 + i] = IV[i]; }
        const uint64_t t = off;
        v[12] ^= (uint32_t)t;
        v[13] ^= (uint32_t)(t >> 32);
        if (last) v[14] = ~v[14];
        auto G = [&](int a, int b, int c, int d, uint32_t x, uint32_t y) {
            v[a] = v[a] + v[b] + x; v[d] = rotr(v[d] ^ v[a], 16);
            v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 12);
            v[a] = v[a] + v[b] + y; v[d] = rotr(v[d] ^ v[a], 8);
            v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 7);
        };
        for (int r = 0; r < 10; ++r) {
            const uint8_t* sg = SIGMA[r];
            G(0, 4, 8, 12, m[sg[0]], m[sg[1]]);   G(1, 5, 9, 13, m[sg[2]], m[sg[3]]);
            G(2, 6, 10, 14, m[sg[4]], m[sg[5]]);  G(3, 7, 11, 15, m[sg[6]], m[sg[7]]);
            G(0, 5, 10, 15, m[sg[8]], m[sg[9]]);  G(1, 6, 11, 12, m[sg[10]], m[sg[11]]);
            G(2, 7, 8, 13, m[sg[12]], m[sg[13]]); G(3, 4, 9, 14, m[sg[14]], m[sg[15]]);
        }
        for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[8 + i];
    } while (off < n);
    for (int i = 0; i < 32; ++i) out[i] = (uint8_t)(h[i >> 2] >> (8 * (i & 3)));
}

// ---- 256-bit unsigned helpers (values are < 2^254, bit sizes <= 256) ----
inline U256 mask_bits(const U256& a, uint32_t bits) {
    if (bits >= 256) return a;
    U256 r;
    for (int i = 0; i < 4; ++i) {
        int lo = 64 * i;
        if ((int)bits >= lo + 64) r.l[i] = a.l[i];
        else if ((int)bits <= lo) r.l[i] = 0;
        else r.l[i] = a.l[i] & ((1ull << (bits - lo)) - 1);
    }
    return r;
}
inline U256 mul_lo(const U256& a, const U256& b) {  // low 256 bits of a*b
    U256 r;
    for (int i = 0; i < 4; ++i) {
        unsigned __int128 c = 0;
        for (int j = 0; i + j < 4; ++j) {
            c += (unsigned __int128)a.l[i] * b.l[j] + r.l[i + j];
            r.l[i + j] = (uint64_t)c;
            c >>= 64;
        }
    }
    return r;
}
inline void divmod(const U256& a, const U256& b, U256& q, U256& r) {  // b != 0
    q = U256{};
    r = U256{};
    for (int i = 255; i >= 0; --i) {
        uint64_t top = r.l[3] >> 63;
        r.l[3] = (r.l[3] << 1) | (r.l[2] >> 63);
        r.l[2] = (r.l[2] << 1) | (r.l[1] >> 63);
        r.l[1] = (r.l[1] << 1) | (r.l[0] >> 63);
        r.l[0] = (r.l[0] << 1) | ((a.l[i / 64] >> (i % 64)) & 1);
        if (top || hf::cmp(r, b) >= 0) {
            hf::sub_raw(r, r, b);
            q.l[i / 64] |= 1ull << (i % 64);
        }
    }
}
inline U256 shl(const U256& a, uint64_t s) {
    if (s >= 256) return U256{};
    U256 r;
    int w = (int)(s / 64), b = (int)(s % 64);
    for (int i = 3; i >= 0; --i) {
        uint64_t v = 0;
        if (i - w >= 0) v = a.l[i - w] << b;
        if (b && i - w - 1 >= 0) v |= a.l[i - w - 1] >> (64 - b);
        r.l[i] = v;
    }
    return r;
}
inline U256 shr(const U256& a, uint64_t s) {
    if (s >= 256) return U256{};
    U256 r;
    int w = (int)(s / 64), b = (int)(s % 64);
    for (int i = 0; i < 4; ++i) {
        uint64_t v = 0;
        if (i + w < 4) v = a.l[i + w] >> b;
        if (b && i + w + 1 < 4) v |= a.l[i + w + 1] << (64 - b);
        r.l[i] = v;
    }
    return r;
}
inline size_t to_usize(const U256& v) {  // Value::to_usize: panics unless the value fits u64
    if (v.l[1] | v.l[2] | v.l[3]) throw PanicEx{"register does not fit into u64"};
    return (size_t)v.l[0];
}

// brillig_vm/src/arithmetic.rs:23-81
inline U256 bigint_op(uint32_t op, const U256& a, const U256& b, uint32_t bs) {
    const U256 one = hf::from_u64(1);
    auto modm = [&](const U256& v) { return mask_bits(v, bs); };
    switch (op) {
        case 0: {  // Add
            U256 r;
            uint64_t c = hf::add_raw(r, a, b);
            (void)c;  // a, b < 2^254: no carry out of 256 bits
            return modm(r);
        }
        case 1: {  // Sub: (2^bs + a - b) % 2^bs ; BigUint underflow panics
            if (bs < 256) {
                U256 m = shl(one, bs), t;
                uint64_t c = hf::add_raw(t, m, a);
                if (!c && hf::cmp(t, b) < 0) throw PanicEx{"attempt to subtract with overflow"};
                U256 r;
                hf::sub_raw(r, t, b);
                return modm(r);
            }
            U256 r;
            hf::sub_raw(r, a, b);  // 2^bs + a >= b always for bs >= 256
            return modm(r);
        }
        case 2:  // Mul
            if (bs > 256) throw PanicEx{"bigint Mul with bit_size > 256 is not supported"};
            return modm(mul_lo(a, b));
        case 3: {  // SignedDiv (arithmetic.rs:83-96)
            if (bs == 0 || bs > 255) throw PanicEx{"SignedDiv bit_size out of range"};
            U256 half = shl(one, bs - 1), full = shl(one, bs);
            // to_big_signed: x if x < 2^(bs-1) else x - 2^bs  (still non-negative when x >= 2^bs)
            auto to_signed = [&](const U256& x, U256& mag) {
                if (hf::cmp(x, half) < 0) { mag = x; return false; }
                if (hf::cmp(x, full) >= 0) { hf::sub_raw(mag, x, full); return false; }
                hf::sub_raw(mag, full, x);
                return true;
            };
            U256 ma, mb;
            bool na = to_signed(a, ma), nb = to_signed(b, mb);
            if (mb.is_zero()) throw PanicEx{"attempt to divide by zero"};
            U256 q, r;
            divmod(ma, mb, q, r);   // BigInt division truncates toward zero
            if (na != nb && !q.is_zero()) {   // to_big_unsigned: 2^bs - |q| ; BigUint underflow panics
                if (hf::cmp(full, q) < 0) throw PanicEx{"attempt to subtract with overflow"};
                U256 t;
                hf::sub_raw(t, full, q);
                return t;
            }
            return q;
        }
        case 4: {  // UnsignedDiv
            U256 mb = modm(b);
            if (mb.is_zero()) throw PanicEx{"attempt to divide by zero"};
            U256 q, r;
            divmod(modm(a), mb, q, r);
            return q;
        }
        case 5: return modm(a) == modm(b) ? one : U256{};
        case 6: return hf::cmp(modm(a), modm(b)) < 0 ? one : U256{};
        case 7: return hf::cmp(modm(a), modm(b)) <= 0 ? one : U256{};
        case 8: { U256 r; for (int i = 0; i < 4; ++i) r.l[i] = a.l[i] & b.l[i]; return modm(r); }
        case 9: { U256 r; for (int i = 0; i < 4; ++i) r.l[i] = a.l[i] | b.l[i]; return modm(r); }
        case 10: { U256 r; for (int i = 0; i < 4; ++i) r.l[i] = a.l[i] ^ b.l[i]; return modm(r); }
        case 11:  // Shl
        case 12: {  // Shr
            if (bs > 128) throw PanicEx{"unsupported bit size for right shift"};
            if (b.l[2] | b.l[3]) throw PanicEx{"shift amount does not fit u128"};
            uint64_t s = (b.l[1] || b.l[0] >= 256) ? 256 : b.l[0];
            return modm(op == 11 ? shl(a, s) : shr(a, s));   // (a << s) % 2^bs == low bits; zero once s >= 256 > bs
        }
        default:
            throw PanicEx{"unknown BinaryIntOp"};
    }
}

struct VM {
    std::vector<U256> regs, mem;
    std::vector<uint64_t> call_stack;
    size_t pc = 0, fc_counter = 0;

    U256 get(uint64_t r) const {
        if (r >= MAX_REGISTERS) throw PanicEx{"Reading register past maximum!"};
        return r < regs.size() ? regs[r] : U256{};
    }
    void set(uint64_t r, const U256& v) {
        if (r >= MAX_REGISTERS) throw PanicEx{"Writing register past maximum!"};
        if (regs.size() <= r) regs.resize(r + 1);
        regs[r] = v;
    }
    const U256& mread(size_t p) const {
        if (p >= mem.size()) throw PanicEx{"memory read out of bounds"};
        return mem[p];
    }
    void mwrite(size_t p, const U256* v, size_t n) {
        if (p + n > MAX_MEMORY) throw PanicEx{"memory write beyond the supported size"};
        if (mem.size() < p + n) mem.resize(p + n);
        for (size_t i = 0; i < n; ++i) mem[p + i] = v[i];
    }
    std::vector<uint8_t> bytes_of(size_t p, size_t n) const {  // to_u8_vec: last byte of every value
        std::vector<uint8_t> out(n);
        for (size_t i = 0; i < n; ++i) out[i] = (uint8_t)mread(p + i).l[0];
        return out;
    }

    Result fail(const std::string& msg) {
        Result r;
        r.status = Status::Failure;
        r.message = msg;
        r.call_stack = call_stack;
        r.call_stack.push_back(pc);
        return r;
    }

    Result run(const Brillig& b) {
        const auto& code = b.bytecode;
        try {
            for (uint64_t steps = 0;; ++steps) {
                if (steps > MAX_STEPS) throw PanicEx{"brillig step limit exceeded"};
                if (pc >= code.size()) {
                    // set_program_counter marks Finished once pc leaves the bytecode; an EMPTY bytecode indexes [0] and panics
                    if (code.empty() || steps == 0) throw PanicEx{"program counter outside the bytecode"};
                    return Result{};
                }
                const BrilligOp& o = code[pc];
                switch (o.tag) {
                    case 0: {  // BinaryFieldOp (arithmetic.rs:7-20)
                        U256 a = get(o.r1), c = get(o.r2), r;
                        switch (o.bop) {
                            case 0: r = hf::add(a, c); break;
                            case 1: r = hf::sub(a, c); break;
                            case 2: r = hf::mul(a, c); break;
                            case 3: r = hf::mul(a, hf::inverse(c)); break;
                            case 4: r = (a == c) ? hf::from_u64(1) : U256{}; break;
                            default: throw PanicEx{"unknown BinaryFieldOp"};
                        }
                        set(o.r0, r);
                        ++pc;
                        break;
                    }
                    case 1: {  // BinaryIntOp -> from_be_bytes_reduce of the big-int result (lib.rs:371-389)
                        U256 r = bigint_op(o.bop, get(o.r1), get(o.r2), o.bit_size);
                        set(o.r0, hf::reduce(r));
                        ++pc;
                        break;
                    }
                    case 2:  // JumpIfNot
                        if (get(o.r0).is_zero()) pc = (size_t)o.r1; else ++pc;
                        break;
                    case 3:  // JumpIf
                        if (!get(o.r0).is_zero()) pc = (size_t)o.r1; else ++pc;
                        break;
                    case 4:
                        pc = (size_t)o.r0;
                        break;
                    case 5:  // Call
                        call_stack.push_back(pc);
                        pc = (size_t)o.r0;
                        break;
                    case 6:
                        set(o.r0, o.value);
                        ++pc;
                        break;
                    case 7:  // Return
                        if (call_stack.empty()) return fail("return opcode hit, but callstack already empty");
                        pc = (size_t)call_stack.back() + 1;
                        call_stack.pop_back();
                        break;
                    case 8: {  // ForeignCall (lib.rs:189-263)
                        if (fc_counter >= b.foreign_call_results.size()) {
                            Result r;
                            r.status = Status::ForeignCallWait;
                            r.message = o.function;
                            for (const RegOrMem& in : o.inputs) {   // get_register_value_or_memory_values (lib.rs:334-354)
                                std::vector<U256> vals;
                                if (in.kind == 0) {
                                    vals.push_back(get(in.a));
                                } else {
                                    size_t p = to_usize(get(in.a));
                                    size_t n = in.kind == 1 ? (size_t)in.b : to_usize(get(in.b));
                                    if (p + n > mem.size()) throw PanicEx{"memory read out of bounds"};
                                    vals.assign(mem.begin() + p, mem.begin() + p + n);
                                }
                                r.fc_inputs.push_back(std::move(vals));
                            }
                            return r;
                        }
                        const auto& values = b.foreign_call_results[fc_counter];
                        bool invalid = false;
                        size_t n = std::min(o.destinations.size(), values.size());
                        for (size_t i = 0; i < n; ++i) {
                            const RegOrMem& d = o.destinations[i];
                            const ForeignCallOutput& out = values[i];
                            if (d.kind == 0) {
                                if (out.is_array) throw PanicEx{"Function result size does not match brillig bytecode (expected 1 result)"};
                                set(d.a, out.values[0]);
                            } else if (d.kind == 1) {
                                if (!out.is_array) throw PanicEx{"Function result size does not match brillig bytecode size"};
                                if (out.values.size() != d.b) {
                                    invalid = true;
                                    break;
                                }
                                mwrite(to_usize(get(d.a)), out.values.data(), out.values.size());
                            } else {
                                if (!out.is_array) throw PanicEx{"Function result size does not match brillig bytecode size"};
                                set(d.b, hf::from_u64(out.values.size()));
                                mwrite(to_usize(get(d.a)), out.values.data(), out.values.size());
                            }
                        }
                        // the reference records a failure status here but keeps executing; the LAST status wins
                        Result pending;
                        bool has_pending = false;
                        if (o.destinations.size() != values.size()) {
                            pending = fail(std::to_string(values.size()) + " output values were provided as a foreign call result for " +
                                           std::to_string(o.destinations.size()) + " destination slots");
                            has_pending = true;
                        }
                        if (invalid) {
                            pending = fail("Function result size does not match brillig bytecode");
                            has_pending = true;
                        }
                        ++fc_counter;
                        ++pc;
                        if (has_pending && pc >= code.size()) return Result{};  // Finished overwrites the failure status
                        if (has_pending) {
                            // process_opcodes() stops on the Failure status returned by this opcode (lib.rs:136-142)
                            return pending;
                        }
                        break;
                    }
                    case 9:
                        set(o.r0, get(o.r1));
                        ++pc;
                        break;
                    case 10:  // Load
                        set(o.r0, mread(to_usize(get(o.r1))));
                        ++pc;
                        break;
                    case 11: {  // Store
                        U256 v = get(o.r1);
                        mwrite(to_usize(get(o.r0)), &v, 1);
                        ++pc;
                        break;
                    }
                    case 12: {  // BlackBox (black_box.rs:42-165)
                        if (o.bb_tag <= 3) {   // Sha256 / Blake2s / Keccak256 {message: HeapVector, output: HeapArray}; HashToField {.., output: register}
                            size_t p = to_usize(get(o.bb[0])), n = to_usize(get(o.bb[1]));
                            if (p + n > mem.size()) throw PanicEx{"memory read out of bounds"};
                            std::vector<uint8_t> msg = bytes_of(p, n);
                            uint8_t d[32];
                            if (o.bb_tag == 0) sha256_host(msg.data(), msg.size(), d);
                            else if (o.bb_tag == 2) gk::keccak256_host(msg.data(), msg.size(), d);
                            else blake2s_host(msg.data(), msg.size(), d);
                            if (o.bb_tag == 3) {   // digest as a big-endian integer, reduced mod p (blackbox_solver/src/lib.rs:62-65,94-99)
                                set(o.bb[2], hf::from_be_bytes_reduce(d, 32));
                            } else {
                                U256 vals[32];
                                for (int i = 0; i < 32; ++i) vals[i] = hf::from_u64(d[i]);
                                mwrite(to_usize(get(o.bb[2])), vals, 32);
                            }
                        } else if (o.bb_tag == 4 || o.bb_tag == 5) {   // EcdsaSecp256k1 / r1 (black_box.rs:73-127)
                            static const char* what[3] = {"Invalid public key x length", "Invalid public key y length", "Invalid signature length"};
                            static const size_t want[3] = {32, 32, 64};
                            std::vector<uint8_t> part[3];
                            for (int k = 0; k < 3; ++k) {
                                size_t p = to_usize(get(o.bb[2 + 2 * k])), n = (size_t)o.bb[3 + 2 * k];
                                if (p + n > mem.size()) throw PanicEx{"memory read out of bounds"};
                                if (n != want[k]) return fail(std::string("Failed to solve blackbox function: ecdsa, reason: ") + what[k]);
                                part[k] = bytes_of(p, n);
                            }
                            size_t p = to_usize(get(o.bb[0])), n = to_usize(get(o.bb[1]));
                            if (p + n > mem.size()) throw PanicEx{"memory read out of bounds"};
                            std::vector<uint8_t> hashed = bytes_of(p, n);
                            if (hashed.size() != 32) throw PanicEx{"GenericArray::from_slice: hashed message is not 32 bytes"};
                            int res = ec::ecdsa_verify(o.bb_tag == 4 ? 0 : 1, hashed.data(), part[0].data(), part[1].data(), part[2].data());
                            if (res == ec::EC_PANIC) throw PanicEx{"ecdsa verification panics in the reference"};
                            set(o.bb[8], hf::from_u64(res == ec::EC_TRUE));
                        } else if (o.bb_tag == 8) {               // FixedBaseScalarMul {low, high, result: HeapArray}
                            U256 lo = get(o.bb[0]), hi = get(o.bb[1]);
                            if (lo.l[2] | lo.l[3] | hi.l[2] | hi.l[3])
                                return fail("Failed to solve blackbox function: fixed_base_scalar_mul, reason: limb is not less than 2^128");
                            U256 s;
                            s.l[0] = lo.l[0]; s.l[1] = lo.l[1]; s.l[2] = hi.l[0]; s.l[3] = hi.l[1];
                            static const U256 n = {{0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}};
                            if (hf::cmp(s, n) >= 0)
                                return fail("Failed to solve blackbox function: fixed_base_scalar_mul, reason: not a valid grumpkin scalar");
                            gk::Pt pt = gk::mul(s, gk::generator());
                            U256 xy[2];
                            if (!pt.inf) { xy[0] = pt.x; xy[1] = pt.y; }
                            mwrite(to_usize(get(o.bb[2])), xy, 2);
                        } else {
                            throw PanicEx{"brillig blackbox op not supported by the host VM"};
                        }
                        ++pc;
                        break;
                    }
                    case 13:
                        return fail("explicit trap hit in brillig");
                    case 14:
                        return Result{};
                    default:
                        throw PanicEx{"unknown brillig opcode"};
                }
            }
        } catch (const PanicEx& e) {
            Result r;
            r.status = Status::Panic;
            r.message = e.what;
            return r;
        }
    }
};

}  // namespace bvm
}  // namespace acvmb
