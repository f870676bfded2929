// Step-VM kernels for sm_100a: one CTA = one tile of T witness instances, S micro-op slots per step.
//
// Replaces the reference's serial interpreter loop (acvm/src/pwg/mod.rs:236-303) and the opcode
// solvers it dispatches to (arithmetic.rs:27-239, blackbox/logic.rs, blackbox/range.rs, ...).
//
//  * thread (slot, lane) executes micro-op `slot` of the current step for instance `tile*T + lane`;
//    lanes of the same slot read the SAME record -> shared-memory broadcast, zero divergence.
//  * the record stream is shared by every CTA: it is staged global -> shared with TMA bulk copies
//    (cp.async.bulk + mbarrier complete_tx) through a 4-deep ring, one elected thread issuing.
//  * witness columns are tile-major in HBM: each operand load is T*16 contiguous bytes per plane.
//  * one __syncthreads() per step orders the CTA's own global stores/loads (the plan guarantees a
//    slot never reads a column written in the same step).
#include <cuda_runtime.h>
#include <stdint.h>

#include "fr.cuh"
#include "plan.hpp"
#include "vm_kernel.cuh"
#ifdef ACVMB_HEAVY_OPS
#include "heavy_ops.cuh"
#endif

namespace acvmb {

using fr::Fe;

constexpr int MAX_NSTAGE = 8;   // depth of the TMA staging ring is a launch parameter (VmArgs::n_stage)

// ---------------------------------------------------------------------------------------------
// TMA bulk copy + mbarrier (inline PTX; SASS: UBLKCP / SYNCS)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}

// ---------------------------------------------------------------------------------------------
// column access
// ---------------------------------------------------------------------------------------------
template <int T>
__device__ __forceinline__ void load_w(Fe& v, const uint4* cb, uint32_t w) {
    const uint4* p = cb + (size_t)w * (2 * T);
    uint4 lo = p[0], hi = p[T];
    v.l[0] = lo.x; v.l[1] = lo.y; v.l[2] = lo.z; v.l[3] = lo.w;
    v.l[4] = hi.x; v.l[5] = hi.y; v.l[6] = hi.z; v.l[7] = hi.w;
}
template <int T>
__device__ __forceinline__ void store_w(uint4* cb, uint32_t w, const Fe& v) {
    uint4* p = cb + (size_t)w * (2 * T);
    p[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    p[T] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ void lds_fe(Fe& v, const uint32_t* c) {
    const uint4* p = reinterpret_cast<const uint4*>(c);
    uint4 lo = p[0], hi = p[1];
    v.l[0] = lo.x; v.l[1] = lo.y; v.l[2] = lo.z; v.l[3] = lo.w;
    v.l[4] = hi.x; v.l[5] = hi.y; v.l[6] = hi.z; v.l[7] = hi.w;
}

__device__ __forceinline__ void record_fail(unsigned long long* fail, uint32_t opcode, uint32_t kind, uint32_t aux) {
    unsigned long long key = ((unsigned long long)opcode << 32) | ((unsigned long long)(kind & 0xF) << 28) | (aux & 0x0FFFFFFFu);
    atomicMin(fail, key);
}

// b-limb source for the gate dot products: plan constants, fetched limb by limb from shared memory
struct SmemLimbs3 {
    const uint32_t* c0;
    const uint32_t* c1;
    const uint32_t* c2;
    __device__ __forceinline__ uint32_t operator()(int k, int i) const { return k == 0 ? c0[i] : (k == 1 ? c1[i] : c2[i]); }
};
struct RegLimbs {
    const Fe& b;
    __device__ __forceinline__ uint32_t operator()(int, int i) const { return b.l[i]; }
};

// GF_MUL : out = cM*(x+alpha)*(y+beta) + c1*w1 + gamma   -- two Montgomery reductions: u = (x+alpha)(y+beta)/R,
//          then <u, w1> . <cM*R^2, c1*R> / R in ONE interleaved reduction (fr::mont_dot_fn).
// linear : out = cY*y + c1*w1 + c2*w2 + cC               -- one reduction for up to three products.
template <int T, int SPLIT>
__device__ __forceinline__ void exec_gate(const OpRec* r, uint32_t kind, uint32_t flags, uint4* cb, unsigned long long* fail) {
    Fe res;
    if (flags & GF_Y) {
        const uint32_t nlin = (flags >> GF_NLIN_SHIFT) & 3;
        if (flags & GF_MUL) {
            Fe x, y, u, t;
            load_w<T>(x, cb, r->w[3]);
            load_w<T>(y, cb, r->w[4]);
            // lazy reduction: x+alpha, y+beta < 2p stay unreduced (4p^2/R + p < 1.76p), and so does u
            // ((1.76 + 1) p^2 / R + p < 1.52p for the second product): one conditional subtraction per gate instead of three
            lds_fe(t, r->c[1]);
            fr::add_raw(x, x, t);
            lds_fe(t, r->c[2]);
            fr::add_raw(y, y, t);
            const Fe* a1[1] = {&x};
            fr::mont_dot_fn<1, RegLimbs, SPLIT>(u, a1, RegLimbs{y});   // (x+alpha)(y+beta)/R < 1.76p
            if (nlin == 0) {
                const Fe* a[1] = {&u};
                fr::mont_dot_fn<1, SmemLimbs3, SPLIT>(res, a, SmemLimbs3{r->c[0], nullptr, nullptr});
            } else {
                Fe w1;
                load_w<T>(w1, cb, r->w[5]);
                const Fe* a[2] = {&u, &w1};
                fr::mont_dot_fn<2, SmemLimbs3, SPLIT>(res, a, SmemLimbs3{r->c[0], r->c[3], nullptr});
            }
        } else if (flags & GF_ADDSUB) {
            // coefficients are all +-1: out = +-y +-w1 +-w2 + cC with modular additions only
            Fe y, t;
            load_w<T>(y, cb, r->w[4]);
            lds_fe(res, r->c[4]);
            if (flags & GF_NEG_Y) fr::sub_mod(res, res, y); else fr::add_mod(res, res, y);
            if (nlin >= 1) {
                load_w<T>(t, cb, r->w[5]);
                if (flags & GF_NEG_W1) fr::sub_mod(res, res, t); else fr::add_mod(res, res, t);
            }
            if (nlin >= 2) {
                load_w<T>(t, cb, r->w[6]);
                if (flags & GF_NEG_W2) fr::sub_mod(res, res, t); else fr::add_mod(res, res, t);
            }
        } else {
            Fe y;
            load_w<T>(y, cb, r->w[4]);
            if (nlin == 0) {
                const Fe* a[1] = {&y};
                fr::mont_dot_fn<1, SmemLimbs3, SPLIT>(res, a, SmemLimbs3{r->c[1], nullptr, nullptr});
            } else {
                Fe w1;
                load_w<T>(w1, cb, r->w[5]);
                if (nlin == 1) {
                    const Fe* a[2] = {&y, &w1};
                    fr::mont_dot_fn<2, SmemLimbs3, SPLIT>(res, a, SmemLimbs3{r->c[1], r->c[2], nullptr});
                } else {
                    Fe w2;
                    load_w<T>(w2, cb, r->w[6]);
                    const Fe* a[3] = {&y, &w1, &w2};
                    fr::mont_dot_fn<3, SmemLimbs3, SPLIT>(res, a, SmemLimbs3{r->c[1], r->c[2], r->c[3]});
                }
            }
        }
        if (!(flags & GF_ADDSUB)) {
            fr::cond_sub_p(res);
            Fe cC;
            lds_fe(cC, r->c[4]);
            fr::add_raw(res, res, cC);
            fr::cond_sub_p(res);
        }
    } else {
        lds_fe(res, r->c[4]);
    }
    if (kind == MK_GATE_ASSIGN) {
        if (flags & GF_OUT_CHECK) {
            Fe old;
            load_w<T>(old, cb, r->w[2]);
            if (!fr::eq(old, res)) {  // insert_value replaces the old value before it reports the mismatch (mod.rs:343-354)
                record_fail(fail, r->w[1], EK_UNSATISFIED_CONSTRAIN, 0);
                store_w<T>(cb, r->w[2], res);
            }
        } else {
            store_w<T>(cb, r->w[2], res);
        }
    } else {
        if (!fr::is_zero(res)) record_fail(fail, r->w[1], EK_UNSATISFIED_CONSTRAIN, 0);
    }
}

// AND / XOR on the low `nb` bits of the canonical values (acir_field/src/generic_ark.rs:322-354,446-473)
template <int T>
__device__ __forceinline__ void exec_logic(const OpRec* r, uint32_t kind, uint32_t flags, uint4* cb, unsigned long long* fail) {
    Fe x, y, res;
    load_w<T>(x, cb, r->w[3]);
    load_w<T>(y, cb, r->w[4]);
    const uint32_t nb = r->w[7];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint32_t m;
        if (nb >= 32u * (i + 1)) m = 0xFFFFFFFFu;
        else if (nb <= 32u * i) m = 0u;
        else m = (1u << (nb - 32u * i)) - 1u;
        uint32_t a = x.l[i] & m, b = y.l[i] & m;
        res.l[i] = (kind == MK_AND) ? (a & b) : (a ^ b);
    }
    if (nb >= 254) fr::reduce_256(res);
    if (flags & GF_OUT_CHECK) {
        Fe old;
        load_w<T>(old, cb, r->w[2]);
        if (!fr::eq(old, res)) {  // insert_value replaces the old value before it reports the mismatch (mod.rs:343-354)
                record_fail(fail, r->w[1], EK_UNSATISFIED_CONSTRAIN, 0);
                store_w<T>(cb, r->w[2], res);
            }
    } else {
        store_w<T>(cb, r->w[2], res);
    }
}

template <int T>
__device__ __forceinline__ void exec_range(const OpRec* r, uint4* cb, unsigned long long* fail) {
    Fe x;
    load_w<T>(x, cb, r->w[3]);
    if (fr::num_bits(x) > r->w[7]) record_fail(fail, r->w[1], EK_UNSATISFIED_CONSTRAIN, 0);
}

// FULL variant: the curve / hash code wants ~128 registers, which would allow only 4 CTAs of 128 threads per SM; the
// plan serialises heavy ops along each instance's dependency chain, so run time is (number of CTA waves) x (sum of heavy-op
// latencies) and keeping the whole sub-batch in ONE wave matters more than spill-free heavy ops: cap at 7 CTAs/SM.
// CAP selects the register-capped build; the launcher uses it only when the uncapped one could not hold the sub-batch in a
// single wave (it costs ~30 % on Keccak, whose state then spills).
template <int T, int S, bool FULL, int SPLIT, bool CAP = false>
__global__ void __launch_bounds__(T* S, (CAP && T * S <= 128) ? (896 / (T * S)) : 1) vm_kernel(const VmArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t chunk_bytes = a.chunk_steps * S * (uint32_t)sizeof(OpRec);
    const uint32_t NSTAGE = a.n_stage;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)NSTAGE * chunk_bytes);

    const uint32_t tid = threadIdx.x;
    const uint32_t slot = tid / T;
    const uint32_t lane = tid % T;
    const uint32_t tile = blockIdx.x;
    uint4* cb = a.cols + (size_t)tile * a.n_slots * (2 * T) + lane;
    unsigned long long* fail = a.fail + (size_t)tile * T + lane;
    uint32_t* mu = a.mu_assign + (size_t)tile * a.n_mu * T + lane;

    const uint32_t n_chunks = a.n_steps / a.chunk_steps;
    const uint8_t* stream = a.stream + (size_t)a.first_step * S * sizeof(OpRec);
    if (tid == 0) {
        for (uint32_t s = 0; s < NSTAGE; ++s) mbar_init(&bars[s], 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t pre = n_chunks < NSTAGE ? n_chunks : NSTAGE;
        for (uint32_t c = 0; c < pre; ++c) {
            mbar_expect_tx(&bars[c], chunk_bytes);
            tma_bulk_g2s(smem + (size_t)c * chunk_bytes, stream + (size_t)c * chunk_bytes, chunk_bytes, &bars[c]);
        }
    }

    for (uint32_t c = 0; c < n_chunks; ++c) {
        const uint32_t st = c % NSTAGE;
        mbar_wait(&bars[st], (c / NSTAGE) & 1);
        const OpRec* recs = reinterpret_cast<const OpRec*>(smem + (size_t)st * chunk_bytes);
        for (uint32_t s = 0; s < a.chunk_steps; ++s) {
            const OpRec* r = recs + s * S + slot;
            const uint32_t hdr = r->w[0];
            const uint32_t kind = hdr & 0xFF, flags = hdr >> 8;
            switch (kind) {
                case MK_NOP:
                    break;
                case MK_GATE_ASSIGN:
                case MK_GATE_CHECK:
                    exec_gate<T, SPLIT>(r, kind, flags, cb, fail);
                    break;
                case MK_AND:
                case MK_XOR:
                    exec_logic<T>(r, kind, flags, cb, fail);
                    break;
                case MK_RANGE:
                    exec_range<T>(r, cb, fail);
                    break;
                default:
#ifdef ACVMB_HEAVY_OPS
                    if constexpr (FULL) exec_heavy<T>(r, kind, flags, cb, fail, a.payload, mu);
#endif
                    break;
            }
            __syncthreads();
        }
        // every thread is past the last read of this stage: refill it
        if (tid == 0 && c + NSTAGE < n_chunks) {
            mbar_expect_tx(&bars[st], chunk_bytes);
            tma_bulk_g2s(smem + (size_t)st * chunk_bytes, stream + (size_t)(c + NSTAGE) * chunk_bytes, chunk_bytes, &bars[st]);
        }
    }
}

template <int T, int S, bool FULL, int SPLIT = FR_ALU_SPLIT>
static cudaError_t launch_one(const VmArgs& args, cudaStream_t stream) {
    if (args.n_stage < 1 || args.n_stage > MAX_NSTAGE) return cudaErrorInvalidValue;
    size_t smem = (size_t)args.n_stage * args.chunk_steps * S * sizeof(OpRec) + args.n_stage * sizeof(uint64_t);
    auto k = vm_kernel<T, S, FULL, SPLIT, false>;
    if constexpr (FULL && T * S <= 128) {
        // uncapped FULL build: ~128 registers -> 65536 / (128 * threads) CTAs per SM
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
        const uint32_t one_wave = (uint32_t)sms * (65536u / (128u * T * S));
        if (args.n_tiles > one_wave) k = vm_kernel<T, S, FULL, SPLIT, true>;
    }
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<args.n_tiles, T * S, smem, stream>>>(args);
    return cudaGetLastError();
}

// tile shapes (T instances x S slots).  The FULL variant (hash / curve / general ops) is instantiated for fewer shapes:
// it dominates compile time.
#define ACVMB_CONFIGS_LIGHT(X) X(8, 16) X(4, 16) X(2, 16) X(16, 16) X(16, 8) X(32, 4) X(32, 2) X(32, 1) X(4, 32) X(2, 64)
#define ACVMB_CONFIGS_FULL(X) X(8, 16) X(4, 16) X(16, 8) X(32, 4) X(32, 2) X(32, 1) X(4, 32) X(2, 64)

cudaError_t set_curve_tables(const uint32_t* fixed_base, const uint32_t* pedersen) {
#ifdef ACVMB_HEAVY_OPS
    CurveTables t{fixed_base, pedersen};
    return cudaMemcpyToSymbol(g_curve_tables, &t, sizeof(t));
#else
    (void)fixed_base; (void)pedersen;
    return cudaSuccess;
#endif
}

bool vm_config_supported(int T, int S, bool full) {
#define X(t, s) if (T == t && S == s) return true;
    if (full) { ACVMB_CONFIGS_FULL(X) } else { ACVMB_CONFIGS_LIGHT(X) }
#undef X
    return false;
}

cudaError_t launch_vm(const KernelConfig& cfg, const VmArgs& args, cudaStream_t stream) {
    // tuning hook: explicit FMA/ALU pipe-split level for the production tile shape of the arithmetic kernel
    if (cfg.split >= 0 && !cfg.full && cfg.S == 16 && cfg.T == 8) {
#define Y(sp) if (cfg.split == sp) return launch_one<8, 16, false, sp>(args, stream);
        Y(0) Y(1) Y(2) Y(3) Y(4)
#undef Y
    }
    if (cfg.full) {
#define X(t, s) if (cfg.T == t && cfg.S == s) return launch_one<t, s, true>(args, stream);
        ACVMB_CONFIGS_FULL(X)
#undef X
    } else {
#define X(t, s) if (cfg.T == t && cfg.S == s) return launch_one<t, s, false>(args, stream);
        ACVMB_CONFIGS_LIGHT(X)
#undef X
    }
    return cudaErrorInvalidConfiguration;
}

// ---------------------------------------------------------------------------------------------
// input scatter: [inst][k][32 B big-endian]  ->  canonical LE limbs in column input_slots[k]
// (FieldElement::from_be_bytes_reduce semantics, acir_field/src/generic_ark.rs:281-283)
// ---------------------------------------------------------------------------------------------
__global__ void scatter_inputs_kernel(const uint8_t* __restrict__ in_be, const uint32_t* __restrict__ input_slots,
                                      uint32_t n_inputs, uint4* cols, uint32_t n_slots, int T, uint32_t n_inst) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)n_inst * n_inputs) return;
    uint32_t inst = (uint32_t)(gid / n_inputs), k = (uint32_t)(gid % n_inputs);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(in_be + gid * 32);
    Fe v;
#pragma unroll
    for (int m = 0; m < 8; ++m) v.l[7 - m] = __byte_perm(src[m], 0, 0x0123);
    fr::reduce_256(v);
    uint32_t tile = inst / T, lane = inst % T;
    uint4* p = cols + ((size_t)tile * n_slots + input_slots[k]) * (2 * T) + lane;
    p[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    p[T] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

cudaError_t launch_scatter_inputs(const uint8_t* in_be, const uint32_t* input_slots, uint32_t n_inputs, uint4* cols,
                                  uint32_t n_slots, int T, uint32_t n_inst, cudaStream_t stream) {
    size_t n = (size_t)n_inst * n_inputs;
    if (n == 0) return cudaSuccess;
    scatter_inputs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(in_be, input_slots, n_inputs, cols, n_slots, T, n_inst);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// output gather: columns -> [inst][n_out][32 B big-endian]; a witness that the instance never
// assigned (it failed earlier, or nothing assigns it) is written as zeros.
// ---------------------------------------------------------------------------------------------
__global__ void gather_outputs_kernel(const GatherArgs g) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)g.n_inst * g.n_out) return;
    uint32_t li = (uint32_t)(gid / g.n_out), k = (uint32_t)(gid % g.n_out);
    uint32_t inst = g.first_inst + li;
    uint32_t w = g.witness_ids ? g.witness_ids[k] : k;
    uint32_t fail_op = (uint32_t)(g.fail[inst] >> 32);
    if (g.static_fail_opcode < fail_op) fail_op = g.static_fail_opcode;
    uint32_t tile = inst / g.T, lane = inst % g.T;
    bool present = true;
    if (!g.raw) {
        uint32_t ao = g.assign_opcode[w];
        if (ao == 0xFFFFFFFDu)   // value-dependent: this lane's own record of which opcode assigned it
            ao = g.mu_assign[((size_t)tile * g.n_mu + g.mu_index_of[w]) * g.T + lane];
        present = (ao == 0xFFFFFFFEu) || (ao != 0xFFFFFFFFu && ao < fail_op);
    }
    if (g.out_present) g.out_present[gid] = present ? 1 : 0;
    if (!g.out_be) return;
    uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
    if (present) {
        const uint4* p = g.cols + ((size_t)tile * g.n_slots + w) * (2 * g.T) + lane;
        lo = p[0];
        hi = p[g.T];
    }
    uint4* dst = reinterpret_cast<uint4*>(g.out_be + gid * 32);
    dst[0] = make_uint4(__byte_perm(hi.w, 0, 0x0123), __byte_perm(hi.z, 0, 0x0123), __byte_perm(hi.y, 0, 0x0123), __byte_perm(hi.x, 0, 0x0123));
    dst[1] = make_uint4(__byte_perm(lo.w, 0, 0x0123), __byte_perm(lo.z, 0, 0x0123), __byte_perm(lo.y, 0, 0x0123), __byte_perm(lo.x, 0, 0x0123));
}

cudaError_t launch_gather_outputs(const GatherArgs& g, cudaStream_t stream) {
    size_t n = (size_t)g.n_inst * g.n_out;
    if (n == 0) return cudaSuccess;
    gather_outputs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(g);
    return cudaGetLastError();
}

// per-instance checksum of the solved witness map: sum over present witnesses of an FNV-style mix of (index, limbs)
__device__ __forceinline__ unsigned long long mix_witness(uint32_t w, const uint4& lo, const uint4& hi) {
    unsigned long long h = ((unsigned long long)w + 1ull) * 0x9E3779B97F4A7C15ull;
    const uint32_t l[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) h = (h ^ l[k]) * 0x100000001B3ull;
    return h;
}

__global__ void checksum_kernel(const GatherArgs g, unsigned long long* out) {
    // one CTA per instance, threads stride over the witnesses
    const uint32_t inst = blockIdx.x;
    uint32_t fail_op = (uint32_t)(g.fail[inst] >> 32);
    if (g.static_fail_opcode < fail_op) fail_op = g.static_fail_opcode;
    const uint32_t tile = inst / g.T, lane = inst % g.T;
    unsigned long long acc = 0;
    for (uint32_t w = threadIdx.x; w < g.n_out; w += blockDim.x) {
        uint32_t ao = g.assign_opcode[w];
        if (ao == 0xFFFFFFFDu) ao = g.mu_assign[((size_t)tile * g.n_mu + g.mu_index_of[w]) * g.T + lane];
        if (!((ao == 0xFFFFFFFEu) || (ao != 0xFFFFFFFFu && ao < fail_op))) continue;
        const uint4* p = g.cols + ((size_t)tile * g.n_slots + w) * (2 * g.T) + lane;
        acc += mix_witness(w, p[0], p[g.T]);
    }
    __shared__ unsigned long long red[256];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[inst] = red[0];
}

cudaError_t launch_checksum(const GatherArgs& g, unsigned long long* out, cudaStream_t stream) {
    if (g.n_inst == 0) return cudaSuccess;
    checksum_kernel<<<g.n_inst, 256, 0, stream>>>(g, out);
    return cudaGetLastError();
}

__global__ void fill_u64_kernel(unsigned long long* p, size_t n, unsigned long long v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
cudaError_t launch_fill_u64(unsigned long long* p, size_t n, unsigned long long v, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    fill_u64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p, n, v);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// IMAD roofline micro-benchmark ("IMAD_peak must be measured on the box").
// Each thread runs 8 independent chains; variant 0: 32-bit IMAD, 1: IMAD.WIDE.U32 (64-bit
// accumulate), 2: IMAD.WIDE.U32 with the carry-in/out pattern the Montgomery rows use.
// ---------------------------------------------------------------------------------------------
template <int VARIANT>
__global__ void __launch_bounds__(256) imad_bench_kernel(uint32_t* out, uint32_t seed, int iters) {
    uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
    if (VARIANT == 0) {
        uint32_t acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = a + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = acc[i] * a + b;
            }
        }
        uint32_t s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s ^= acc[i];
        if (s == 0x12345678u) out[0] = s;
    } else if (VARIANT == 1) {
        unsigned long long acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = a + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a + i), "r"(b));
            }
        }
        unsigned long long s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s ^= acc[i];
        if (s == 0x12345678ull) out[0] = (uint32_t)s;
    } else {
        uint32_t e[8], o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { e[i] = a + i; o[i] = b + i; }
        uint32_t av[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) av[i] = a * (i + 1);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                fr::cmad_n(o, av + 1 - 1, b + u);   // 4 wide IMADs in one carry chain
                fr::cmad_n(e, av, a + u);           // 4 more in an independent chain
                fr::cmad_n(o, av, a ^ u);
                fr::cmad_n(e, av, b ^ u);
            }
        }
        uint32_t s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s ^= e[i] ^ o[i];
        if (s == 0x12345678u) out[0] = s;
    }
}

// Fr-mul ceiling: register-resident Montgomery multiplications, two dependent chains per thread.
template <int SPLIT>
__global__ void __launch_bounds__(256) frmul_bench_kernel(uint32_t* out, uint32_t seed, int iters) {
    fr::Fe a, b;
#pragma unroll
    for (int i = 0; i < 8; ++i) { a.l[i] = seed * (i + 3) + threadIdx.x; b.l[i] = seed * (i + 7) + blockIdx.x; }
    a.l[7] &= 0x0FFFFFFF; b.l[7] &= 0x0FFFFFFF;
    for (int it = 0; it < iters; ++it) {
        fr::mont_mul_s<SPLIT>(a, a, b);
        fr::mont_mul_s<SPLIT>(b, b, a);
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a.l[i] ^ b.l[i];
    if (s == 0x12345678u) out[0] = s;
}

cudaError_t frmul_microbench(double* fr_mul_per_s /*[5]: pipe-split level 0..4*/) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint32_t* d = nullptr;
    cudaError_t e = cudaMalloc(&d, 4);
    if (e != cudaSuccess) return e;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    const int iters = 512, blocks = sms * 8, threads = 256;
    for (int sp = 0; sp < 5; ++sp) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(t0);
            switch (sp) {
                case 0: frmul_bench_kernel<0><<<blocks, threads>>>(d, 17 + rep, iters); break;
                case 1: frmul_bench_kernel<1><<<blocks, threads>>>(d, 17 + rep, iters); break;
                case 2: frmul_bench_kernel<2><<<blocks, threads>>>(d, 17 + rep, iters); break;
                case 3: frmul_bench_kernel<3><<<blocks, threads>>>(d, 17 + rep, iters); break;
                default: frmul_bench_kernel<4><<<blocks, threads>>>(d, 17 + rep, iters); break;
            }
            cudaEventRecord(t1);
            e = cudaEventSynchronize(t1);
            if (e != cudaSuccess) return e;
            float ms;
            cudaEventElapsedTime(&ms, t0, t1);
            if (rep > 0 && ms < best) best = ms;
        }
        if (fr_mul_per_s) fr_mul_per_s[sp] = 2.0 * iters * blocks * threads / (best * 1e-3);
    }
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(d);
    return cudaGetLastError();
}

cudaError_t imad_microbench(double* imad32_per_s, double* imad_wide_per_s, double* imad_wide_carry_per_s, double* sm_clock_mhz) {
    int dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    if (sm_clock_mhz) *sm_clock_mhz = khz / 1000.0;
    uint32_t* d = nullptr;
    cudaError_t e = cudaMalloc(&d, 4);
    if (e != cudaSuccess) return e;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    const int iters = 4096, blocks = sms * 8, threads = 256;
    double res[3] = {0, 0, 0};
    for (int v = 0; v < 3; ++v) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(t0);
            if (v == 0) imad_bench_kernel<0><<<blocks, threads>>>(d, 17 + rep, iters);
            if (v == 1) imad_bench_kernel<1><<<blocks, threads>>>(d, 17 + rep, iters);
            if (v == 2) imad_bench_kernel<2><<<blocks, threads>>>(d, 17 + rep, iters);
            cudaEventRecord(t1);
            e = cudaEventSynchronize(t1);
            if (e != cudaSuccess) return e;
            float ms;
            cudaEventElapsedTime(&ms, t0, t1);
            if (rep > 0 && ms < best) best = ms;
        }
        double per_thread = (v == 2) ? (double)iters * 2 * 16 : (double)iters * 4 * 8;
        res[v] = per_thread * blocks * threads / (best * 1e-3);
    }
    if (imad32_per_s) *imad32_per_s = res[0];
    if (imad_wide_per_s) *imad_wide_per_s = res[1];
    if (imad_wide_carry_per_s) *imad_wide_carry_per_s = res[2];
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(d);
    return cudaGetLastError();
}

}  // namespace acvmb
