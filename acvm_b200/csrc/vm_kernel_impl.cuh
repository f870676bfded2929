// Step-VM kernel template shared by the two translation units that instantiate it (vm_kernel.cu: arithmetic/logic variant,
// vm_kernel_full.cu: + hash / curve / general ops), so they compile in parallel.
#pragma once
// (see vm_kernel.cu for the design notes)
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
#ifdef ACVMB_HEAVY_OPS_TU
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
// Operand of a gate / logic / range micro-op: a column, or (bit 31 of the field) an entry of the tile's shared-memory ring of
// recent values -- same [plane][lane] layout as a column, ~30 cycles away instead of an L2 round trip.  One generic load
// serves both, so slots of a warp that differ in where their operand lives do not diverge.
template <int T, bool RING>
__device__ __forceinline__ void load_op(Fe& v, const uint4* cb, const uint4* ring, uint32_t w) {
    const uint4* p = cb + (size_t)w * (2 * T);
    if constexpr (RING) p = (w & RING_FLAG) ? ring + (size_t)(w & ~RING_FLAG) * (2 * T) : p;
    uint4 lo = p[0], hi = p[T];
    v.l[0] = lo.x; v.l[1] = lo.y; v.l[2] = lo.z; v.l[3] = lo.w;
    v.l[4] = hi.x; v.l[5] = hi.y; v.l[6] = hi.z; v.l[7] = hi.w;
}
template <int T, bool RING>
__device__ __forceinline__ void store_ring(uint4* ring, uint32_t entry, const Fe& v) {
    if constexpr (!RING) return;
    if (ring == nullptr || entry == RING_NONE) return;
    uint4* p = ring + (size_t)entry * (2 * T);
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

// b-limb source of the gate's dot product: plan constants fetched limb by limb from shared memory, except that the first
// operand of a one-reduction multiplicative gate is the lane's own (y + beta)
struct GateLimbs {
    const uint32_t* c0;
    const uint32_t* c1;
    const uint32_t* c2;
    const Fe& yreg;
    bool k0_reg;
    __device__ __forceinline__ uint32_t operator()(int k, int i) const {
        if (k == 0) return k0_reg ? yreg.l[i] : c0[i];
        return k == 1 ? c1[i] : c2[i];
    }
};

// Arithmetic gate on STORED column values (plan.cpp "scaled columns"; canonical columns are the special case lambda = 1):
//   GF_MUL|GF_ONE_RED : out = ( (x+c1)*(y+c2) + w1*c3 + c4 ) / R                        one reduction, width 1..2
//   GF_MUL            : u = (x+c1)*(y+c2)/R ; out = ( u*c0 + w1*c3 + c4 ) / R            two reductions
//   otherwise         : out = ( y*c1 + w1*c2 + w2*c3 + c4 ) / R                          one reduction, width 0..3
// where only the first GF_NPROD linear operands are products; the rest are added or subtracted after the reduction, and a
// gate without any product is c4 +- operands with no multiplication at all.
template <int T, int SPLIT, bool RING>
__device__ __forceinline__ void exec_gate(const OpRec* r, uint32_t kind, uint32_t flags, uint4* cb, uint4* ring,
                                          unsigned long long* fail) {
    Fe res;
    if (flags & GF_Y) {
        const uint32_t nlin = (flags >> GF_NLIN_SHIFT) & 3;
        const uint32_t nprod = (flags >> GF_NPROD_SHIFT) & 3;
        const bool mul = (flags & GF_MUL) != 0;
        // Every operand load is issued here, before the code paths part: the slots of a warp may hold different gate forms,
        // the forms then run one after the other, and with the loads inside each form their L2 latencies added up.
        Fe x, y, w1, w2;
        if (mul) load_op<T, RING>(x, cb, ring, r->w[3]);
        load_op<T, RING>(y, cb, ring, r->w[4]);
        if (nlin >= 1) load_op<T, RING>(w1, cb, ring, r->w[5]);
        if (nlin >= 2) load_op<T, RING>(w2, cb, ring, r->w[6]);
        const uint32_t K = (mul ? 1u : 0u) + nprod;   // width of this lane's dot product
        if (K == 0) {
            lds_fe(res, r->c[4]);
        } else {
            // Every form shares ONE reduction of warp-uniform width: the lanes of this warp that are here agree on the widest
            // dot product among them, narrower lanes pad with zero operands.
            const unsigned lanes = __activemask();
            const bool one_red = (flags & GF_ONE_RED) != 0;
            Fe a1, a2;   // product operands 1 and 2 (zero when this lane has none)
            if (mul) {
                // lazy reduction: x+c1, y+c2 < 2p stay unreduced ((4p^2 + p^2)/R + p < 1.95p: one conditional subtraction)
                Fe t;
                lds_fe(t, r->c[1]);
                fr::add_raw(x, x, t);
                lds_fe(t, r->c[2]);
                fr::add_raw(y, y, t);
                if (!one_red) {
                    const Fe* p1[1] = {&x};
                    fr::mont_dot_fn<1, RegLimbs, SPLIT>(t, p1, RegLimbs{y});   // u = (x+c1)(y+c2)/R < 1.76p
                    x = t;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) { a1.l[i] = nprod >= 1 ? w1.l[i] : 0u; a2.l[i] = 0u; }
            } else {
                x = y;
#pragma unroll
                for (int i = 0; i < 8; ++i) { a1.l[i] = nprod >= 2 ? w1.l[i] : 0u; a2.l[i] = nprod >= 3 ? w2.l[i] : 0u; }
            }
            __syncwarp(lanes);
            const uint32_t Kmax = __reduce_max_sync(lanes, K);
            // operand k pairs with: multiplicative (x', w1) . (y' or c0, c3); linear (y, w1, w2) . (c1, c2, c3)
            const GateLimbs bl{mul ? r->c[0] : r->c[1], mul ? r->c[3] : r->c[2], r->c[3], y, mul && one_red};
            const Fe* a[3] = {&x, &a1, &a2};
            if (Kmax == 1) fr::mont_dot_fn<1, GateLimbs, SPLIT>(res, a, bl, r->c[4]);
            else if (Kmax == 2) fr::mont_dot_fn<2, GateLimbs, SPLIT>(res, a, bl, r->c[4]);
            else fr::mont_dot_fn<3, GateLimbs, SPLIT>(res, a, bl, r->c[4]);
            fr::cond_sub_p(res);
        }
        // linear operands that are plain additions: index >= nprod in (y, w1, w2), resp. (w1) for a multiplicative gate
        if (mul) {
            if (nlin >= 1 && nprod == 0) {
                if (flags & GF_NEG_W1) fr::sub_mod(res, res, w1); else fr::add_mod(res, res, w1);
            }
        } else {
            if (nprod == 0) {
                if (flags & GF_NEG_Y) fr::sub_mod(res, res, y); else fr::add_mod(res, res, y);
            }
            if (nlin >= 1 && nprod <= 1) {
                if (flags & GF_NEG_W1) fr::sub_mod(res, res, w1); else fr::add_mod(res, res, w1);
            }
            if (nlin >= 2 && nprod <= 2) {
                if (flags & GF_NEG_W2) fr::sub_mod(res, res, w2); else fr::add_mod(res, res, w2);
            }
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
        store_ring<T, RING>(ring, r->w[7], res);
    } else {
        if (!fr::is_zero(res)) record_fail(fail, r->w[1], EK_UNSATISFIED_CONSTRAIN, 0);
    }
}

// AND / XOR on the low `nb` bits of the canonical values (acir_field/src/generic_ark.rs:322-354,446-473)
template <int T, bool RING>
__device__ __forceinline__ void exec_logic(const OpRec* r, uint32_t kind, uint32_t flags, uint4* cb, uint4* ring,
                                           unsigned long long* fail) {
    Fe x, y, res;
    load_op<T, RING>(x, cb, ring, r->w[3]);
    load_op<T, RING>(y, cb, ring, r->w[4]);
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
    store_ring<T, RING>(ring, r->w[5], res);
}

template <int T, bool RING>
__device__ __forceinline__ void exec_range(const OpRec* r, uint4* cb, const uint4* ring, unsigned long long* fail) {
    Fe x;
    load_op<T, RING>(x, cb, ring, r->w[3]);
    if (fr::num_bits(x) > r->w[7]) record_fail(fail, r->w[1], EK_UNSATISFIED_CONSTRAIN, 0);
}

// FULL variant: the curve / hash code wants ~128 registers, which would allow only 4 CTAs of 128 threads per SM; the
// plan serialises heavy ops along each instance's dependency chain, so run time is (number of CTA waves) x (sum of heavy-op
// latencies) and keeping the whole sub-batch in ONE wave matters more than spill-free heavy ops: cap at 7 CTAs/SM.
// CAP selects the register-capped build; the launcher uses it only when the uncapped one could not hold the sub-batch in a
// single wave (it costs ~30 % on Keccak, whose state then spills).
template <int T, int S, bool FULL, int SPLIT, bool CAP = false, bool RING = false>
__global__ void __launch_bounds__(T* S, (T * S <= 128) ? (CAP ? 896 / (T * S) : (FULL ? 512 / (T * S) : 1)) : ((FULL && T * S <= 512) ? 512 / (T * S) : 1)) vm_kernel(const VmArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t chunk_bytes = a.chunk_steps * S * (uint32_t)sizeof(OpRec);
    const uint32_t NSTAGE = a.n_stage;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)NSTAGE * chunk_bytes);
    // ring of recent values: [entry][plane][lane] uint4, after the staging ring and its barriers (16 B aligned)
    uint4* ring = (RING && a.ring_slots)
                      ? reinterpret_cast<uint4*>(smem + (((size_t)NSTAGE * chunk_bytes + NSTAGE * sizeof(uint64_t) + 15) & ~(size_t)15))
                      : nullptr;

    const uint32_t tid = threadIdx.x;
    const uint32_t slot = tid / T;   // T < 32: a warp holds 32 / T consecutive slots (plan.cpp Scheduler::emit orders them for that)
    const uint32_t lane = tid % T;
    const uint32_t tile = blockIdx.x;
    uint4* cb = a.cols + (size_t)tile * a.n_slots * (2 * T) + lane;
    uint4* ring_lane = ring ? ring + lane : nullptr;
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
                    exec_gate<T, SPLIT, RING>(r, kind, flags, cb, ring_lane, fail);
                    break;
                case MK_AND:
                case MK_XOR:
                    exec_logic<T, RING>(r, kind, flags, cb, ring_lane, fail);
                    break;
                case MK_RANGE:
                    exec_range<T, RING>(r, cb, ring_lane, fail);
                    break;
                default:
#ifdef ACVMB_HEAVY_OPS_TU
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
    if (args.ring_slots) {
        if (FULL) return cudaErrorInvalidValue;   // the ring variant is built for the arithmetic / logic kernel only
        smem = ((smem + 15) & ~(size_t)15) + (size_t)args.ring_slots * T * 32;
    }
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    auto k = vm_kernel<T, S, FULL, SPLIT, false>;
    if constexpr (!FULL) {
        if (args.ring_slots) k = vm_kernel<T, S, FULL, SPLIT, false, true>;
    }
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
#define ACVMB_CONFIGS_LIGHT(X) X(8, 16) X(4, 16) X(2, 16) X(16, 16) X(16, 8) X(32, 4) X(32, 2) X(32, 1) X(4, 32) X(2, 64) X(32, 8)
// the FULL shapes are compiled in two translation units (vm_kernel_full.cu / vm_kernel_full_b.cu) to halve the build's critical path
#define ACVMB_CONFIGS_FULL_A(X) X(8, 16) X(4, 16) X(16, 8)
#define ACVMB_CONFIGS_FULL_B(X) X(32, 4) X(32, 1) X(32, 8)
#define ACVMB_CONFIGS_FULL(X) ACVMB_CONFIGS_FULL_A(X) ACVMB_CONFIGS_FULL_B(X)

cudaError_t launch_vm_full(const KernelConfig& cfg, const VmArgs& args, cudaStream_t stream);   // vm_kernel_full.cu
cudaError_t launch_vm_full_b(const KernelConfig& cfg, const VmArgs& args, cudaStream_t stream); // vm_kernel_full_b.cu
cudaError_t set_curve_tables_b(const uint32_t* fixed_base, const uint32_t* pedersen);

}  // namespace acvmb
