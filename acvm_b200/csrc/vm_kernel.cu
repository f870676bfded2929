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
#include "vm_kernel_impl.cuh"

namespace acvmb {

bool vm_config_supported(int T, int S, bool full) {
#define X(t, s) if (T == t && S == s) return true;
    if (full) { ACVMB_CONFIGS_FULL(X) } else { ACVMB_CONFIGS_LIGHT(X) }
#undef X
    return false;
}

cudaError_t launch_vm(const KernelConfig& cfg, const VmArgs& args, cudaStream_t stream) {
    if (cfg.full) return launch_vm_full(cfg, args, stream);
    // tuning hook: explicit FMA/ALU pipe-split level for the production tile shape of the arithmetic kernel
    if (cfg.split >= 0 && cfg.S == 16 && cfg.T == 8) {
#define Y(sp) if (cfg.split == sp) return launch_one<8, 16, false, sp>(args, stream);
        Y(0) Y(1) Y(2) Y(3) Y(4)
#undef Y
    }
#define X(t, s) if (cfg.T == t && cfg.S == s) return launch_one<t, s, false>(args, stream);
    ACVMB_CONFIGS_LIGHT(X)
#undef X
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
    uint32_t slot = input_slots[k];
    if (slot & 0x80000000u) {   // scaled column (plan.cpp): the plan wants this input in Montgomery form, value * R
        slot &= 0x7FFFFFFFu;
        Fe r2;
        r2.l[0] = 0xae216da7u; r2.l[1] = 0x1bb8e645u; r2.l[2] = 0xe35c59e3u; r2.l[3] = 0x53fe3ab1u;
        r2.l[4] = 0x53bb8085u; r2.l[5] = 0x8c49833du; r2.l[6] = 0x7f4e44a5u; r2.l[7] = 0x0216d0b1u;
        fr::mont_mul(v, v, r2);
    }
    uint32_t tile = inst / T, lane = inst % T;
    uint4* p = cols + ((size_t)tile * n_slots + slot) * (2 * T) + lane;
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
// scaled columns (plan.cpp): the column holds lambda_w * value; unscale[w] = (1/lambda_w) * R, so one Montgomery product
// gives the canonical value
__device__ __forceinline__ void unscale_value(const uint32_t* unscale, uint32_t w, uint4& lo, uint4& hi) {
    const uint4* u = reinterpret_cast<const uint4*>(unscale + (size_t)w * 8);
    const uint4 ul = u[0], uh = u[1];
    Fe s, m;
    s.l[0] = lo.x; s.l[1] = lo.y; s.l[2] = lo.z; s.l[3] = lo.w; s.l[4] = hi.x; s.l[5] = hi.y; s.l[6] = hi.z; s.l[7] = hi.w;
    m.l[0] = ul.x; m.l[1] = ul.y; m.l[2] = ul.z; m.l[3] = ul.w; m.l[4] = uh.x; m.l[5] = uh.y; m.l[6] = uh.z; m.l[7] = uh.w;
    fr::mont_mul(s, s, m);
    lo = make_uint4(s.l[0], s.l[1], s.l[2], s.l[3]);
    hi = make_uint4(s.l[4], s.l[5], s.l[6], s.l[7]);
}

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
        if (g.unscale && !g.raw) unscale_value(g.unscale, w, lo, hi);
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
        uint4 lo = p[0], hi = p[g.T];
        if (g.unscale) unscale_value(g.unscale, w, lo, hi);
        acc += mix_witness(w, lo, hi);
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
    } else if (VARIANT >= 3) {
        // 3: wide multiply-add with carry-OUT only, the carry captured by an ALU addc into a side counter (lazy-carry rows)
        // 4: carry-out only, carry dropped      5: chains of two [carry-out, carry-in+out] with one capture per chain
        uint32_t lo[8], hi[8], c[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { lo[i] = a + i; hi[i] = b + i; c[i] = i; }
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (VARIANT == 5) {
#pragma unroll
                    for (int i = 0; i < 8; i += 2) {
                        fr::mad_lo_cc(lo[i], a + i, b + u, lo[i]);
                        fr::madc_hi_cc(hi[i], a + i, b + u, hi[i]);
                        fr::madc_lo_cc(lo[i + 1], a + i, b ^ u, lo[i + 1]);
                        fr::madc_hi_cc(hi[i + 1], a + i, b ^ u, hi[i + 1]);
                        fr::addc(c[i], c[i], 0);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        fr::mad_lo_cc(lo[i], a + i, b + u, lo[i]);
                        fr::madc_hi_cc(hi[i], a + i, b + u, hi[i]);
                        if (VARIANT == 3) fr::addc(c[i], c[i], 0);
                    }
                }
            }
        }
        uint32_t s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s ^= lo[i] ^ hi[i] ^ c[i];
        if (s == 0x12345678u) out[0] = s;
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

// experiment: rates of the carry-out-only forms (wide IMADs per second), see imad_bench_kernel variants 3..5
cudaError_t run_imad_cc_microbench(double* out3) {
    const int blocks = 148 * 8, threads = 256, iters = 4096;
    uint32_t* d = nullptr;
    cudaError_t e = cudaMalloc(&d, 64);
    if (e != cudaSuccess) return e;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    for (int v = 0; v < 3; ++v) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(t0);
            if (v == 0) imad_bench_kernel<3><<<blocks, threads>>>(d, 17 + rep, iters);
            if (v == 1) imad_bench_kernel<4><<<blocks, threads>>>(d, 17 + rep, iters);
            if (v == 2) imad_bench_kernel<5><<<blocks, threads>>>(d, 17 + rep, iters);
            cudaEventRecord(t1);
            e = cudaEventSynchronize(t1);
            if (e != cudaSuccess) return e;
            float ms;
            cudaEventElapsedTime(&ms, t0, t1);
            if (rep > 0 && ms < best) best = ms;
        }
        out3[v] = (double)iters * 4 * 8 * blocks * threads / (best * 1e-3);
    }
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(d);
    return cudaGetLastError();
}

}  // namespace acvmb
