// Device-side interface between the host runtime (runtime.cu) and the step-VM kernels (vm_kernel.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace acvmb {

// Witness storage ("columns"), tile-major so one CTA streams one contiguous region:
//   uint4 cols[tile][slot][plane(2)][T]      plane 0 = limbs 0..3, plane 1 = limbs 4..7 (canonical, LE)
// A tile is T consecutive instances; every load/store of a witness by a tile is T*16 contiguous bytes
// per plane (one 128 B line for T = 8).
struct VmArgs {
    const uint8_t* stream;        // OpRec[n_steps][S]
    const uint32_t* payload;      // variable-length operand lists
    uint4* cols;
    unsigned long long* fail;     // per instance: min over failures of (opcode<<32 | kind<<28 | aux)
    uint32_t first_step;          // this launch covers steps [first_step, first_step + n_steps)
    uint32_t n_steps;             // multiple of chunk_steps
    uint32_t chunk_steps;
    uint32_t n_stage;             // depth of the shared-memory staging ring (1..8)
    uint32_t n_slots;
    uint32_t n_tiles;
    uint32_t* mu_assign;          // [tile][n_mu][T]: opcode index that assigned a value-dependent witness, ~0 = unassigned
    uint32_t n_mu;
    uint32_t ring_slots;          // entries of the shared-memory ring of recent values the stream was compiled for (0 = none)
};

struct KernelConfig {
    int T, S;
    bool full;
    int split = -1;   // FMA/ALU pipe-split level of the field multiplier (-1 = build default), see fr.cuh
};

// returns cudaSuccess or the launch error; smem_bytes/regs reported for the run record
cudaError_t launch_vm(const KernelConfig& cfg, const VmArgs& args, cudaStream_t stream);
bool vm_config_supported(int T, int S, bool full);
// device pointers of the Grumpkin lookup tables (built by the host once per context)
cudaError_t set_curve_tables(const uint32_t* fixed_base, const uint32_t* pedersen);

cudaError_t launch_scatter_inputs(const uint8_t* in_be, const uint32_t* input_slots, uint32_t n_inputs, uint4* cols,
                                  uint32_t n_slots, int T, uint32_t n_inst, cudaStream_t stream);
struct GatherArgs {
    const uint4* cols;
    uint32_t n_slots;
    int T;
    const uint32_t* witness_ids;   // NULL = dense 0..n_out-1
    uint32_t n_out, first_inst, n_inst;
    const unsigned long long* fail;
    const uint32_t* assign_opcode;
    const uint32_t* mu_index_of;   // per witness, for value-dependent presence
    const uint32_t* mu_assign;
    uint32_t n_mu;
    uint32_t static_fail_opcode;
    uint8_t* out_be;               // [n_inst][n_out][32] or NULL
    uint8_t* out_present;          // [n_inst][n_out] or NULL
    int raw;                       // 1 = witness_ids are raw column slots (temporaries allowed), no presence logic
    const uint32_t* unscale;       // scaled columns: per witness 8 limbs of (1/lambda_w)*R, or NULL when every column is canonical
};
cudaError_t launch_gather_outputs(const GatherArgs& g, cudaStream_t stream);
// out[inst] = sum over the witnesses instance `inst` holds of mix(index, value)  (g.n_out = num_witnesses, g.n_inst instances from 0)
cudaError_t launch_checksum(const GatherArgs& g, unsigned long long* out, cudaStream_t stream);
cudaError_t launch_fill_u64(unsigned long long* p, size_t n, unsigned long long v, cudaStream_t stream);

// IMAD roofline micro-benchmark: returns measured multiply-accumulates per second for each variant
cudaError_t imad_microbench(double* imad32_per_s, double* imad_wide_per_s, double* imad_wide_carry_per_s, double* sm_clock_mhz);

// register-resident Montgomery multiplications per second (practical Fr-mul ceiling of this field library)
cudaError_t frmul_microbench(double* fr_mul_per_s /*[5]*/);
cudaError_t run_imad_cc_microbench(double* out3);

}  // namespace acvmb
