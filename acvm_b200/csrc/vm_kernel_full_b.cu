// FULL step-VM variant, tile shapes with T = 32 (see vm_kernel_full.cu).
#define ACVMB_HEAVY_OPS_TU 1
#include "vm_kernel_impl.cuh"

namespace acvmb {

cudaError_t set_curve_tables_b(const uint32_t* fixed_base, const uint32_t* pedersen) {
    CurveTables t{fixed_base, pedersen};
    return cudaMemcpyToSymbol(g_curve_tables, &t, sizeof(t));
}

cudaError_t launch_vm_full_b(const KernelConfig& cfg, const VmArgs& args, cudaStream_t stream) {
#define X(t, s) if (cfg.T == t && cfg.S == s) return launch_one<t, s, true>(args, stream);
    ACVMB_CONFIGS_FULL_B(X)
#undef X
    return cudaErrorInvalidConfiguration;
}

}  // namespace acvmb
