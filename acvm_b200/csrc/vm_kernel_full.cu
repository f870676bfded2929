// FULL step-VM variant (hash / curve / general / directive / memory micro-ops), tile shapes with T <= 16: its own translation
// unit so that it compiles in parallel with the arithmetic variant and with the T = 32 shapes (vm_kernel_full_b.cu).
// Each of the two units has its own copy of the __constant__ curve-table pointers.
#define ACVMB_HEAVY_OPS_TU 1
#include "vm_kernel_impl.cuh"

namespace acvmb {

cudaError_t set_curve_tables(const uint32_t* fixed_base, const uint32_t* pedersen) {
    CurveTables t{fixed_base, pedersen};
    cudaError_t e = cudaMemcpyToSymbol(g_curve_tables, &t, sizeof(t));
    return e != cudaSuccess ? e : set_curve_tables_b(fixed_base, pedersen);
}

cudaError_t launch_vm_full(const KernelConfig& cfg, const VmArgs& args, cudaStream_t stream) {
#define X(t, s) if (cfg.T == t && cfg.S == s) return launch_one<t, s, true>(args, stream);
    ACVMB_CONFIGS_FULL_A(X)
#undef X
    return launch_vm_full_b(cfg, args, stream);
}

}  // namespace acvmb
