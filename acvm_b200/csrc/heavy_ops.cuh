// Heavy blackbox micro-ops (SHA-256, Keccak-256, Grumpkin) -- FULL kernel variant only.
#pragma once
#include "fr.cuh"
#include "plan.hpp"
namespace acvmb {
template <int T>
__device__ __forceinline__ void exec_heavy(const OpRec* r, uint32_t kind, uint32_t flags, uint4* cb, unsigned long long* fail,
                                           const uint32_t* payload) {
    (void)r; (void)kind; (void)flags; (void)cb; (void)fail; (void)payload;
}
}  // namespace acvmb
