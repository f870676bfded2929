// Host build of acvm_b200/csrc/ecdsa.cuh so the verification routine the GPU lanes run can be checked against
// oracle/ecdsa.py on a CPU-only box.  Test shim only -- never shipped.
#include "../../acvm_b200/csrc/ecdsa.cuh"
extern "C" int t_ecdsa_verify(int curve, const uint8_t* hashed_msg, const uint8_t* pkx, const uint8_t* pky, const uint8_t* sig) {
    return ec::ecdsa_verify(curve, hashed_msg, pkx, pky, sig);
}
