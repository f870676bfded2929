// Host build of acvm_b200/csrc/fr.cuh (carry flag emulated) so the limb algorithms can be
// checked against Python big ints on a CPU-only box.  Test shim only -- never shipped.
#include "../../acvm_b200/csrc/fr.cuh"
#include <cstring>
using namespace fr;
extern "C" {
void t_mont_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) {
    Fe A, B, R; memcpy(A.l, a, 32); memcpy(B.l, b, 32); mont_mul(R, A, B); memcpy(r, R.l, 32);
}
void t_mont_dot(int k, const uint32_t* a, const uint32_t* b, uint32_t* r) {
    Fe A[4], B[4], R;
    for (int i = 0; i < k; ++i) { memcpy(A[i].l, a + 8 * i, 32); memcpy(B[i].l, b + 8 * i, 32); }
    const Fe* pa[4] = {&A[0], &A[1], &A[2], &A[3]};
    const Fe* pb[4] = {&B[0], &B[1], &B[2], &B[3]};
    if (k == 1) mont_dot_raw<1>(R, pa, pb);
    else if (k == 2) mont_dot_raw<2>(R, pa, pb);
    else if (k == 3) mont_dot_raw<3>(R, pa, pb);
    else mont_dot_raw<4>(R, pa, pb);
    memcpy(r, R.l, 32);
}
void t_mont_dot_init(int k, const uint32_t* a, const uint32_t* b, const uint32_t* init, uint32_t* r) {
    Fe A[4], B[4], R;
    for (int i = 0; i < k; ++i) { memcpy(A[i].l, a + 8 * i, 32); memcpy(B[i].l, b + 8 * i, 32); }
    const Fe* pa[4] = {&A[0], &A[1], &A[2], &A[3]};
    const Fe* pb[4] = {&B[0], &B[1], &B[2], &B[3]};
    if (k == 1) mont_dot_fn<1>(R, pa, PtrLimbs{pb}, init);
    else if (k == 2) mont_dot_fn<2>(R, pa, PtrLimbs{pb}, init);
    else mont_dot_fn<3>(R, pa, PtrLimbs{pb}, init);
    memcpy(r, R.l, 32);
}
void t_mont_mul_split(int split, const uint32_t* a, const uint32_t* b, uint32_t* r) {
    Fe A, B, R; memcpy(A.l, a, 32); memcpy(B.l, b, 32);
    switch (split) { case 0: mont_mul_s<0>(R, A, B); break; case 1: mont_mul_s<1>(R, A, B); break; case 2: mont_mul_s<2>(R, A, B); break;
                     case 3: mont_mul_s<3>(R, A, B); break; default: mont_mul_s<4>(R, A, B); break; }
    memcpy(r, R.l, 32);
}
void t_add(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fe A, B, R; memcpy(A.l, a, 32); memcpy(B.l, b, 32); add_mod(R, A, B); memcpy(r, R.l, 32); }
void t_sub(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fe A, B, R; memcpy(A.l, a, 32); memcpy(B.l, b, 32); sub_mod(R, A, B); memcpy(r, R.l, 32); }
void t_reduce(uint32_t* a) { Fe A; memcpy(A.l, a, 32); reduce_256(A); memcpy(a, A.l, 32); }
void t_inv_bea(const uint32_t* a, uint32_t* r) { Fe A, R; memcpy(A.l, a, 32); inv_bea(R, A); memcpy(r, R.l, 32); }
uint32_t t_num_bits(const uint32_t* a) { Fe A; memcpy(A.l, a, 32); return num_bits(A); }
}

// ---- 9 x 29-bit carry-free representation (fr29.cuh) ----
#include "fr29.cuh"   // rejected experiment (measured slower), kept as a tested reference for a future 29-bit-limb column format
extern "C" {
void t_dot9(int k, const uint32_t* a8, const uint32_t* b8, uint32_t* r8) {
    fr::Fe A[4], B[4];
    fr29::Fe9 A9[4], B9[4];
    for (int i = 0; i < k; ++i) {
        memcpy(A[i].l, a8 + 8 * i, 32); memcpy(B[i].l, b8 + 8 * i, 32);
        fr29::to9(A9[i], A[i]); fr29::to9(B9[i], B[i]);
    }
    const fr29::Fe9* pa[4] = {&A9[0], &A9[1], &A9[2], &A9[3]};
    const fr29::Fe9* pb[4] = {&B9[0], &B9[1], &B9[2], &B9[3]};
    uint32_t out[9];
    if (k == 1) fr29::mont_dot9<1>(out, pa, fr29::PtrLimbs9{pb});
    else if (k == 2) fr29::mont_dot9<2>(out, pa, fr29::PtrLimbs9{pb});
    else if (k == 3) fr29::mont_dot9<3>(out, pa, fr29::PtrLimbs9{pb});
    else fr29::mont_dot9<4>(out, pa, fr29::PtrLimbs9{pb});
    fr::Fe R; fr29::from9(R, out); memcpy(r8, R.l, 32);
}
void t_roundtrip9(const uint32_t* a8, uint32_t* r8) {
    fr::Fe A, R; memcpy(A.l, a8, 32);
    fr29::Fe9 x; fr29::to9(x, A); fr29::from9(R, x.l); memcpy(r8, R.l, 32);
}
}
