// Host driver for the REFERENCE's own CUDA helper kernels (TEST INFRASTRUCTURE).
//
// The reference ships the pack / unpack / pointwise pieces of the vpsi + rhoofr path as plain CUDA
// kernels in src/cuuser_utils_kernels.cu (called from fftmain_utils / vpsi_utils / rhoofr_utils when
// cp_cuda_env%use_fft).  They are the only part of the path that compiles from its own source files
// without the Fortran tool chain, so they are compiled here - from where they lie under
// /root/reference, nothing is copied - for the host (see ref_shim/) and used by tests/test_oracle_ref.py
// to pin the oracle's restatement of exactly these steps:
//   set_psi_2_states_g / set_psi_1_state_g, build_density_sum, the pointwise V*psi, phasen,
//   putz / getz, unpack_x2y / pack_y2x (ray -> plane index convention).
// The launch geometry of the reference's wrappers (cuuser_utils.cu) is irrelevant to the result: every
// kernel guards its indices, so the driver runs each over a grid that covers the index range.
#include <cmath>
#include <cstdio>
#include <stdbool.h>

#ifndef REF_KERNELS_CU
#error "compile with -DREF_KERNELS_CU='\"/root/reference/src/cuuser_utils_kernels.cu\"'"
#endif
#define __CUDA
#include REF_KERNELS_CU

ref_uint3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0};
dim3 blockDim(1, 1, 1), gridDim(1, 1, 1);

namespace {
// run `body` once per thread of a (nx, ny, nz) index space, one thread per block
template <class F>
void for_each_thread(long nx, long ny, long nz, F&& body) {
  blockDim = dim3(1, 1, 1);
  gridDim = dim3((unsigned)nx, (unsigned)ny, (unsigned)nz);
  threadIdx = {0, 0, 0};
  for (unsigned z = 0; z < (unsigned)nz; ++z)
    for (unsigned y = 0; y < (unsigned)ny; ++y)
      for (unsigned x = 0; x < (unsigned)nx; ++x) {
        blockIdx = {x, y, z};
        body();
      }
}
}  // namespace

extern "C" {

void ref_set_psi_2_states_g(void* c1, void* c2, void* psi, int jgw, int* nzfs, int* inzs, int geq0) {
  for_each_thread(jgw, 1, 1, [&] {
    CuUser_Kernel_Set_Psi_2_Stages_G((cuDoubleComplex*)c1, (cuDoubleComplex*)c2, (cuDoubleComplex*)psi, jgw, nzfs, inzs,
                                     geq0 != 0);
  });
}

void ref_set_psi_1_state_g(double alpha_re, double alpha_im, void* c1, void* psi, int jgw, int* nzfs, int* inzs,
                           int geq0) {
  for_each_thread(jgw, 1, 1, [&] {
    CuUser_Kernel_Set_Psi_1_Stage_G(alpha_re, alpha_im, (cuDoubleComplex*)c1, (cuDoubleComplex*)psi, jgw, nzfs, inzs,
                                    geq0 != 0);
  });
}

void ref_build_density_sum(double a_re, double a_im, void* psi, double* rho, int n) {
  for_each_thread(n, 1, 1, [&] { CuUser_Kernel_Build_Density_Sum(a_re, a_im, (cuDoubleComplex*)psi, rho, n); });
}

void ref_pointwise_cxr(void* xf, double* yf, int n) {
  for_each_thread(n, 1, 1, [&] { CuUser_Kernel_Pointwise_CxR((cuDoubleComplex*)xf, yf, n); });
}

void ref_phasen(void* f, int kr1, int kr2s, int kr3s, int n1u, int n1o, int nr2s, int nr3s) {
  for_each_thread(n1o - n1u + 1, nr2s, nr3s,
                  [&] { CuUser_Kernel_PhaseN((cuDoubleComplex*)f, kr1, kr2s, kr3s, n1u, n1o, nr2s, nr3s); });
}

// CuUser_C_PutZ (cuuser_utils.cu:54-73): zero b(kr,m), then MatMov(n, m, a, n, b(krmin,1), kr)
void ref_putz(void* a, void* b, int krmin, int krmax, int kr, int m) {
  cuDoubleComplex* b_p = (cuDoubleComplex*)b;
  const int size = m * kr, n = krmax - krmin + 1;
  for_each_thread(size, 1, 1, [&] { CuUser_Kernel_Zeroing(b_p, size); });
  for_each_thread(n, m, 1, [&] { CuUser_Kernel_MatMov(n, m, (cuDoubleComplex*)a, n, &b_p[krmin - 1], kr); });
}

// CuUser_C_GetZ (cuuser_utils.cu:75-86)
void ref_getz(void* a, void* b, int krmin, int krmax, int kr, int m) {
  cuDoubleComplex* a_p = (cuDoubleComplex*)a;
  const int n = krmax - krmin + 1;
  for_each_thread(n, m, 1, [&] { CuUser_Kernel_MatMov(n, m, &a_p[krmin - 1], kr, (cuDoubleComplex*)b, n); });
}

void ref_unpack_x2y(void* xf, void* yf, int m, int lr1, int lda, int* msp, int lmsp, int* sp8, int maxfft, int mproc) {
  int mxrp = 0;
  for (int ip = 0; ip < mproc; ++ip) mxrp = sp8[ip] > mxrp ? sp8[ip] : mxrp;
  for_each_thread(mproc, mxrp, lr1, [&] {
    CuUser_Kernel_Unpack_x2y_8((cuDoubleComplex*)xf, (cuDoubleComplex*)yf, m, lr1, lda, msp, lmsp, sp8, maxfft, mproc);
  });
}

void ref_pack_y2x(void* xf, void* yf, int m, int lr1, int lda, int* msp, int lmsp, int* sp8, int maxfft, int mproc) {
  int mxrp = 0;
  for (int ip = 0; ip < mproc; ++ip) mxrp = sp8[ip] > mxrp ? sp8[ip] : mxrp;
  for_each_thread(mproc, mxrp, lr1, [&] {
    CuUser_Kernel_Pack_y2x_8((cuDoubleComplex*)xf, (cuDoubleComplex*)yf, m, lr1, lda, msp, lmsp, sp8, maxfft, mproc);
  });
}

// CuUser_C_SetBlock2Zero (cuuser_utils.cu:18-29): mltfft_cuda clears the padding of its output with it
void ref_setblock2zero(void* a, int trans_is_n, int N, int M, int LDBX, int LDBY) {
  for_each_thread(LDBX, LDBY, 1, [&] {
    if (trans_is_n) CuUser_Kernel_SetBlock2Zero((cuDoubleComplex*)a, N, M, LDBX, LDBY);
    else CuUser_Kernel_SetBlock2Zero((cuDoubleComplex*)a, M, N, LDBX, LDBY);
  });
}

const char* ref_source(void) { return REF_KERNELS_CU; }

}  // extern "C"
