// One translation unit per mesh length: nvcc -DCPB_N=192 -DCPB_R1=16 -DCPB_R2=12 -c axis_tu.cu
// Instantiates the six pipeline kernels for N = R1*R2 and exports their launchers.
#include "axis.h"

#if !defined(CPB_N) || !defined(CPB_R1) || !defined(CPB_R2)
#error "compile with -DCPB_N=<n> -DCPB_R1=<r1> -DCPB_R2=<r2>"
#endif
static_assert(CPB_R1 * CPB_R2 == CPB_N, "R1*R2 must equal N");

namespace cpb {
namespace {

constexpr int R1 = CPB_R1, R2 = CPB_R2, N = CPB_N, B = CPB_B, SL = CPB_SL;
constexpr int RM = R1 > R2 ? R1 : R2;

// opt in to more than the default 48 KB (static + dynamic) once per kernel instantiation
template <class K>
void allow_smem(K kern, size_t bytes) {
#if !defined(CPB_EMULATE)
  static bool done = false;  // one static per K (each kernel instantiation is a distinct type? no:
                             // same function-pointer type may be shared, so key on the pointer)
  static K last = nullptr;
  if (!done || last != kern) {
    if (bytes > 32 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    done = true;
    last = kern;
  }
#else
  (void)kern;
  (void)bytes;
#endif
}

constexpr size_t kSmemYZ = (size_t)N * B * sizeof(cplx);

template <bool HALF>
void x_inv_t(cudaStream_t st, const cplx* c0, long ldc, cplx* T1, const PlanDev& pd, const PairDev& pr,
             int npair, int ppg) {
  using C = XCfg<R1, R2, SL, HALF>;
  auto k = k_x_inv<R1, R2, SL, B, HALF>;
  allow_smem(k, C::SMEM);
  CPB_LAUNCH(k, dim3(pd.ntiles, (npair + ppg - 1) / ppg), dim3(C::NT), C::SMEM, st, c0, ldc, T1, pd, pr, npair, ppg);
}
void x_inv(cudaStream_t st, const cplx* c0, long ldc, cplx* T1, const PlanDev& pd, const PairDev& pr,
           int npair, int ppg, bool half) {
  if (half) x_inv_t<true>(st, c0, ldc, T1, pd, pr, npair, ppg);
  else x_inv_t<false>(st, c0, ldc, T1, pd, pr, npair, ppg);
}

template <bool HALF, bool ACC>
void x_fwd_t(cudaStream_t st, const cplx* T1, const cplx* c0, cplx* c2, long ldc, const PlanDev& pd,
             const PairDev& pr, int npair, int ppg) {
  using C = XCfg<R1, R2, SL, HALF>;
  auto k = k_x_fwd<R1, R2, SL, B, HALF, ACC>;
  allow_smem(k, C::SMEM);
  CPB_LAUNCH(k, dim3(pd.ntiles, (npair + ppg - 1) / ppg), dim3(C::NT), C::SMEM, st, T1, c0, c2, ldc, pd, pr, npair, ppg);
}
void x_fwd(cudaStream_t st, const cplx* T1, const cplx* c0, cplx* c2, long ldc, const PlanDev& pd,
           const PairDev& pr, int npair, int ppg, bool half, bool accumulate) {
  if (half) {
    if (accumulate) x_fwd_t<true, true>(st, T1, c0, c2, ldc, pd, pr, npair, ppg);
    else x_fwd_t<true, false>(st, T1, c0, c2, ldc, pd, pr, npair, ppg);
  } else {
    if (accumulate) x_fwd_t<false, true>(st, T1, c0, c2, ldc, pd, pr, npair, ppg);
    else x_fwd_t<false, false>(st, T1, c0, c2, ldc, pd, pr, npair, ppg);
  }
}

constexpr size_t kSmemYZ2 = 2 * kSmemYZ;  // double-buffered exchange

template <bool HALF>
void y_inv_t(cudaStream_t st, const cplx* T1, cplx* T2, const PlanDev& pd, int npair, int xt0, int nxc, int ppg) {
  auto k = k_y_inv<R1, R2, B, HALF>;
  allow_smem(k, kSmemYZ2);
  CPB_LAUNCH(k, dim3(nxc, pd.nzb, (npair + ppg - 1) / ppg), dim3(B * RM), kSmemYZ2, st, T1, T2, pd, xt0, npair, ppg);
}
void y_inv(cudaStream_t st, const cplx* T1, cplx* T2, const PlanDev& pd, int npair, int xt0, int nxc, int ppg,
           bool half) {
  if (half) y_inv_t<true>(st, T1, T2, pd, npair, xt0, nxc, ppg);
  else y_inv_t<false>(st, T1, T2, pd, npair, xt0, nxc, ppg);
}

template <bool HALF>
void y_fwd_t(cudaStream_t st, const cplx* T2, cplx* T1, const PlanDev& pd, int npair, int xt0, int nxc, int ppg) {
  auto k = k_y_fwd<R1, R2, B, HALF>;
  allow_smem(k, kSmemYZ2);
  CPB_LAUNCH(k, dim3(nxc, pd.nzb, (npair + ppg - 1) / ppg), dim3(B * RM), kSmemYZ2, st, T2, T1, pd, xt0, npair, ppg);
}
void y_fwd(cudaStream_t st, const cplx* T2, cplx* T1, const PlanDev& pd, int npair, int xt0, int nxc, int ppg,
           bool half) {
  if (half) y_fwd_t<true>(st, T2, T1, pd, npair, xt0, nxc, ppg);
  else y_fwd_t<false>(st, T2, T1, pd, npair, xt0, nxc, ppg);
}

template <bool HALF>
void z_rho_t(cudaStream_t st, const cplx* T2, double* rho, const PlanDev& pd, const PairDev& pr, int npair,
             int xt0, int nxc) {
  auto k = k_z_rho<R1, R2, B, HALF>;
  allow_smem(k, kSmemYZ2);
  CPB_LAUNCH(k, dim3(nxc, pd.n2), dim3(B * RM), kSmemYZ2, st, T2, rho, pd, pr, npair, xt0);
}
void z_rho(cudaStream_t st, const cplx* T2, double* rho, const PlanDev& pd, const PairDev& pr, int npair,
           int xt0, int nxc, bool half) {
  if (half) z_rho_t<true>(st, T2, rho, pd, pr, npair, xt0, nxc);
  else z_rho_t<false>(st, T2, rho, pd, pr, npair, xt0, nxc);
}

template <bool HALF>
void z_vpsi_t(cudaStream_t st, cplx* T2, const double* vpot, const PlanDev& pd, int npair, int xt0, int nxc,
              int ppg) {
  auto k = k_z_vpsi<R1, R2, B, HALF>;
  allow_smem(k, kSmemYZ2);
  CPB_LAUNCH(k, dim3(nxc, pd.n2, (npair + ppg - 1) / ppg), dim3(B * RM), kSmemYZ2, st, T2, vpot, pd, xt0, npair, ppg);
}
void z_vpsi(cudaStream_t st, cplx* T2, const double* vpot, const PlanDev& pd, int npair, int xt0, int nxc,
            int ppg, bool half) {
  if (half) z_vpsi_t<true>(st, T2, vpot, pd, npair, xt0, nxc, ppg);
  else z_vpsi_t<false>(st, T2, vpot, pd, npair, xt0, nxc, ppg);
}

const AxisKernels kTable = {N, R1, R2, B, SL, KRange<R1, true>::lo, KRange<R1, true>::hi,
                            x_inv, x_fwd, y_inv, y_fwd, z_rho, z_vpsi, YZBlocks<R1, R2>::v,
                            XCfg<R1, R2, SL, true>::NT, XCfg<R1, R2, SL, true>::EPT, XCfg<R1, R2, SL, false>::EPT};

}  // namespace

#define CPB_CAT2(a, b) a##b
#define CPB_CAT(a, b) CPB_CAT2(a, b)
const AxisKernels* CPB_CAT(axis_kernels_n, CPB_N)() { return &kTable; }

}  // namespace cpb
