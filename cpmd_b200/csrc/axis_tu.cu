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

template <class K>
void allow_smem(K kern, size_t bytes) {
#if !defined(CPB_EMULATE)
  if (bytes > 48 * 1024) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  }
#else
  (void)kern;
  (void)bytes;
#endif
}

constexpr size_t kSmemX = (size_t)N * (SL + 1) * sizeof(cplx);
constexpr size_t kSmemYZ = (size_t)N * B * sizeof(cplx);

void x_inv(cudaStream_t st, const cplx* c0, long ldc, cplx* T1, const PlanDev& pd, const PairDev& pr,
           int npair) {
  auto k = k_x_inv<R1, R2, SL>;
  allow_smem(k, kSmemX);
  CPB_LAUNCH(k, dim3(pd.ntiles, npair), dim3(SL * RM), kSmemX, st, c0, ldc, T1, pd, pr);
}

void x_fwd(cudaStream_t st, const cplx* T1, const cplx* c0, cplx* c2, long ldc, const PlanDev& pd,
           const PairDev& pr, int npair, bool accumulate) {
  if (accumulate) {
    auto k = k_x_fwd<R1, R2, SL, true>;
    allow_smem(k, kSmemX);
    CPB_LAUNCH(k, dim3(pd.ntiles, npair), dim3(SL * RM), kSmemX, st, T1, c0, c2, ldc, pd, pr);
  } else {
    auto k = k_x_fwd<R1, R2, SL, false>;
    allow_smem(k, kSmemX);
    CPB_LAUNCH(k, dim3(pd.ntiles, npair), dim3(SL * RM), kSmemX, st, T1, c0, c2, ldc, pd, pr);
  }
}

void y_inv(cudaStream_t st, const cplx* T1, cplx* T2, const PlanDev& pd, int npair) {
  auto k = k_y_inv<R1, R2, B>;
  allow_smem(k, kSmemYZ);
  CPB_LAUNCH(k, dim3((pd.n1 + B - 1) / B, pd.nzb, npair), dim3(B * RM), kSmemYZ, st, T1, T2, pd);
}

void y_fwd(cudaStream_t st, const cplx* T2, cplx* T1, const PlanDev& pd, int npair) {
  auto k = k_y_fwd<R1, R2, B>;
  allow_smem(k, kSmemYZ);
  CPB_LAUNCH(k, dim3((pd.n1 + B - 1) / B, pd.nzb, npair), dim3(B * RM), kSmemYZ, st, T2, T1, pd);
}

void z_rho(cudaStream_t st, const cplx* T2, double* rho, const PlanDev& pd, const PairDev& pr,
           int npair) {
  auto k = k_z_rho<R1, R2, B>;
  allow_smem(k, kSmemYZ);
  CPB_LAUNCH(k, dim3((pd.n1 + B - 1) / B, pd.n2), dim3(B * RM), kSmemYZ, st, T2, rho, pd, pr, npair);
}

void z_vpsi(cudaStream_t st, cplx* T2, const double* vpot, const PlanDev& pd, int npair) {
  auto k = k_z_vpsi<R1, R2, B>;
  allow_smem(k, kSmemYZ);
  CPB_LAUNCH(k, dim3((pd.n1 + B - 1) / B, pd.n2), dim3(B * RM), kSmemYZ, st, T2, vpot, pd, npair);
}

const AxisKernels kTable = {N, R1, R2, B, SL, x_inv, x_fwd, y_inv, y_fwd, z_rho, z_vpsi};

}  // namespace

#define CPB_CAT2(a, b) a##b
#define CPB_CAT(a, b) CPB_CAT2(a, b)
const AxisKernels* CPB_CAT(axis_kernels_n, CPB_N)() { return &kTable; }

}  // namespace cpb
