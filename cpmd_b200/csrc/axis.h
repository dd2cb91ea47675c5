// Per-mesh-length kernel launchers.  One translation unit per supported length (axis_tu.cu
// compiled with -DCPB_N=.. -DCPB_R1=.. -DCPB_R2=..) fills one AxisKernels; registry.cu collects
// them.  The x, y and z launchers of a plan may come from three different lengths.
#pragma once
#include "kernels.h"

namespace cpb {

#ifndef CPB_B
#define CPB_B 8    // batch columns (consecutive x) per block in the y/z passes: 8*16 B = 128 B rows
#endif
#ifndef CPB_SL
#define CPB_SL 16  // ray slots per block in the x pass (a tile of rays plus their mirrors)
#endif

struct AxisKernels {
  int n, r1, r2;
  int b, sl;
  void (*x_inv)(cudaStream_t, const cplx* c0, long ldc, cplx* T1, const PlanDev&, const PairDev&,
                int npair);
  void (*x_fwd)(cudaStream_t, const cplx* T1, const cplx* c0, cplx* c2, long ldc, const PlanDev&,
                const PairDev&, int npair, bool accumulate);
  void (*y_inv)(cudaStream_t, const cplx* T1, cplx* T2, const PlanDev&, int npair);
  void (*y_fwd)(cudaStream_t, const cplx* T2, cplx* T1, const PlanDev&, int npair);
  void (*z_rho)(cudaStream_t, const cplx* T2, double* rho, const PlanDev&, const PairDev&, int npair);
  void (*z_vpsi)(cudaStream_t, cplx* T2, const double* vpot, const PlanDev&, int npair);
};

const AxisKernels* find_axis_kernels(int n);
int num_axis_kernels();
const AxisKernels* axis_kernels_at(int i);

}  // namespace cpb
