// Per-mesh-length kernel launchers.  One translation unit per supported length (axis_tu.cu
// compiled with -DCPB_N=.. -DCPB_R1=.. -DCPB_R2=..) fills one AxisKernels; registry.cu collects
// them.  The x, y and z launchers of a plan may come from three different lengths.
#pragma once
#include "kernels.h"
#include "kernels_zw.h"
#include "kernels_xw.h"

namespace cpb {

#ifndef CPB_B
#define CPB_B 8    // batch columns (consecutive x) per block in the y/z passes: 8*16 B = 128 B rows
#endif
#ifndef CPB_SL
#define CPB_SL 8   // consecutive rays per block in the x pass
#endif

struct AxisKernels {
  int n, r1, r2;
  int b, sl;
  int klo, khi;  // band-pruned k range of the first radix pass: index band must lie in [r2*klo, r2*khi)
  // x passes on the band-ray storage G (see kernels.h): `ppg` = pairs per block (pair groups in
  // grid.y), `half` = band-pruned instantiation
  void (*x_inv)(cudaStream_t, const cplx* c0, long ldc, cplx* T1, const PlanDev&, const PairDev&, int npair,
                int ppg, bool half);
  // same with every coefficient scaled by +-gk[3*ig] (gradient component; tauofr / vtaupsi)
  void (*x_inv_gk)(cudaStream_t, const cplx* c0, long ldc, cplx* T1, const PlanDev&, const PairDev&, int npair,
                   int ppg, bool half, const double* gk);
  void (*x_fwd)(cudaStream_t, const cplx* T1, cplx* G, const PlanDev&, int npair, int ppg, bool half);
  // mirror-pair x passes of the Gamma-point hot path (kernels.h): a block owns rays and their mirrors.
  // x_inv_m: kin_part != nullptr also accumulates the kin_energy / dotp partials, [pair][block][4];
  // x_fwd_m: forward x pass fused with the unpack + kinetic + c2 update (acc: c2 += result)
  void (*x_inv_m)(cudaStream_t, const cplx* c0, long ldc, cplx* T1, const PlanDev&, const PairDev&, int npair,
                  int ppg, bool half, double* kin_part, int geq0);
  void (*x_fwd_m)(cudaStream_t, const cplx* T1, const cplx* c0, cplx* c2, long ldc, const PlanDev&, const PairDev&,
                  int npair, int ppg, bool half, bool acc);
  // y/z passes work on one chunk of x tiles [xt0, xt0+nxc) (T2 holds that chunk only); `half`
  // selects the band-pruned instantiation (KRange), `ppg` = pairs per block (pair groups in grid.z)
  void (*y_inv)(cudaStream_t, const cplx* T1, cplx* T2, const PlanDev&, int npair, int xt0, int nxc,
                int ppg, bool half);
  void (*y_fwd)(cudaStream_t, const cplx* T2, cplx* T1, const PlanDev&, int npair, int xt0, int nxc,
                int ppg, bool half);
  void (*z_rho)(cudaStream_t, const cplx* T2, double* rho, const PlanDev&, const PairDev&, int npair,
                int xt0, int nxc, bool half);
  void (*z_vpsi)(cudaStream_t, cplx* T2, const double* vpot, const PlanDev&, int npair, int xt0, int nxc,
                 int ppg, bool half);
  // warp-autonomous mirror-pair x passes (kernels_xw.h); null if the length has no CPB_XW factorisation.
  // Band-pruned only: usable when the x band lies inside [8*xw_klo, 8*xw_khi).  kin partials: [pair][units][4]
  void (*x_inv_w)(cudaStream_t, const cplx* c0, long ldc, cplx* T1, const PlanDev&, const PairDev&, int npair,
                  int ppg, double* kin_part, int geq0);
  void (*x_fwd_w)(cudaStream_t, const cplx* T1, const cplx* c0, cplx* c2, long ldc, const PlanDev&, const PairDev&,
                  int npair, int ppg, bool acc);
  int xw_ra, xw_klo, xw_khi;
  int xw_rays;            // rays per warp (each warp also owns their mirrors)
  int xw_warps;           // warps per block
  int xw_inv_blocks, xw_fwd_blocks;  // occupancy the kernels are compiled for
  // warp-autonomous z passes (kernels_zw.h); null if the length has no CPB_ZW factorisation.  Band-pruned
  // only: usable when the z band lies inside [zw_rb*zw_klo, zw_rb*zw_khi)
  void (*z_rho_w)(cudaStream_t, const cplx* T2, double* rho, const PlanDev&, const PairDev&, int npair,
                  int xt0, int nxc);
  void (*z_vpsi_w)(cudaStream_t, cplx* T2, const double* vpot, const PlanDev&, int npair, int xt0, int nxc,
                   int ppg);
  int zw_ra, zw_rb, zw_klo, zw_khi;
  int zw_units_per_row;   // warps (column groups) per 128-byte row of a tile
  int zw_warps;           // warps per block
  int zw_blocks_per_sm;   // occupancy the warp kernels are compiled for
  // dense transforms (real fields on the density-cutoff sphere): z passes with phasen
  void (*z_fwd_real)(cudaStream_t, const double* fre, const double* fim, cplx* T2, const PlanDev&, int xt0, int nxc,
                     bool half, const double* mul, double scale);
  void (*z_inv_real)(cudaStream_t, const cplx* T2, double* ore, double* oim, const PlanDev&, int xt0, int nxc,
                     bool acc, bool half);
  int yz_blocks_per_sm;  // occupancy the y/z kernels are compiled for
  int x_inv_blocks, x_fwd_blocks;  // same for the x kernels
  int x_inv_m_blocks, x_fwd_m_blocks;
};

const AxisKernels* find_axis_kernels(int n);
int num_axis_kernels();
const AxisKernels* axis_kernels_at(int i);

}  // namespace cpb
