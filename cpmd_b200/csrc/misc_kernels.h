// Small reduction kernels (included by cpb200.cu only).
#pragma once
#include "cpb_defs.h"

namespace cpb {

// ---------------------------------------------------------------------------------------------
// G-space reductions of kin_energy (kin_energy_utils.mod.F90:62-110) and dotp
// (dotp_utils.mod.F90:26-53): one block per state, fixed-order tree reduction (bit-stable).
// out[2*i] = sum_G hg |c|^2 ; out[2*i+1] = dotp(c,c).   block = 256
// ---------------------------------------------------------------------------------------------
CPB_GLOBAL k_kin_energy(const cplx* CPB_RESTRICT c0, long ldc, int first_state, int ngw, int geq0,
                        const double* CPB_RESTRICT hg, double* CPB_RESTRICT out) {
  CPB_DYN_SMEM(double, red);  // 2*256
  const int tid = threadIdx.x;
  const int st = blockIdx.x;
  const cplx* c = c0 + (size_t)(first_state + st) * ldc;
  double sk = 0.0, sd = 0.0;
  for (int ig = tid; ig < ngw; ig += 256) {
    const cplx a = c[ig];
    const double m = a.x * a.x + a.y * a.y;
    sk += hg[ig] * m;
    if (ig == 0) {
      sd += geq0 ? a.x * a.x : 2.0 * m;
    } else {
      sd += 2.0 * m;
    }
  }
  red[tid] = sk;
  red[256 + tid] = sd;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) {
      red[tid] += red[tid + s];
      red[256 + tid] += red[256 + tid + s];
    }
    __syncthreads();
  }
  if (tid == 0) {
    out[2 * st] = red[0];
    out[2 * st + 1] = red[256];
  }
}

// sum of rho over the padded array (pads are zero): per-block partials, fixed order. block = 256
CPB_GLOBAL k_sum(const double* CPB_RESTRICT a, size_t n, double* CPB_RESTRICT partial) {
  CPB_DYN_SMEM(double, red);
  const int tid = threadIdx.x;
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * 256 + tid; i < n; i += (size_t)gridDim.x * 256) s += a[i];
  red[tid] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (tid < k) red[tid] += red[tid + k];
    __syncthreads();
  }
  if (tid == 0) partial[blockIdx.x] = red[0];
}

// ---------------------------------------------------------------------------------------------
// Stage the plane-wave columns of a group of pairs in L2 with sequential TMA bulk prefetches.
// The x passes gather/scatter 16-byte coefficients at random positions of these columns
// (nzhs/indzs order vs |G|^2 order); straight from HBM that access pattern is bound by the DRAM
// row-activation rate, from L2 it is not.  grid = (ceil(ngw*16 / (256*CHUNK)), ncols), block = 256
// ---------------------------------------------------------------------------------------------
constexpr unsigned kPrefetchChunk = 2048;  // bytes per thread

CPB_GLOBAL k_l2_prefetch_cols(const cplx* CPB_RESTRICT base, long ldc, const int* CPB_RESTRICT st1,
                              const int* CPB_RESTRICT st2, int npair, int ngw) {
  const int col = blockIdx.y;  // 0 .. 2*npair-1
  const int s = (col < npair) ? st1[col] : st2[col - npair];
  if (s < 0) return;
  const size_t bytes = (size_t)ngw * sizeof(cplx);
  const size_t off = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * kPrefetchChunk;
  if (off >= bytes) return;
  const size_t n = (bytes - off < kPrefetchChunk) ? bytes - off : kPrefetchChunk;
  l2_prefetch(reinterpret_cast<const char*>(base + (size_t)s * ldc) + off, (unsigned)n);
}

}  // namespace cpb
