// One translation unit per mesh length: nvcc -DCPB_N=192 -DCPB_R1=16 -DCPB_R2=12 -c axis_tu.cu
// Instantiates the six pipeline kernels for N = R1*R2 and exports their launchers.
#include <map>
#include <mutex>
#include <utility>

#include "axis.h"
#include "rt.h"

#if !defined(CPB_N) || !defined(CPB_R1) || !defined(CPB_R2)
#error "compile with -DCPB_N=<n> -DCPB_R1=<r1> -DCPB_R2=<r2>"
#endif
static_assert(CPB_R1 * CPB_R2 == CPB_N, "R1*R2 must equal N");

namespace cpb {
namespace {

constexpr int R1 = CPB_R1, R2 = CPB_R2, N = CPB_N, B = CPB_B, SL = CPB_SL;
constexpr int RM = R1 > R2 ? R1 : R2;

// opt in to more than the default 48 KB (static + dynamic); the size may depend on the plan (band
// width), so remember the largest value set per (device, kernel): the attribute belongs to the
// device that is current when it is set, and one process may hold plans on several devices
template <class K>
void allow_smem(K kern, size_t bytes) {
#if !defined(CPB_EMULATE)
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> set;  // per device and kernel instantiation
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) throw Error(-2, "cudaGetDevice failed");
  std::lock_guard<std::mutex> lock(mu);
  size_t& cur = set[std::make_pair(dev, reinterpret_cast<const void*>(kern))];
  if (bytes > cur) {
    if (bytes > 227 * 1024) throw Error(-4, "kernel needs more shared memory than an SM has (mesh too large)");
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
      throw Error(-2, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed");
    cur = bytes;
  }
#else
  (void)kern;
  (void)bytes;
#endif
}

constexpr size_t kSmemYZ = (size_t)N * B * sizeof(cplx);

template <bool HALF>
void x_inv_t(cudaStream_t st, const cplx* c0, long ldc, cplx* T1, const PlanDev& pd, const PairDev& pr, int npair,
             int ppg) {
  using C = XCfg<R1, R2, SL>;
  auto k = k_x_inv<R1, R2, SL, B, HALF, false>;
  const size_t smem = C::smem_inv(KRange<R1, HALF>::cnt);
  allow_smem(k, smem);
  CPB_LAUNCH_PDL(0, k, dim3(pd.nrp / SL, (npair + ppg - 1) / ppg), dim3(C::NT), smem, st, c0, ldc, T1, pd, pr, npair, ppg,
             (const double*)nullptr);
}
void x_inv(cudaStream_t st, const cplx* c0, long ldc, cplx* T1, const PlanDev& pd, const PairDev& pr, int npair,
           int ppg, bool half) {
  if (half) x_inv_t<true>(st, c0, ldc, T1, pd, pr, npair, ppg);
  else x_inv_t<false>(st, c0, ldc, T1, pd, pr, npair, ppg);
}

template <bool HALF>
void x_inv_gk_t(cudaStream_t st, const cplx* c0, long ldc, cplx* T1, const PlanDev& pd, const PairDev& pr, int npair,
                int ppg, const double* gk) {
  using C = XCfg<R1, R2, SL>;
  auto k = k_x_inv<R1, R2, SL, B, HALF, true>;
  const size_t smem = C::smem_inv(KRange<R1, HALF>::cnt);
  allow_smem(k, smem);
  CPB_LAUNCH_PDL(0, k, dim3(pd.nrp / SL, (npair + ppg - 1) / ppg), dim3(C::NT), smem, st, c0, ldc, T1, pd, pr, npair, ppg, gk);
}
void x_inv_gk(cudaStream_t st, const cplx* c0, long ldc, cplx* T1, const PlanDev& pd, const PairDev& pr, int npair,
              int ppg, bool half, const double* gk) {
  if (half) x_inv_gk_t<true>(st, c0, ldc, T1, pd, pr, npair, ppg, gk);
  else x_inv_gk_t<false>(st, c0, ldc, T1, pd, pr, npair, ppg, gk);
}

template <bool HALF>
void x_fwd_t(cudaStream_t st, const cplx* T1, cplx* G, const PlanDev& pd, int npair, int ppg) {
  using C = XCfg<R1, R2, SL>;
  auto k = k_x_fwd<R1, R2, SL, B, HALF>;
  allow_smem(k, C::SMEM_FWD);
  CPB_LAUNCH_PDL(5, k, dim3(pd.nrp / SL, (npair + ppg - 1) / ppg), dim3(C::NT), C::SMEM_FWD, st, T1, G, pd, npair, ppg);
}
void x_fwd(cudaStream_t st, const cplx* T1, cplx* G, const PlanDev& pd, int npair, int ppg, bool half) {
  if (half) x_fwd_t<true>(st, T1, G, pd, npair, ppg);
  else x_fwd_t<false>(st, T1, G, pd, npair, ppg);
}

// blocks of the mirror-pair x kernels: H = SL/2 rays (+ their mirrors) each
int x_m_blocks(const PlanDev& pd) { return ((pd.nrays + 1) / 2 + SL / 2 - 1) / (SL / 2); }

template <bool HALF, bool KIN>
void x_inv_m_t(cudaStream_t st, const cplx* c0, long ldc, cplx* T1, const PlanDev& pd, const PairDev& pr, int npair,
               int ppg, double* kin_part, int geq0) {
  using C = XCfg<R1, R2, SL>;
  using M = XMCfg<R1, R2, SL, HALF>;
  auto k = k_x_inv_m<R1, R2, SL, B, HALF, KIN>;
  allow_smem(k, M::SMEM_INV);
  CPB_LAUNCH_PDL(0, k, dim3(x_m_blocks(pd), (npair + ppg - 1) / ppg), dim3(C::NT), M::SMEM_INV, st, c0, ldc, T1, pd, pr,
             npair, ppg, kin_part, geq0);
}
void x_inv_m(cudaStream_t st, const cplx* c0, long ldc, cplx* T1, const PlanDev& pd, const PairDev& pr, int npair,
             int ppg, bool half, double* kin_part, int geq0) {
  if (half) {
    if (kin_part) x_inv_m_t<true, true>(st, c0, ldc, T1, pd, pr, npair, ppg, kin_part, geq0);
    else x_inv_m_t<true, false>(st, c0, ldc, T1, pd, pr, npair, ppg, kin_part, geq0);
  } else {
    if (kin_part) x_inv_m_t<false, true>(st, c0, ldc, T1, pd, pr, npair, ppg, kin_part, geq0);
    else x_inv_m_t<false, false>(st, c0, ldc, T1, pd, pr, npair, ppg, kin_part, geq0);
  }
}

template <bool HALF, bool ACC>
void x_fwd_m_t(cudaStream_t st, const cplx* T1, const cplx* c0, cplx* c2, long ldc, const PlanDev& pd,
               const PairDev& pr, int npair, int ppg) {
  using C = XCfg<R1, R2, SL>;
  using M = XMCfg<R1, R2, SL, HALF>;
  auto k = k_x_fwd_m<R1, R2, SL, B, HALF, ACC>;
  allow_smem(k, M::SMEM_FWD);
  CPB_LAUNCH_PDL(5, k, dim3(x_m_blocks(pd), (npair + ppg - 1) / ppg), dim3(C::NT), M::SMEM_FWD, st, T1, c0, c2, ldc, pd, pr,
             npair, ppg);
}
void x_fwd_m(cudaStream_t st, const cplx* T1, const cplx* c0, cplx* c2, long ldc, const PlanDev& pd, const PairDev& pr,
             int npair, int ppg, bool half, bool acc) {
  if (half) {
    if (acc) x_fwd_m_t<true, true>(st, T1, c0, c2, ldc, pd, pr, npair, ppg);
    else x_fwd_m_t<true, false>(st, T1, c0, c2, ldc, pd, pr, npair, ppg);
  } else {
    if (acc) x_fwd_m_t<false, true>(st, T1, c0, c2, ldc, pd, pr, npair, ppg);
    else x_fwd_m_t<false, false>(st, T1, c0, c2, ldc, pd, pr, npair, ppg);
  }
}

constexpr size_t kSmemYZ2 = 2 * kSmemYZ;  // double-buffered exchange
// exchange buffers of the staged kernels: the tuning macro, unless two buffers would leave one block per SM (YZXB)
constexpr int kXBy = (CPB_YINV_XB == 2) ? YZXB<R1, R2, B>::v : CPB_YINV_XB;
constexpr int kXBz = (CPB_ZRHO_XB == 2) ? YZXB<R1, R2, B>::v : CPB_ZRHO_XB;
constexpr int kXBv = YZXB<R1, R2, B>::v;

template <bool HALF>
void y_inv_t(cudaStream_t st, const cplx* T1, cplx* T2, const PlanDev& pd, int npair, int xt0, int nxc, int ppg) {
  auto k = k_y_inv<R1, R2, B, HALF, kXBy>;
  const size_t smem = YZCfg<R1, R2, B>::smem(pd.nyb * B, kXBy);
  allow_smem(k, smem);
  CPB_LAUNCH_PDL(1, k, CPB_Y_ZFAST ? dim3(pd.nzb, nxc, (npair + ppg - 1) / ppg) : dim3(nxc, pd.nzb, (npair + ppg - 1) / ppg),
             dim3(B * RM), smem, st, T1, T2, pd, xt0, npair, ppg);
}
void y_inv(cudaStream_t st, const cplx* T1, cplx* T2, const PlanDev& pd, int npair, int xt0, int nxc, int ppg,
           bool half) {
  if (half) y_inv_t<true>(st, T1, T2, pd, npair, xt0, nxc, ppg);
  else y_inv_t<false>(st, T1, T2, pd, npair, xt0, nxc, ppg);
}

template <bool HALF>
void y_fwd_t(cudaStream_t st, const cplx* T2, cplx* T1, const PlanDev& pd, int npair, int xt0, int nxc, int ppg) {
  // two exchange buffers (+ the private cp.async slots of the next pair's rows, R2 elements per thread, where
  // they do not cost a resident block)
  constexpr size_t slots = (size_t)R2 * B * RM * sizeof(cplx);
  constexpr bool async = CPB_YFWD_ASYNC && (YZBlocks<R1, R2>::v * (kSmemYZ2 + slots + 1024) <= (size_t)228 * 1024);
  auto k = k_y_fwd<R1, R2, B, HALF, async>;
  const size_t smem = kSmemYZ2 + (async ? slots : 0);
  allow_smem(k, smem);
  CPB_LAUNCH_PDL(4, k, CPB_Y_ZFAST ? dim3(pd.nzb, nxc, (npair + ppg - 1) / ppg) : dim3(nxc, pd.nzb, (npair + ppg - 1) / ppg),
             dim3(B * RM), smem, st, T2, T1, pd, xt0, npair, ppg);
}
void y_fwd(cudaStream_t st, const cplx* T2, cplx* T1, const PlanDev& pd, int npair, int xt0, int nxc, int ppg,
           bool half) {
  if (half) y_fwd_t<true>(st, T2, T1, pd, npair, xt0, nxc, ppg);
  else y_fwd_t<false>(st, T2, T1, pd, npair, xt0, nxc, ppg);
}

template <bool HALF>
void z_rho_t(cudaStream_t st, const cplx* T2, double* rho, const PlanDev& pd, const PairDev& pr, int npair,
             int xt0, int nxc) {
  auto k = k_z_rho<R1, R2, B, HALF, kXBz>;
  const size_t smem = YZCfg<R1, R2, B>::smem(pd.nzb * B, kXBz);
  allow_smem(k, smem);
  CPB_LAUNCH_PDL(2, k, dim3(nxc, pd.n2), dim3(B * RM), smem, st, T2, rho, pd, pr, npair, xt0);
}
void z_rho(cudaStream_t st, const cplx* T2, double* rho, const PlanDev& pd, const PairDev& pr, int npair,
           int xt0, int nxc, bool half) {
  if (half) z_rho_t<true>(st, T2, rho, pd, pr, npair, xt0, nxc);
  else z_rho_t<false>(st, T2, rho, pd, pr, npair, xt0, nxc);
}

template <bool HALF>
void z_vpsi_t(cudaStream_t st, cplx* T2, const double* vpot, const PlanDev& pd, int npair, int xt0, int nxc,
              int ppg) {
  auto k = k_z_vpsi<R1, R2, B, HALF, kXBv>;
  const size_t smem = YZCfg<R1, R2, B>::smem(pd.nzb * B, kXBv);
  allow_smem(k, smem);
  CPB_LAUNCH_PDL(3, k, dim3(nxc, pd.n2, (npair + ppg - 1) / ppg), dim3(B * RM), smem, st, T2, vpot, pd, xt0, npair, ppg);
}
void z_vpsi(cudaStream_t st, cplx* T2, const double* vpot, const PlanDev& pd, int npair, int xt0, int nxc,
            int ppg, bool half) {
  if (half) z_vpsi_t<true>(st, T2, vpot, pd, npair, xt0, nxc, ppg);
  else z_vpsi_t<false>(st, T2, vpot, pd, npair, xt0, nxc, ppg);
}

// warp-autonomous mirror-pair x passes (kernels_xw.h), band-pruned instantiation only
constexpr bool kHasXW = XWPick<N>::ra != 0 && B == 8;
template <bool HAS, int DUMMY = 0>
struct XWLaunch {
  static constexpr void (*inv)(cudaStream_t, const cplx*, long, cplx*, const PlanDev&, const PairDev&, int, int, double*,
                               int) = nullptr;
  static constexpr void (*fwd)(cudaStream_t, const cplx*, const cplx*, cplx*, long, const PlanDev&, const PairDev&, int,
                               int, bool) = nullptr;
  static constexpr int ra = 0, klo = 0, khi = 0, rays = 0, warps = 0, inv_blocks = 0, fwd_blocks = 0;
};
template <int DUMMY>
struct XWLaunch<true, DUMMY> {  // partial specialisation: members are only instantiated where used
  static constexpr int ra = XWPick<N>::ra ? XWPick<N>::ra : 8;
  using C = XWCfg<ra, true>;
  static constexpr int klo = C::KR::lo, khi = C::KR::hi, rays = C::H, warps = C::WARPS, inv_blocks = C::MINB_INV,
                       fwd_blocks = C::MINB_FWD;
  static int grid_x(const PlanDev& pd) { return (xw_units<ra, true>(pd.nrays) + warps - 1) / warps; }
  static void inv(cudaStream_t st, const cplx* c0, long ldc, cplx* T1, const PlanDev& pd, const PairDev& pr, int npair,
                  int ppg, double* kin_part, int geq0) {
    if (kin_part) {
      auto k = k_xw_inv<ra, B, true, true>;
      allow_smem(k, C::SMEM_INV);
      CPB_LAUNCH_PDL(0, k, dim3(grid_x(pd), (npair + ppg - 1) / ppg), dim3(C::NT), C::SMEM_INV, st, c0, ldc, T1, pd, pr, npair,
                 ppg, kin_part, geq0);
    } else {
      auto k = k_xw_inv<ra, B, true, false>;
      allow_smem(k, C::SMEM_INV);
      CPB_LAUNCH_PDL(0, k, dim3(grid_x(pd), (npair + ppg - 1) / ppg), dim3(C::NT), C::SMEM_INV, st, c0, ldc, T1, pd, pr, npair,
                 ppg, kin_part, geq0);
    }
  }
  static void fwd(cudaStream_t st, const cplx* T1, const cplx* c0, cplx* c2, long ldc, const PlanDev& pd,
                  const PairDev& pr, int npair, int ppg, bool acc) {
    if (acc) {
      auto k = k_xw_fwd<ra, B, true, true>;
      allow_smem(k, C::SMEM_FWD);
      CPB_LAUNCH_PDL(5, k, dim3(grid_x(pd), (npair + ppg - 1) / ppg), dim3(C::NT), C::SMEM_FWD, st, T1, c0, c2, ldc, pd, pr,
                 npair, ppg);
    } else {
      auto k = k_xw_fwd<ra, B, true, false>;
      allow_smem(k, C::SMEM_FWD);
      CPB_LAUNCH_PDL(5, k, dim3(grid_x(pd), (npair + ppg - 1) / ppg), dim3(C::NT), C::SMEM_FWD, st, T1, c0, c2, ldc, pd, pr,
                 npair, ppg);
    }
  }
};
using XWL = XWLaunch<kHasXW>;

// warp-autonomous z passes (kernels_zw.h), band-pruned instantiation only
using ZP = ZWPick<N>;
constexpr bool kHasZW = ZP::ra != 0;
template <bool HAS, int DUMMY = 0>
struct ZWLaunch {
  static constexpr void (*rho)(cudaStream_t, const cplx*, double*, const PlanDev&, const PairDev&, int, int, int) = nullptr;
  static constexpr void (*vpsi)(cudaStream_t, cplx*, const double*, const PlanDev&, int, int, int, int) = nullptr;
  static constexpr int ra = 0, rb = 0, klo = 0, khi = 0, upr = 0, warps = 0, minb = 0;
};
template <int DUMMY>
struct ZWLaunch<true, DUMMY> {  // partial specialisation: members are only instantiated where used
  static constexpr int ra = ZP::ra ? ZP::ra : 4, rb = ZP::rb ? ZP::rb : 4, L = ZP::l ? ZP::l : 4;
  using C = ZWCfg<ra, rb, L, true>;
  static constexpr int klo = C::KR::lo, khi = C::KR::hi, upr = B / C::CW, warps = C::WARPS, minb = C::MINB;
  static int grid_y(const PlanDev& pd) { return (pd.n2 * upr + warps - 1) / warps; }
  static void rho(cudaStream_t st, const cplx* T2, double* rho_, const PlanDev& pd, const PairDev& pr, int npair,
                  int xt0, int nxc) {
    auto k = k_zw_rho<ra, rb, L, B, true>;
    allow_smem(k, C::SMEM);
    CPB_LAUNCH_PDL(2, k, dim3(nxc, grid_y(pd)), dim3(C::NT), C::SMEM, st, T2, rho_, pd, pr, npair, xt0);
  }
  static void vpsi(cudaStream_t st, cplx* T2, const double* vpot, const PlanDev& pd, int npair, int xt0, int nxc,
                   int ppg) {
    auto k = k_zw_vpsi<ra, rb, L, B, true>;
    allow_smem(k, C::SMEM);
    CPB_LAUNCH_PDL(3, k, dim3(nxc, grid_y(pd), (npair + ppg - 1) / ppg), dim3(C::NT), C::SMEM, st, T2, vpot, pd, xt0, npair,
               ppg);
  }
};
using ZWL = ZWLaunch<kHasZW>;

template <bool HALF>
void z_fwd_real_t(cudaStream_t st, const double* fre, const double* fim, cplx* T2, const PlanDev& pd, int xt0,
                  int nxc, const double* mul, double scale) {
  auto k = k_z_fwd_real<R1, R2, B, HALF>;
  allow_smem(k, kSmemYZ);
  CPB_LAUNCH_PDL(6, k, dim3(nxc, pd.n2), dim3(B * RM), kSmemYZ, st, fre, fim, T2, pd, xt0, mul, scale);
}
void z_fwd_real(cudaStream_t st, const double* fre, const double* fim, cplx* T2, const PlanDev& pd, int xt0, int nxc,
                bool half, const double* mul, double scale) {
  if (half) z_fwd_real_t<true>(st, fre, fim, T2, pd, xt0, nxc, mul, scale);
  else z_fwd_real_t<false>(st, fre, fim, T2, pd, xt0, nxc, mul, scale);
}

template <bool HALF>
void z_inv_real_t(cudaStream_t st, const cplx* T2, double* ore, double* oim, const PlanDev& pd, int xt0, int nxc,
                  bool acc) {
  auto k = k_z_inv_real<R1, R2, B, HALF>;
  allow_smem(k, kSmemYZ);
  CPB_LAUNCH_PDL(6, k, dim3(nxc, pd.n2), dim3(B * RM), kSmemYZ, st, T2, ore, oim, pd, xt0, acc ? 1 : 0);
}
void z_inv_real(cudaStream_t st, const cplx* T2, double* ore, double* oim, const PlanDev& pd, int xt0, int nxc,
                bool acc, bool half) {
  if (half) z_inv_real_t<true>(st, T2, ore, oim, pd, xt0, nxc, acc);
  else z_inv_real_t<false>(st, T2, ore, oim, pd, xt0, nxc, acc);
}

const AxisKernels kTable = {N, R1, R2, B, SL, KRange<R1, true>::lo, KRange<R1, true>::hi,
                            x_inv, x_inv_gk, x_fwd, x_inv_m, x_fwd_m, y_inv, y_fwd, z_rho, z_vpsi,
                            XWL::inv, XWL::fwd, XWL::ra, XWL::klo, XWL::khi, XWL::rays, XWL::warps, XWL::inv_blocks,
                            XWL::fwd_blocks,
                            ZWL::rho, ZWL::vpsi, ZWL::ra, ZWL::rb, ZWL::klo, ZWL::khi, ZWL::upr, ZWL::warps, ZWL::minb,
                            z_fwd_real, z_inv_real,
                            YZBlocks<R1, R2>::v, XCfg<R1, R2, SL>::MINB, XCfg<R1, R2, SL>::MINB_FWD,
                            XMCfg<R1, R2, SL, true>::MINB_INV, XMCfg<R1, R2, SL, true>::MINB_FWD};

}  // namespace

#define CPB_CAT2(a, b) a##b
#define CPB_CAT(a, b) CPB_CAT2(a, b)
const AxisKernels* CPB_CAT(axis_kernels_n, CPB_N)() { return &kTable; }

}  // namespace cpb
