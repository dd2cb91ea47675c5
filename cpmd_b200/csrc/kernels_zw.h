// Warp-autonomous z passes (sm_100a): k_zw_rho / k_zw_vpsi.
//
// Same contract, same memory layout and the same arithmetic per column as k_z_rho / k_z_vpsi
// (kernels.h); what changes is who synchronises with whom.  The block kernels spread one column
// transform over max(R1,R2) threads of a 4-warp block and separate the radix passes with block
// barriers; on a B200 they sit at 60-66 % of the FP64 pipe with the block barrier as the top stall
// (profiles/r02e_full_zrho_192x128.txt): three blocks per SM march in lock step and in two of three
// phases a quarter of the role rows has nothing to do.  Here ONE WARP owns CW columns of a tile for the
// whole pair loop:
//   * N = RA * RB, both multiples of the L = 32/CW role lanes of a column.  Lane (l, col) runs the
//     band-side radix-RA sub-transforms of roles a = l + L i  (i < RB/L) and the real-space-side
//     radix-RB sub-transforms of roles p = l + L j  (j < RA/L): every lane has work in every phase.
//   * the exchange between the two radix passes goes through a per-warp region of shared memory,
//     ordered by __syncwarp() only.  No block barrier, no mbarrier: warps never wait for each other,
//     the scheduler of an SM sub-partition always has independent warps to pick from.
//   * a lane fetches exactly the band elements it transforms (z = a + RB k) with 16-byte cp.async copies
//     into private shared-memory slots, one pair ahead: no registers held by loads in flight and no
//     hand-over between threads (same idea as the gather of k_x_inv).
//   * the band-side transforms are the streaming codelets dft_in_dif / dft_out_dit (codelets.h): the
//     radix-RA transform is consumed in D groups of RA/D, which keeps the live set at (band elements +
//     RA/D) instead of RA complex registers; k_zw_vpsi runs the real-space side in place (the lane that
//     owns row p of the exchange region reads it, multiplies by V, transforms forward and writes the row
//     back).
// Replaces putz + z mltfft + build_density_sum (rhoofr_utils.mod.F90:369-374, density_utils.mod.F90:
// 61-83) and putz + z mltfft + V psi + z mltfft + getz (vpsi_utils.mod.F90:487-493,
// fftmain_utils.mod.F90:100-104,122-127) like the block kernels.
#pragma once
#include "kernels.h"

namespace cpb {

// Factorisations the warp kernels are built for: CPB_ZW(N, RA, RB, L).  RA: radix of the band-side pass,
// RB: radix of the real-space-side pass, L: role lanes per column (32/L columns per warp).  Chosen so that
// N/L accumulators (rho) or potential values (V) plus one band-side working set fit the register budget
// of three 128-thread blocks per SM.  Lengths without an entry keep the block kernels.
template <int N>
struct ZWPick {
  static constexpr int ra = 0, rb = 0, l = 0;
};
#define CPB_ZW(N_, RA_, RB_, L_)                        \
  template <>                                           \
  struct ZWPick<N_> {                                   \
    static constexpr int ra = RA_, rb = RB_, l = L_;    \
  };
CPB_ZW(16, 4, 4, 4)
CPB_ZW(32, 8, 4, 4)
CPB_ZW(48, 12, 4, 4)
CPB_ZW(64, 8, 8, 8)
CPB_ZW(96, 12, 8, 4)
CPB_ZW(128, 16, 8, 8)
CPB_ZW(144, 12, 12, 4)
CPB_ZW(192, 24, 8, 8)
CPB_ZW(256, 16, 16, 8)
#undef CPB_ZW

template <int RA, int RB, int L, bool HALF>
struct ZWCfg {
  static constexpr int N = RA * RB;
  static constexpr int CW = 32 / L;       // columns per warp
  static constexpr int NA = RB / L;       // band-side sub-transforms (radix RA) per lane
  static constexpr int NB = RA / L;       // real-space-side sub-transforms (radix RB) per lane
  static constexpr int D = (RA % 3 == 0) ? 3 : ((RA % 2 == 0) ? 2 : 1);
  using KR = KRange<RA, HALF>;
  static constexpr int KC = KR::cnt;
  static constexpr int WARPS = 4;         // per block
  static constexpr int NT = 32 * WARPS;
  static constexpr int EX_ELEMS = N * CW;          // per warp: exchange region
  static constexpr int ST_ELEMS = NA * KC * 32;    // per warp: private slots of the band elements
  static constexpr size_t SMEM = (size_t)(WARPS * (EX_ELEMS + ST_ELEMS) + N) * sizeof(cplx);
  static constexpr int MINB = (SMEM * 3 + 3 * 1024 <= 228 * 1024) ? 3 : ((SMEM * 2 + 2 * 1024 <= 228 * 1024) ? 2 : 1);
  static_assert(RA % L == 0 && RB % L == 0, "both radices must be multiples of the role lanes");
  // position of (p, a) in units of CW elements.  CW = 4: a quarter warp (the unit of a 16-byte shared-memory
  // access) spans two role lanes; the XOR puts rows p and p+1 into different halves of a 128-byte bank line
  // when the lanes run over p, and is a permutation inside a row when they run over a.
  CPB_HD static constexpr int pos(int p, int a) { return p * RB + ((CW == 4) ? (a ^ (p & 1)) : a); }
};

// ---------------------------------------------------------------------------------------------
// Common pieces.  lane = l * CW + col.
// ---------------------------------------------------------------------------------------------
template <int RA, int RB, int L, bool HALF>
struct ZWLane {
  using C = ZWCfg<RA, RB, L, HALF>;
  using KR = typename C::KR;
  static constexpr int CW = C::CW, NA = C::NA, NB = C::NB, KC = C::KC, D = C::D;

  // issue the copies of one pair's band elements of this lane: src points at (zr = 0, my column), row
  // pitch B elements; out-of-band slots were zeroed once and are never written
  template <int B>
  static CPB_D void gather(cplx* ST, int lane, int l, const cplx* src, int zlo, int nzb) {
    static_for<0, NA>([&](auto ii) {
      constexpr int i = decltype(ii)::value;
      static_for<0, KC>([&](auto kk) {
        constexpr int j = decltype(kk)::value;
        const int zr = (l + L * i) + RB * (KR::lo + j) - zlo;
        if (zr >= 0 && zr < nzb) cp_async16(&ST[(i * KC + j) * 32 + lane], src + (size_t)zr * B);
      });
    });
    cp_async_commit();
  }
  static CPB_D void zero_slots(cplx* ST, int lane, int l, int zlo, int nzb) {
    static_for<0, NA>([&](auto ii) {
      constexpr int i = decltype(ii)::value;
      static_for<0, KC>([&](auto kk) {
        constexpr int j = decltype(kk)::value;
        const int zr = (l + L * i) + RB * (KR::lo + j) - zlo;
        if (!(zr >= 0 && zr < nzb)) ST[(i * KC + j) * 32 + lane] = mk(0.0, 0.0);
      });
    });
  }

  // band-side inverse pass of sub-transform i: radix-RA of the slots, twiddle w^(a p), store row-wise into
  // the exchange region
  template <int I>
  static CPB_D void inv_band(const cplx* ST, cplx* EX, const cplx* TW, int lane, int l, int col) {
    cplx v[RA];
    static_for<0, KC>([&](auto kk) {
      constexpr int j = decltype(kk)::value;
      v[KR::lo + j] = ST[(I * KC + j) * 32 + lane];
    });
    const int a = l + L * I;
    dft_in_dif<RA, D, true, KR::lo, KR::hi>(v, [&](auto pp, cplx o) {
      constexpr int p = decltype(pp)::value;
      const int e = p * RB + ((CW == 4) ? (a ^ (p & 1)) : a);
      if constexpr (p != 0) o = cmul(o, TW[e]);
      EX[e * CW + col] = o;
    });
  }
};

// ---------------------------------------------------------------------------------------------
// z pass of rhoofr.  grid = (x tiles of the chunk, ceil(n2 * halves / WARPS)), block = 128:
// warp w of block (bx, by) owns unit u = by * WARPS + w -> y = u / halves, column group u % halves of
// x tile bx (halves = B / CW column groups per 128-byte row).
// ---------------------------------------------------------------------------------------------
template <int RA, int RB, int L, int B, bool HALF>
CPB_GLOBAL CPB_LAUNCH_BOUNDS((ZWCfg<RA, RB, L, HALF>::NT), (ZWCfg<RA, RB, L, HALF>::MINB))
    k_zw_rho(const cplx* CPB_RESTRICT T2, double* rho, PlanDev pd, PairDev pr, int npair, int xt0) {
  using C = ZWCfg<RA, RB, L, HALF>;
  using LN = ZWLane<RA, RB, L, HALF>;
  constexpr int N = C::N, CW = C::CW, NB = C::NB, NA = C::NA;
  static_assert(B % CW == 0, "a row of B columns splits into whole column groups");
  constexpr int HV = B / CW;
  CPB_DYN_SMEM(cplx, S);
  const int tid = threadIdx.x;
  const int w = tid >> 5, lane = tid & 31;
  const int col = lane % CW, l = lane / CW;
  cplx* EX = S + (size_t)w * C::EX_ELEMS;
  cplx* ST = S + (size_t)C::WARPS * C::EX_ELEMS + (size_t)w * C::ST_ELEMS;
  cplx* TW = S + (size_t)C::WARPS * (C::EX_ELEMS + C::ST_ELEMS);
  pdl_trigger();
  for (int i = tid; i < N; i += C::NT) {
    const int p = i / RB, a = i % RB;
    TW[C::pos(p, a)] = pd.tw3[a * p];
  }
  pdl_wait();  // T2, rho / V come from preceding kernels
  __syncthreads();  // the only block barrier: twiddles visible
  const int unit = blockIdx.y * C::WARPS + w;
  const int y = unit / HV, hv = unit % HV;
  if (y >= pd.n2) return;  // warp-uniform
  const int xtc = blockIdx.x, nxc = gridDim.x;
  const int x = (xt0 + xtc) * B + hv * CW + col;
  const bool xok = x < pd.n1;
  const int zlo = pd.zlo, nzb = pd.nzb;
  const size_t pstride = (size_t)nxc * pd.n2 * nzb * B;
  const cplx* src = T2 + ((size_t)xtc * pd.n2 + y) * nzb * B + hv * CW + col;
  LN::zero_slots(ST, lane, l, zlo, nzb);
  if (npair > 0) LN::template gather<B>(ST, lane, l, src, zlo, nzb);
  // the accumulators start from rho itself (one read-modify-write of rho per batch)
  double acc[NB * RB];
  static_for<0, NB>([&](auto jj) {
    constexpr int j = decltype(jj)::value;
    static_for<0, RB>([&](auto qq) {
      constexpr int q = decltype(qq)::value;
      const int z = (l + L * j) + RA * q;
      acc[j * RB + q] = xok ? rho[((size_t)z * pd.kr2 + y) * pd.kr1 + x] : 0.0;
    });
  });
  for (int pair = 0; pair < npair; ++pair) {
    const double ca = __ldg(&pr.ca[pair]), cb = __ldg(&pr.cb[pair]);
    cp_async_wait_all();
    static_for<0, NA>([&](auto ii) {
      constexpr int i = decltype(ii)::value;
      LN::template inv_band<i>(ST, EX, TW, lane, l, col);
      // the slots of the last sub-transform have been read: the next pair's copies may land
      if constexpr (i == NA - 1) {
        if (pair + 1 < npair) LN::template gather<B>(ST, lane, l, src + (size_t)(pair + 1) * pstride, zlo, nzb);
      }
    });
    __syncwarp();
    static_for<0, NB>([&](auto jj) {
      constexpr int j = decltype(jj)::value;
      const int p = l + L * j;
      cplx u[RB];
      static_for<0, RB>([&](auto aa) {
        constexpr int a = decltype(aa)::value;
        u[a] = EX[(p * RB + ((CW == 4) ? (a ^ (p & 1)) : a)) * CW + col];
      });
      dft<RB, true>(u);
      static_for<0, RB>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        acc[j * RB + q] += ca * (u[q].x * u[q].x) + cb * (u[q].y * u[q].y);
      });
    });
    __syncwarp();  // the exchange region is single buffered
  }
  if (xok) {
    static_for<0, NB>([&](auto jj) {
      constexpr int j = decltype(jj)::value;
      static_for<0, RB>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        const int z = (l + L * j) + RA * q;
        rho[((size_t)z * pd.kr2 + y) * pd.kr1 + x] = acc[j * RB + q];
      });
    });
  }
}

// ---------------------------------------------------------------------------------------------
// z pass of vpsi: z-inverse, multiply by V(r), z-forward, band stored back in place.
// grid = (x tiles of the chunk, ceil(n2 * halves / WARPS), pair groups)
// ---------------------------------------------------------------------------------------------
template <int RA, int RB, int L, int B, bool HALF>
CPB_GLOBAL CPB_LAUNCH_BOUNDS((ZWCfg<RA, RB, L, HALF>::NT), (ZWCfg<RA, RB, L, HALF>::MINB))
    k_zw_vpsi(cplx* T2, const double* CPB_RESTRICT vpot, PlanDev pd, int xt0, int npair, int ppg) {
  using C = ZWCfg<RA, RB, L, HALF>;
  using LN = ZWLane<RA, RB, L, HALF>;
  using KR = typename C::KR;
  constexpr int N = C::N, CW = C::CW, NB = C::NB, NA = C::NA, D = C::D;
  static_assert(B % CW == 0, "a row of B columns splits into whole column groups");
  constexpr int HV = B / CW;
  CPB_DYN_SMEM(cplx, S);
  const int tid = threadIdx.x;
  const int w = tid >> 5, lane = tid & 31;
  const int col = lane % CW, l = lane / CW;
  cplx* EX = S + (size_t)w * C::EX_ELEMS;
  cplx* ST = S + (size_t)C::WARPS * C::EX_ELEMS + (size_t)w * C::ST_ELEMS;
  cplx* TW = S + (size_t)C::WARPS * (C::EX_ELEMS + C::ST_ELEMS);
  pdl_trigger();
  for (int i = tid; i < N; i += C::NT) {
    const int p = i / RB, a = i % RB;
    TW[C::pos(p, a)] = pd.tw3[a * p];
  }
  pdl_wait();  // T2, rho / V come from preceding kernels
  __syncthreads();  // the only block barrier: twiddles visible
  const int unit = blockIdx.y * C::WARPS + w;
  const int y = unit / HV, hv = unit % HV;
  if (y >= pd.n2) return;  // warp-uniform
  const int xtc = blockIdx.x, nxc = gridDim.x;
  const int x = (xt0 + xtc) * B + hv * CW + col;
  const bool xok = x < pd.n1;
  const int zlo = pd.zlo, nzb = pd.nzb;
  const int p0 = blockIdx.z * ppg;
  const int p1 = (p0 + ppg < npair) ? p0 + ppg : npair;
  const size_t pstride = (size_t)nxc * pd.n2 * nzb * B;
  cplx* tile = T2 + ((size_t)xtc * pd.n2 + y) * nzb * B + hv * CW + col;
  LN::zero_slots(ST, lane, l, zlo, nzb);
  if (p0 < p1) LN::template gather<B>(ST, lane, l, tile + (size_t)p0 * pstride, zlo, nzb);
  double vv[NB * RB];
  static_for<0, NB>([&](auto jj) {
    constexpr int j = decltype(jj)::value;
    static_for<0, RB>([&](auto qq) {
      constexpr int q = decltype(qq)::value;
      const int z = (l + L * j) + RA * q;
      vv[j * RB + q] = xok ? __ldg(&vpot[((size_t)z * pd.kr2 + y) * pd.kr1 + x]) : 0.0;
    });
  });
  for (int pair = p0; pair < p1; ++pair) {
    cp_async_wait_all();
    static_for<0, NA>([&](auto ii) {
      constexpr int i = decltype(ii)::value;
      LN::template inv_band<i>(ST, EX, TW, lane, l, col);
      if constexpr (i == NA - 1) {
        if (pair + 1 < p1) LN::template gather<B>(ST, lane, l, tile + (size_t)(pair + 1) * pstride, zlo, nzb);
      }
    });
    __syncwarp();
    // real-space side, in place: row p of the exchange region belongs to this lane alone
    static_for<0, NB>([&](auto jj) {
      constexpr int j = decltype(jj)::value;
      const int p = l + L * j;
      cplx u[RB];
      static_for<0, RB>([&](auto aa) {
        constexpr int a = decltype(aa)::value;
        u[a] = EX[(p * RB + ((CW == 4) ? (a ^ (p & 1)) : a)) * CW + col];
      });
      dft<RB, true>(u);
      static_for<0, RB>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        u[q].x *= vv[j * RB + q];
        u[q].y *= vv[j * RB + q];
      });
      // forward transform, mirrored factorisation: radix RB over q, twiddle conj(w^(p s)), row p again
      dft<RB, false>(u);
      static_for<0, RB>([&](auto ss) {
        constexpr int s = decltype(ss)::value;
        const int e = p * RB + ((CW == 4) ? (s ^ (p & 1)) : s);
        cplx o = u[s];
        if constexpr (s != 0) o = cmulc(o, TW[e]);
        EX[e * CW + col] = o;
      });
    });
    __syncwarp();
    // band side, forward: radix RA over p of column s = l + L i, only the band outputs; every lane stores
    // exactly the elements it loaded (getz)
    cplx* d = tile + (size_t)pair * pstride;
    static_for<0, NA>([&](auto ii) {
      constexpr int i = decltype(ii)::value;
      const int s = l + L * i;
      cplx z[RA];
      dft_out_dit<RA, D, false, KR::lo, KR::hi>(
          [&](auto pp) {
            constexpr int p = decltype(pp)::value;
            return EX[(p * RB + ((CW == 4) ? (s ^ (p & 1)) : s)) * CW + col];
          },
          z);
      static_for<KR::lo, KR::hi>([&](auto tt) {
        constexpr int t = decltype(tt)::value;
        const int zr = s + RB * t - zlo;
        if (zr >= 0 && zr < nzb) d[(size_t)zr * B] = z[t];
      });
    });
    __syncwarp();  // the exchange region is single buffered
  }
}

}  // namespace cpb
