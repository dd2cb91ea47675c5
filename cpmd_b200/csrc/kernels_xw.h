// Warp-autonomous mirror-pair x passes (sm_100a): k_xw_inv / k_xw_fwd.
//
// Same contract and memory layout as k_x_inv_m / k_x_fwd_m (kernels.h): the gather of a packed pair's
// coefficients + x-inverse FFT into T1 (zeroing(psi) + set_psi_2_states_g + x mltfft, state_utils.mod.F90:
// 132-189, fftmain_utils.mod.F90:93-94; with KIN also kin_energy / dotp, kin_energy_utils.mod.F90:62-110,
// dotp_utils.mod.F90:26-53), and the x-forward FFT of T1 fused with the +-G unpack, the kinetic term, the
// -f/2 scale and the c2 update (fwfftn's last mltfft with scale 1/N, fftmain_utils.mod.F90:134-136;
// vpsi_utils.mod.F90:626-673, add_wfn :717).
//
// The block kernels are latency bound (ncu, profiles/r02c_full_x_192x128.txt: no pipe above 45 %, top stalls
// block barrier and long scoreboard): four warps meet at two block barriers per pair and one of them idles in
// the slot-major pass.  Here ONE WARP owns H = 32/L/2 rays and their (-y,-z) mirrors for the whole pair loop,
// like the warp z kernels (kernels_zw.h):
//   * lane = slot * L + l: L = 8 role lanes per ray, N = RA * RB with RB = L.  Lane l runs the band-side
//     radix-RA transform of the positions x = l + RB k and the radix-RB transforms of p = l + L j; in the
//     latter its outputs x = p + RA q cover, over the 8 lanes of a ray, one whole 128-byte row of T1.
//   * every +G coefficient is fetched once with a 16-byte cp.async into the owner's private slots; the lane
//     that owns the -G position reads it there after a __syncwarp().  The exchange between the radix passes is
//     a per-warp region ordered by __syncwarp() only.  No block barrier inside the pair loop.
#pragma once
#include "kernels.h"

namespace cpb {

// Factorisations the warp x kernels are built for (RB = L = 8 role lanes per ray).
template <int N>
struct XWPick {
  static constexpr int ra = 0;
};
#define CPB_XW(N_, RA_)               \
  template <>                         \
  struct XWPick<N_> {                 \
    static constexpr int ra = RA_;    \
  };
CPB_XW(64, 8)
CPB_XW(128, 16)
CPB_XW(192, 24)
#undef CPB_XW

template <int RA, bool HALF>
struct XWCfg {
  static constexpr int L = 8, RB = 8;
  static constexpr int N = RA * RB;
  static constexpr int SLW = 32 / L;      // ray slots per warp
  static constexpr int H = SLW / 2;       // rays per warp (the other slots hold their mirrors)
  static constexpr int NB = RA / L;       // radix-RB sub-transforms per lane
  static constexpr int D = (RA % 3 == 0) ? 3 : 2;
  using KR = KRange<RA, HALF>;
  static constexpr int KC = KR::cnt;
  static constexpr int C0 = RA / 2;       // first decimated index that can hold a +G position
  static constexpr int NPOS = KR::hi - C0;
  static constexpr int WARPS = 4;
  static constexpr int NT = 32 * WARPS;
  static constexpr int EX_ELEMS = N * SLW;             // per warp
  static constexpr int ST_INV = 2 * NPOS * 32;         // per warp: [state][NPOS][lane]
  static constexpr size_t SMEM_INV = (size_t)(WARPS * (EX_ELEMS + ST_INV) + N) * sizeof(cplx);
  // forward: the exchange region only (it also carries the +-G hand-over after the second radix pass)
  static constexpr size_t SMEM_FWD = (size_t)(WARPS * EX_ELEMS + N) * sizeof(cplx);
  static_assert(NPOS * 32 <= EX_ELEMS, "the +-G hand-over fits the exchange region");
  static constexpr int minb(size_t smem) {
    return (smem * 4 + 4 * 1024 <= 228 * 1024) ? 4 : ((smem * 3 + 3 * 1024 <= 228 * 1024) ? 3 : 2);
  }
  static constexpr int MINB_INV = minb(SMEM_INV);
  static constexpr int MINB_FWD = minb(SMEM_FWD) > 3 ? 3 : minb(SMEM_FWD);
  // exchange position of (p, slot, a) in elements: rows of RB, the XOR spreads the 8 rows a quarter warp
  // reads in the radix-RB pass (fixed a, p = l + L j) over the 8 16-byte bank groups
  CPB_HD static constexpr int ex(int p, int s, int a) { return (p * SLW + s) * RB + (a ^ (p & 7)); }
};

// butterfly sum of 4 doubles over the lanes of a warp; every lane ends with the totals
#if defined(CPB_EMULATE)
void emu_warp_sum4(double (&q)[4]);  // simulator: through a per-warp scratch, same pairing order (emu_cuda.cpp)
CPB_D void warp_sum4(double (&q)[4]) { emu_warp_sum4(q); }
#else
CPB_D void warp_sum4(double (&q)[4]) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] += __shfl_xor_sync(0xffffffffu, q[i], m);
  }
}
#endif

// units (warps) of a launch: H rays and their mirrors each
template <int RA, bool HALF>
CPB_HD int xw_units(int nrays) {
  return ((nrays + 1) / 2 + XWCfg<RA, HALF>::H - 1) / XWCfg<RA, HALF>::H;
}

template <int RA, int B, bool HALF, bool KIN>
CPB_GLOBAL CPB_LAUNCH_BOUNDS((XWCfg<RA, HALF>::NT), (XWCfg<RA, HALF>::MINB_INV))
    k_xw_inv(const cplx* CPB_RESTRICT c0, long ldc, cplx* CPB_RESTRICT T1, PlanDev pd, PairDev pr, int npair, int ppg,
             double* CPB_RESTRICT kin_part, int geq0) {
  using C = XWCfg<RA, HALF>;
  using KR = typename C::KR;
  constexpr int N = C::N, L = C::L, RB = C::RB, SLW = C::SLW, H = C::H, NB = C::NB, C0 = C::C0, NPOS = C::NPOS, D = C::D;
  static_assert(B == L, "a T1 row of B consecutive x is written by the L lanes of a ray");
  CPB_DYN_SMEM(cplx, S);
  const int tid = threadIdx.x;
  const int w = tid >> 5, lane = tid & 31;
  const int s = lane / L, l = lane % L;
  cplx* EX = S + (size_t)w * C::EX_ELEMS;
  cplx* ST = S + (size_t)C::WARPS * C::EX_ELEMS + (size_t)w * C::ST_INV;
  cplx* TW = S + (size_t)C::WARPS * (C::EX_ELEMS + C::ST_INV);
  const int p0 = blockIdx.y * ppg;
  const int p1 = (p0 + ppg < npair) ? p0 + ppg : npair;
  pdl_trigger();
  for (int i = tid; i < N; i += C::NT) TW[i] = pd.tw1[((i / RB) * (i % RB)) % N];  // TW[p*RB + a] = w^(a p)
  pdl_wait();  // c0 and the pair descriptors may come from preceding work; T1 may still be read by it
  __syncthreads();  // the only block barrier: twiddles visible
  const int nunits = xw_units<RA, HALF>(pd.nrays);
  const int unit = blockIdx.x * C::WARPS + w;
  if (unit >= nunits) return;  // warp-uniform
  const XSlot sa = x_slot<SLW>(unit, s, pd.nrays);
  const int sp = sa.self ? s : (s + H) % SLW;  // slot that stages my -G partners
#if defined(CPB_DEBUG_KNOBS) && (CPB_DEBUG_KNOBS & 1)
  const size_t t1_pair = 0;  // experiment: every pair writes the same T1 region (stores stay in L2)
#else
  const size_t t1_pair = (size_t)pd.nxt * pd.nrays * B;
#endif
  // my band positions x = l + RB k -> plane-wave index (bit 31: -G partner) or kNoPW
  uint32_t tab[KR::cnt];
  double hgv[NPOS];
  static_for<0, KR::cnt>([&](auto kk) {
    constexpr int j = decltype(kk)::value;
    constexpr int k = KR::lo + j;
    const int xb = l + RB * k - pd.xlo;
    tab[j] = (sa.valid && xb >= 0 && xb < pd.nxb) ? __ldg(&pd.gtab[(size_t)xb * pd.nrp + sa.ray]) : kNoPW;
    if constexpr (k < C0) {
      if (!(tab[j] & kNegPW)) tab[j] = kNoPW;  // below the centre plane only -G partners exist
    }
    if constexpr (KIN && k >= C0) {
      hgv[k - C0] = (tab[j] != kNoPW && !(tab[j] & kNegPW)) ? __ldg(&pd.hg[tab[j]]) : 0.0;
    }
  });
  auto gather = [&](int s1, int s2) {
    const cplx* c1p = c0 + (size_t)s1 * ldc;
    const cplx* c2p = c0 + (size_t)(s2 < 0 ? s1 : s2) * ldc;
    static_for<C0, KR::hi>([&](auto kk) {
      constexpr int k = decltype(kk)::value;
      constexpr int j = k - KR::lo;
      if (tab[j] != kNoPW && !(tab[j] & kNegPW)) {
#if defined(CPB_DEBUG_KNOBS) && (CPB_DEBUG_KNOBS & 2)
        const uint32_t ig = tab[j] & 0x3ffu;  // experiment: the gather reads a 16 KB window of the column
#else
        const uint32_t ig = tab[j];
#endif
        cp_async16(&ST[(0 * NPOS + (k - C0)) * 32 + lane], c1p + ig);
        if (s2 >= 0) cp_async16(&ST[(1 * NPOS + (k - C0)) * 32 + lane], c2p + ig);
      }
    });
    cp_async_commit();
  };
  int s2cur = -1;
  if (p0 < p1) {
    s2cur = __ldg(&pr.st2[p0]);
    gather(__ldg(&pr.st1[p0]), s2cur);
  }
  for (int pair = p0; pair < p1; ++pair) {
    // descriptors of the next pair (used after the stage has been read)
    const int s1n = (pair + 1 < p1) ? __ldg(&pr.st1[pair + 1]) : 0;
    const int s2n = (pair + 1 < p1) ? __ldg(&pr.st2[pair + 1]) : -1;
    cp_async_wait_all();
    __syncwarp();  // the stage of this pair is complete and visible to the partner lanes; EX is free again
    const bool two = s2cur >= 0;
    double sk1 = 0.0, sd1 = 0.0, sk2 = 0.0, sd2 = 0.0;
    cplx v[RA];
    static_for<KR::lo, KR::hi>([&](auto kk) {
      constexpr int k = decltype(kk)::value;
      constexpr int j = k - KR::lo;
      const uint32_t e = tab[j];
      v[k] = mk(0.0, 0.0);
      if (e != kNoPW) {
        if (k >= C0 && !(e & kNegPW)) {
          // +G: c1 + i c2
          const cplx a = ST[(0 * NPOS + (k - C0)) * 32 + lane];
          const cplx bq = two ? ST[(1 * NPOS + (k - C0)) * 32 + lane] : mk(0.0, 0.0);
          v[k] = mk(a.x - bq.y, a.y + bq.x);
          if constexpr (KIN && k >= C0) {
            const double m1 = a.x * a.x + a.y * a.y, m2 = bq.x * bq.x + bq.y * bq.y;
            const bool g0 = (e == 0u) && geq0;  // dotp counts the real part of G = 0 once
            sk1 += hgv[k - C0] * m1;
            sk2 += hgv[k - C0] * m2;
            sd1 += g0 ? a.x * a.x : 2.0 * m1;
            sd2 += g0 ? bq.x * bq.x : 2.0 * m2;
          }
        } else {
          // -G: conj(c1) + i conj(c2) of the coefficient staged by the owner of (n1 - x, mirror ray)
          const int xm = pd.n1 - (l + RB * k);
          const int idx = (xm / RB - C0) * 32 + sp * L + (xm % RB);
          const cplx a = ST[0 * NPOS * 32 + idx];
          const cplx bq = two ? ST[1 * NPOS * 32 + idx] : mk(0.0, 0.0);
          v[k] = mk(a.x + bq.y, bq.x - a.y);
        }
      }
    });
    __syncwarp();  // every lane has read the stage (its own slots and its partners'): the next pair's copies may land
    if (pair + 1 < p1) gather(s1n, s2n);
    s2cur = s2n;
    dft_in_dif<RA, D, true, KR::lo, KR::hi>(v, [&](auto pp, cplx o) {
      constexpr int p = decltype(pp)::value;
      if constexpr (p != 0) o = cmul(o, TW[p * RB + l]);
      EX[(p * SLW + s) * RB + (l ^ (p & 7))] = o;
    });
    if constexpr (KIN) {
      // butterfly sum over the 32 lanes (fixed order, bit-stable), one value per unit and pair
      double q4[4] = {sk1, sd1, sk2, sd2};
      warp_sum4(q4);
      const double o = lane == 0 ? q4[0] : (lane == 1 ? q4[1] : (lane == 2 ? q4[2] : q4[3]));
      if (lane < 4) kin_part[((size_t)pair * nunits + unit) * 4 + lane] = o;
    }
    __syncwarp();
    if (sa.valid) {
      cplx* dst = T1 + (size_t)pair * t1_pair + (size_t)sa.ray * B;
      static_for<0, NB>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        const int p = l + L * j;
        cplx u[RB];
        static_for<0, RB>([&](auto aa) {
          constexpr int a = decltype(aa)::value;
          u[a] = EX[(p * SLW + s) * RB + (a ^ (p & 7))];
        });
        dft<RB, true>(u);
        static_for<0, RB>([&](auto qq) {
          constexpr int q = decltype(qq)::value;
          const int x = p + RA * q;  // x % B == l: the 8 lanes of a ray write one 128-byte row
          st_stream(&dst[(size_t)(x / B) * pd.nrays * B + (x % B)], u[q]);
        });
      });
    }
  }
}

// forward x pass fused with the unpack of vpsi.  pr.ca / pr.cb = fi / fip1 (vpsi_utils.mod.F90:627-633).
// ACC: c2 += result (reference semantics), else c2 = result.  The lanes of a ray read T1 in whole 128-byte rows
// (x = p + RA q, p = l + L j), transform over q in registers, exchange, transform over p with only the band outputs
// x = l + RB t; the lane that holds FFT[V psi](-G) hands it to the owner of the +G position through the (now free)
// exchange region, and the owner combines it with c0 / c2 fetched straight into registers.
template <int RA, int B, bool HALF, bool ACC>
CPB_GLOBAL CPB_LAUNCH_BOUNDS((XWCfg<RA, HALF>::NT), (XWCfg<RA, HALF>::MINB_FWD))
    k_xw_fwd(const cplx* CPB_RESTRICT T1, const cplx* CPB_RESTRICT c0, cplx* c2, long ldc, PlanDev pd, PairDev pr,
             int npair, int ppg) {
  using C = XWCfg<RA, HALF>;
  using KR = typename C::KR;
  constexpr int N = C::N, L = C::L, RB = C::RB, SLW = C::SLW, H = C::H, NB = C::NB, C0 = C::C0, NPOS = C::NPOS, D = C::D;
  static_assert(B == L, "a T1 row of B consecutive x is read by the L lanes of a ray");
  CPB_DYN_SMEM(cplx, S);
  const int tid = threadIdx.x;
  const int w = tid >> 5, lane = tid & 31;
  const int s = lane / L, l = lane % L;
  cplx* EX = S + (size_t)w * C::EX_ELEMS;
  cplx* TW = S + (size_t)C::WARPS * C::EX_ELEMS;
  const int p0 = blockIdx.y * ppg;
  const int p1 = (p0 + ppg < npair) ? p0 + ppg : npair;
  pdl_trigger();
  for (int i = tid; i < N; i += C::NT) TW[i] = pd.tw1[((i / RA) * (i % RA)) % N];  // TW[a*RA + p] = w^(a p): lanes run over p
  pdl_wait();  // T1 comes from the preceding kernel
  __syncthreads();  // the only block barrier: twiddles visible
  const int nunits = xw_units<RA, HALF>(pd.nrays);
  const int unit = blockIdx.x * C::WARPS + w;
  if (unit >= nunits) return;  // warp-uniform
  const XSlot sa = x_slot<SLW>(unit, s, pd.nrays);
  const int sp = sa.self ? s : (s + H) % SLW;
  const size_t t1_pair = (size_t)pd.nxt * pd.nrays * B;
  const double sc = pd.inv_n;
  uint32_t tab[KR::cnt];
  double g2v[NPOS];
  static_for<0, KR::cnt>([&](auto kk) {
    constexpr int j = decltype(kk)::value;
    constexpr int k = KR::lo + j;
    const int xb = l + RB * k - pd.xlo;
    tab[j] = (sa.valid && xb >= 0 && xb < pd.nxb) ? __ldg(&pd.gtab[(size_t)xb * pd.nrp + sa.ray]) : kNoPW;
    if constexpr (k < C0) {
      if (!(tab[j] & kNegPW)) tab[j] = kNoPW;
    }
    if constexpr (k >= C0) {
      g2v[k - C0] = (tab[j] != kNoPW && !(tab[j] & kNegPW)) ? pd.tpiba2 * __ldg(&pd.hg[tab[j]]) : 0.0;
    }
  });
  for (int pair = p0; pair < p1; ++pair) {
    const int s1 = __ldg(&pr.st1[pair]), s2 = __ldg(&pr.st2[pair]);
    const double fi = __ldg(&pr.ca[pair]), fip1 = __ldg(&pr.cb[pair]);
    // radix-RB pass over q of the rows p = l + L j: all loads of the pair are issued first
    const cplx* src = T1 + (size_t)pair * t1_pair + (size_t)(sa.valid ? sa.ray : 0) * B;
    cplx nv[NB * RB];
    static_for<0, NB>([&](auto jj) {
      constexpr int j = decltype(jj)::value;
      static_for<0, RB>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        const int x = (l + L * j) + RA * q;
        nv[j * RB + q] = sa.valid ? ld_stream(&src[(size_t)(x / B) * pd.nrays * B + (x % B)]) : mk(0.0, 0.0);
      });
    });
    static_for<0, NB>([&](auto jj) {
      constexpr int j = decltype(jj)::value;
      const int p = l + L * j;
      cplx u[RB];
      static_for<0, RB>([&](auto qq) { u[decltype(qq)::value] = nv[j * RB + decltype(qq)::value]; });
      dft<RB, false>(u);
      static_for<0, RB>([&](auto ss) {
        constexpr int sx = decltype(ss)::value;
        cplx o = u[sx];
        if constexpr (sx != 0) o = cmulc(o, TW[sx * RA + p]);
        EX[(p * SLW + s) * RB + (sx ^ (p & 7))] = o;
      });
    });
    __syncwarp();
    // radix-RA pass over p of column l, band outputs only, with fwfftn's scale
    cplx z[RA];
    dft_out_dit<RA, D, false, KR::lo, KR::hi>(
        [&](auto pp) {
          constexpr int p = decltype(pp)::value;
          return EX[(p * SLW + s) * RB + (l ^ (p & 7))];
        },
        z);
    __syncwarp();  // every lane has read the exchange region: it now carries the +-G hand-over [NPOS][lane]
    static_for<KR::lo, KR::hi>([&](auto tt) {
      constexpr int t = decltype(tt)::value;
      constexpr int j = t - KR::lo;
      z[t] = cscale(z[t], sc);
      if (tab[j] != kNoPW && (tab[j] & kNegPW)) {
        // hand FFT[V psi](-G) to the owner of the +G position (n1 - x, mirror ray)
        const int xm = pd.n1 - (l + RB * t);
        EX[(xm / RB - C0) * 32 + sp * L + (xm % RB)] = z[t];
      }
    });
    __syncwarp();
    // the coefficients the unpack needs, all loads first (c2 may not be reordered across its own stores otherwise)
    cplx ca1[NPOS], ca2[NPOS], co1[NPOS], co2[NPOS];
    static_for<C0, KR::hi>([&](auto tt) {
      constexpr int t = decltype(tt)::value;
      constexpr int j = t - KR::lo;
      const uint32_t ig = tab[j];
      const bool own = ig != kNoPW && !(ig & kNegPW);
      const size_t o1 = (size_t)s1 * ldc + (own ? ig : 0u), o2 = (size_t)(s2 < 0 ? s1 : s2) * ldc + (own ? ig : 0u);
      ca1[t - C0] = own ? __ldg(&c0[o1]) : mk(0.0, 0.0);
      ca2[t - C0] = (own && s2 >= 0) ? __ldg(&c0[o2]) : mk(0.0, 0.0);
      if constexpr (ACC) {
        co1[t - C0] = own ? ld_stream(&c2[o1]) : mk(0.0, 0.0);
        co2[t - C0] = (own && s2 >= 0) ? ld_stream(&c2[o2]) : mk(0.0, 0.0);
      }
    });
    static_for<C0, KR::hi>([&](auto tt) {
      constexpr int t = decltype(tt)::value;
      constexpr int j = t - KR::lo;
      const uint32_t ig = tab[j];
      if (ig != kNoPW && !(ig & kNegPW)) {
        const cplx psin = z[t];
        // G = 0 is its own partner (vpsi_utils.mod.F90:655-671 reads psi(nzhs) and psi(indzs), the same element)
        const bool selfpos = sa.self && 2 * (l + RB * t) == pd.n1;
        const cplx psii = selfpos ? psin : EX[(t - C0) * 32 + lane];
        const cplx a = ca1[t - C0];
        const cplx fp = cadd(psin, psii);
        const cplx fm = csub(psin, psii);
        const double g2 = g2v[t - C0];
        cplx r1 = mk(-fi * (g2 * a.x + fp.x), -fi * (g2 * a.y + fm.y));
        if constexpr (ACC) r1 = cadd(r1, co1[t - C0]);
        c2[(size_t)s1 * ldc + ig] = r1;
        if (s2 >= 0) {
          const cplx bq = ca2[t - C0];
          cplx r2 = mk(-fip1 * (g2 * bq.x + fp.y), -fip1 * (g2 * bq.y - fm.x));
          if constexpr (ACC) r2 = cadd(r2, co2[t - C0]);
          c2[(size_t)s2 * ldc + ig] = r2;
        }
      }
    });
    __syncwarp();  // the hand-over has been read: the exchange region is free for the next pair
  }
}

}  // namespace cpb
