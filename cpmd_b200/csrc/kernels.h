// sm_100a kernels of the vpsi / rhoofr pipeline (templates; instantiated per mesh length in
// axis_tu.cu).  See DESIGN.md for the data layout and the byte model.
//
// Pipeline per batch of packed state pairs (two real states per complex transform, Gamma point):
//   rhoofr:  k_x_inv  ->  k_y_inv  ->  k_z_rho
//   vpsi:    k_x_inv  ->  k_y_inv  ->  k_z_vpsi (z-inverse * V(r) * z-forward, fused)
//                     ->  k_y_fwd  ->  k_x_fwd (x-forward + unpack + kinetic + scale + c2 update)
//
// Replaces (not ports) the reference's per-pair sequence set_psi_2_states_g -> invfftn ->
// {build_density_sum | V*psi -> fwfftn -> unpack}  (rhoofr_utils.mod.F90:306-410,
// vpsi_utils.mod.F90:376-675, fftmain_utils.mod.F90:92-136) and its cuFFT staging
// (fftcu_methods.mod.F90).  Intermediates:
//   T1[pair][xt][ray][B]       after the x pass, only rays inside the cutoff disc (S_x bytes/pair);
//                              xt = x tile of B consecutive x (B*16 = 128-byte rows)
//   T2[pair][xtc][y][zr][B]    after the y pass, only z planes inside the band (S_y bytes/pair),
//                              held for ONE CHUNK of x tiles at a time (xtc = xt - xt0): the y and z
//                              passes of a chunk run back to back and the chunk buffer is reused, so
//                              it stays resident in the 126 MB L2 and never travels to HBM.
// The full n^3 complex box never exists in memory: the z pass consumes it in registers.
//
// Every 1-D FFT is a two-pass Cooley-Tukey N = RA*RB: each thread owns one radix-RA
// sub-transform in registers (codelets.h), one exchange through shared memory, then one radix-RB
// sub-transform.  Lanes of a warp run along the contiguous batch coordinate (x, or the ray slot
// in the x pass), so global accesses are 128-byte coalesced and shared-memory accesses are
// conflict free (16-byte elements, consecutive lanes -> consecutive elements).
#pragma once
#include "codelets.h"
#if defined(CPB_DBG_CLOCK)
#include <cstdio>
#endif

namespace cpb {

// optional phase timing of one block (debug builds only: -DCPB_DBG_CLOCK)
#if defined(CPB_DBG_CLOCK) && !defined(CPB_EMULATE)
#define CPB_CLK_INIT long long clk_[12]; long long clkacc_[14] = {0}; long long clk_prev_ = clock64(); int clk_n_ = 0; (void)clk_; (void)clk_n_
#define CPB_CLK(i) do { long long t_ = clock64(); clkacc_[i] += t_ - clk_prev_; clk_prev_ = t_; } while (0)
#define CPB_CLK_PRINT(name) do { if ((threadIdx.x == 0 || threadIdx.x == 160) && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0) printf(name " t%d clk: pro %lld %lld %lld %lld | top %lld scat %lld fetch %lld bar1 %lld A:dft %lld A:tw+sts %lld bar2 %lld B:lds+dft %lld B:stg %lld\n", (int)threadIdx.x, clkacc_[0], clkacc_[1], clkacc_[2], clkacc_[3], clkacc_[4], clkacc_[5], clkacc_[6], clkacc_[7], clkacc_[11], clkacc_[8], clkacc_[9], clkacc_[12], clkacc_[10]); } while (0)
#else
#define CPB_CLK_INIT
#define CPB_CLK(i)
#define CPB_CLK_PRINT(name)
#endif

struct PlanDev {
  int n1, n2, n3;     // mesh (spar%nr1s..)
  int kr1, kr2, kr3;  // padded real-space leading dimensions (fpar%kr1, kr2s, kr3s)
  int xlo, xhi;       // 0-based x band that holds G-sphere coefficients
  int zlo, nzb;       // 0-based first z plane of the band, number of planes (kr3min..kr3max)
  int nrays;          // internal ray count (>= msrays; dense in y inside every plane)
  int ntiles;         // x-pass tiles (mirror-closed groups of rays)
  int nxt;            // x tiles of B columns: ceil(n1 / B)
  const int* ylo;     // [nzb] first y with a ray in plane zr (0-based)
  const int* yhi;     // [nzb] last y (ylo > yhi: plane has no ray)
  const int* rayoff;  // [nzb] ray index of (ylo, zr)
  const int* slot_ray;       // [ntiles*SL] ray id of each slot or -1
  const int* ent_off;        // [ntiles+1] G entries of each tile
  const int* ent_ig;         // [nent] 0-based plane-wave index
  const uint32_t* ent_loc;   // [nent] lo16: x*LD+slot of +G ; hi16: same for -G
  const double* hg;          // [ngw]
  const cplx* tw1;           // [n1] exp(+2 pi i m / n1)
  const cplx* tw2;
  const cplx* tw3;
  double tpiba2;
  double inv_n;              // 1/(n1 n2 n3), fwfftn's scale (fftmain_utils.mod.F90:134)
};

// per-batch pair descriptors (device arrays, one entry per packed pair)
constexpr int kMaxGroup = 64;  // pairs one block may loop over (= largest batch a plan accepts)

struct PairDev {
  const int* st1;     // state index of the real part (column of c0), always valid
  const int* st2;     // state index of the imaginary part, or -1 (single-state path)
  const double* ca;   // rhoofr: f1/omega ; vpsi: fi   (vpsi_utils.mod.F90:627-633)
  const double* cb;   // rhoofr: f2/omega ; vpsi: fip1
};

// ---------------------------------------------------------------------------------------------
// two-pass tile FFT through shared memory; element idx of batch column b lives at S[idx*LD + b]
// ---------------------------------------------------------------------------------------------
template <int A, int B>
struct MaxOf {
  static constexpr int v = A > B ? A : B;
};

// HALF = true: the plan verified that the coefficient band lies inside [RB*KLO, RB*KHI) along the
// axis (always the case for the dual = 4 sphere), so only k in [KLO, KHI) of the decimated index
// is loaded (inverse) or stored (forward) and the first radix pass skips the zero terms (dft_in).
template <int R, bool HALF>
struct KRange {
  static constexpr int lo = HALF ? R / 4 : 0;
  static constexpr int hi = HALF ? (3 * R + 3) / 4 : R;  // exclusive
  static constexpr int cnt = hi - lo;
};

// resident blocks per SM the y/z kernels are compiled for (register budget 65536 / threads / this)
template <int R1, int R2>
struct YZBlocks {
  static constexpr int v = (MaxOf<R1, R2>::v <= 16 && R1 + R2 <= 28) ? 3 : 2;
};

// Thread role a (0 <= a < RB) holds v[k] = x[a + RB*k] (zero outside k in [LO,HI)).  Radix-RA
// transform, twiddle w^(a p), then st(p, value) for every p.
template <int RA, int RB, bool INV, int LO, int HI, class ST>
CPB_D void pass_a_st(cplx (&v)[RA], int a, const cplx* CPB_RESTRICT tw, ST&& st) {
  dft_in<RA, INV, LO, HI>(v);
  static_for<0, RA>([&](auto pp) {
    constexpr int p = decltype(pp)::value;
    cplx o = v[p];
    if constexpr (p != 0) {
#ifdef CPB_DBG_NOTW
      cplx t = mk(0.5 + a, 0.25 * p);
#else
      cplx t = __ldg(&tw[a * p]);
#endif
      if constexpr (!INV) t.y = -t.y;
      o = cmul(o, t);
    }
    st(p, o);
  });
}

template <int RA, int RB, bool INV, int LO, int HI>
CPB_D void pass_a_in(cplx (&v)[RA], int a, const cplx* CPB_RESTRICT tw, cplx* Sb, int LD) {
  pass_a_st<RA, RB, INV, LO, HI>(v, a, tw, [&](int p, cplx o) { Sb[(p * RB + a) * LD] = o; });
}

template <int RA, int RB, bool INV>
CPB_D void pass_a(cplx (&v)[RA], int a, const cplx* CPB_RESTRICT tw, cplx* Sb, int LD) {
  pass_a_in<RA, RB, INV, 0, RA>(v, a, tw, Sb, LD);
}

// Thread role p (0 <= p < RA).  On return u[q] = X[p + RA*q].
template <int RA, int RB, bool INV>
CPB_D void pass_b(cplx (&u)[RB], int p, const cplx* Sb, int LD) {
  static_for<0, RB>([&](auto aa) {
    constexpr int a = decltype(aa)::value;
    u[a] = Sb[(p * RB + a) * LD];
  });
  dft<RB, INV>(u);
}

// ---------------------------------------------------------------------------------------------
// x passes.  One block = one mirror-closed tile of up to SL rays (a ray and its (-y,-z) partner
// are in the same tile, so +G and -G of every plane wave are handled by the same block and c0 is
// read once) and loops over a group of packed pairs; the tile's plane-wave list (index into c0,
// tile-local positions of +G and -G) is loaded once and lives in registers.
//
// Shared memory:  SB[x][slot]   (LDB = SL+1)  scatter / gather buffer, band rows only are used;
//                 SX            exchange buffer between the two radix passes.
// Two thread roles:  "slot-major" (slot = tid % SL, row = tid / SL) touches SB conflict free;
//                    "x-major"    (row = tid % R1, slot = tid / R1) touches T1 in 128-byte rows
//                    (consecutive lanes = consecutive x of one ray).
// SX is laid out so that both roles access it conflict free (odd row pitch).
// grid = (ntiles, pair groups), block = SL * max(R1,R2)
// ---------------------------------------------------------------------------------------------
template <int R1, int R2, int SL, bool HALF>
struct XCfg {
  static constexpr int N = R1 * R2;
  static constexpr int RM = MaxOf<R1, R2>::v;
  static constexpr int NT = SL * RM;
  static constexpr int LDB = SL;  // slot-major lanes are contiguous: no padding needed
  static constexpr int P1 = R1 | 1;  // odd pitch of SX rows
  static constexpr int SB_ELEMS = N * LDB;
  static constexpr int SX_ELEMS = R2 * SL * P1;
  static constexpr size_t SMEM = (size_t)(SB_ELEMS + SX_ELEMS) * sizeof(cplx);
  // plane waves per thread: a tile holds at most SL * band / 2 (+G,-G) pairs (checked by the plan)
  static constexpr int EPT = (R2 * KRange<R1, HALF>::cnt + 2 * RM - 1) / (2 * RM) + (HALF ? 0 : 1);
  static constexpr int MINB = 512 / NT;  // 128 registers per thread
};

// x pass, inverse: scatter G coefficients of a packed pair into rays + FFT along x.
// Fuses zeroing(psi) + set_psi_2_states_g / set_psi_1_state_g (state_utils.mod.F90:132-189) +
// the x mltfft of fftnew (fftmain_utils.mod.F90:93-94).
template <int R1, int R2, int SL, int B, bool HALF>
CPB_GLOBAL CPB_LAUNCH_BOUNDS((XCfg<R1, R2, SL, HALF>::NT), (XCfg<R1, R2, SL, HALF>::MINB))
    k_x_inv(const cplx* CPB_RESTRICT c0, long ldc, cplx* CPB_RESTRICT T1, PlanDev pd, PairDev pr, int npair,
            int ppg) {
  using C = XCfg<R1, R2, SL, HALF>;
  using KR = KRange<R1, HALF>;
  constexpr int N = C::N, LDB = C::LDB, NT = C::NT, P1 = C::P1, EPT = C::EPT;
  CPB_CLK_INIT;
  CPB_DYN_SMEM(cplx, S);
  cplx* SB = S;
  cplx* SX = S + C::SB_ELEMS;
  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int p0 = blockIdx.y * ppg;
  const int p1 = (p0 + ppg < npair) ? p0 + ppg : npair;
  const int slotA = tid % SL, rA = tid / SL;  // slot-major role
  const int pB = tid % R1, slotB = tid / R1;  // x-major role (valid if slotB < SL)

  // tile's plane waves -> registers
  const int e0 = pd.ent_off[tile], e1 = pd.ent_off[tile + 1];
  int eig[EPT];
  uint32_t eloc[EPT];
  static_for<0, EPT>([&](auto jj) {
    constexpr int j = decltype(jj)::value;
    const int e = e0 + tid + j * NT;
    eig[j] = (e < e1) ? pd.ent_ig[e] : -1;
    eloc[j] = (e < e1) ? pd.ent_loc[e] : 0u;
  });
  // the group's pair descriptors -> shared memory (no dependent global load inside the pair loop)
  CPB_SHARED int sst1[kMaxGroup], sst2[kMaxGroup];
  for (int i = tid; i < p1 - p0; i += NT) {
    sst1[i] = pr.st1[p0 + i];
    sst2[i] = pr.st2[p0 + i];
  }
  __syncthreads();
  cplx ca[EPT], cb[EPT];
  auto fetch = [&](int pair) {
    const int s1 = sst1[pair - p0], s2 = sst2[pair - p0];
    const cplx* c1p = c0 + (size_t)s1 * ldc;
    const cplx* c2p = c0 + (size_t)(s2 < 0 ? 0 : s2) * ldc;
    static_for<0, EPT>([&](auto jj) {
      constexpr int j = decltype(jj)::value;
      if (eig[j] >= 0) {
#ifdef CPB_DBG_NOGATHER
        ca[j] = mk(1.0 + eig[j], (double)s1);
        cb[j] = mk(2.0, (double)s2 + (size_t)c1p % 3 + (size_t)c2p % 5);
#else
        ca[j] = c1p[eig[j]];
        cb[j] = (s2 >= 0) ? c2p[eig[j]] : mk(0.0, 0.0);
#endif
      }
    });
  };
  CPB_CLK(0);
  if (p0 < p1) fetch(p0);
  CPB_CLK(1);
  // zero the band rows once: the scatter rewrites the same positions for every pair and nothing
  // else writes SB
  for (int i = pd.xlo * LDB + tid; i < (pd.xhi + 1) * LDB; i += NT) SB[i] = mk(0.0, 0.0);
  const int rayB = (slotB < SL) ? pd.slot_ray[tile * SL + slotB] : -1;
  const size_t t1_pair = (size_t)pd.nxt * pd.nrays * B;
  CPB_CLK(2);
  __syncthreads();
  CPB_CLK(3);
  for (int pair = p0; pair < p1; ++pair) {
    CPB_CLK(4);
    static_for<0, EPT>([&](auto jj) {
      constexpr int j = decltype(jj)::value;
#ifdef CPB_DBG_NOSCATTER
      if (eig[j] >= 0 && ca[j].x == 1.2345e-300) {
#else
      if (eig[j] >= 0) {
#endif
        const int lp = eloc[j] & 0xffffu, lm = eloc[j] >> 16;
        const cplx a = ca[j], bq = cb[j];
        SB[lp] = mk(a.x - bq.y, a.y + bq.x);                // c1 + i c2
        if (lm != lp) SB[lm] = mk(a.x + bq.y, bq.x - a.y);  // conj(c1) + i conj(c2)
      }
    });
    CPB_CLK(5);
    if (pair + 1 < p1) fetch(pair + 1);
    CPB_CLK(6);
    __syncthreads();
    CPB_CLK(7);
    if (rA < R2) {
      cplx v[R1];
      static_for<0, R1>([&](auto kk) {
        constexpr int k = decltype(kk)::value;
        if constexpr (k >= KR::lo && k < KR::hi) {
          const int x = rA + R2 * k;
          v[k] = (x >= pd.xlo && x <= pd.xhi) ? SB[x * LDB + slotA] : mk(0.0, 0.0);
        } else {
          v[k] = mk(0.0, 0.0);
        }
      });
      cplx* dst = SX + (rA * SL + slotA) * P1;
#if defined(CPB_DBG_CLOCK)
      dft_in<R1, true, KR::lo, KR::hi>(v);
      if (v[1].x == 1.2345e-300) v[2].x = 0;  // consume
      CPB_CLK(11);
      static_for<0, R1>([&](auto pp) {
        constexpr int p = decltype(pp)::value;
        cplx o = v[p];
        if constexpr (p != 0) o = cmul(o, __ldg(&pd.tw1[rA * p]));
        dst[p] = o;
      });
#else
      pass_a_st<R1, R2, true, KR::lo, KR::hi>(v, rA, pd.tw1, [&](int p, cplx o) { dst[p] = o; });
#endif
    }
    CPB_CLK(8);
    __syncthreads();
    CPB_CLK(9);
    if (slotB < SL) {
      cplx u[R2];
      static_for<0, R2>([&](auto aa) {
        constexpr int a = decltype(aa)::value;
        u[a] = SX[(a * SL + slotB) * P1 + pB];
      });
      dft<R2, true>(u);
#if defined(CPB_DBG_CLOCK)
      if (u[1].x == 1.2345e-300) u[2].x = 0;  // consume
      CPB_CLK(12);
#endif
      if (rayB >= 0) {
        cplx* dst = T1 + (size_t)pair * t1_pair + (size_t)rayB * B;
        static_for<0, R2>([&](auto qq) {
          constexpr int q = decltype(qq)::value;
          const int x = pB + R1 * q;
#ifdef CPB_DBG_NOSTORE
          if (u[q].x == 1.2345e-300) st_stream(&dst[(size_t)(x / B) * pd.nrays * B + (x % B)], u[q]);
#else
          st_stream(&dst[(size_t)(x / B) * pd.nrays * B + (x % B)], u[q]);
#endif
        });
      }
    }
    CPB_CLK(10);
  }
  CPB_CLK_PRINT("x_inv");
}

// x pass, forward: FFT along x (scale 1/N_total) + unpack of the two states + kinetic term +
// occupation scale + accumulation into c2.  Fuses the last mltfft of fwfftn
// (fftmain_utils.mod.F90:134-136) with vpsi_utils.mod.F90:626-673 and add_wfn (:717).
// ACC: c2 += result (reference semantics) ; !ACC: c2 = result.
template <int R1, int R2, int SL, int B, bool HALF, bool ACC>
CPB_GLOBAL CPB_LAUNCH_BOUNDS((XCfg<R1, R2, SL, HALF>::NT), (XCfg<R1, R2, SL, HALF>::MINB))
    k_x_fwd(const cplx* CPB_RESTRICT T1, const cplx* CPB_RESTRICT c0, cplx* CPB_RESTRICT c2, long ldc,
            PlanDev pd, PairDev pr, int npair, int ppg) {
  using C = XCfg<R1, R2, SL, HALF>;
  using KR = KRange<R1, HALF>;
  constexpr int N = C::N, LDB = C::LDB, NT = C::NT, P1 = C::P1, EPT = C::EPT;
  CPB_DYN_SMEM(cplx, S);
  cplx* SB = S;
  cplx* SX = S + C::SB_ELEMS;
  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int p0 = blockIdx.y * ppg;
  const int p1 = (p0 + ppg < npair) ? p0 + ppg : npair;
  const int slotA = tid % SL, rA = tid / SL;
  const int pB = tid % R1, slotB = tid / R1;

  const int e0 = pd.ent_off[tile], e1 = pd.ent_off[tile + 1];
  int eig[EPT];
  uint32_t eloc[EPT];
  double eg2[EPT];
  static_for<0, EPT>([&](auto jj) {
    constexpr int j = decltype(jj)::value;
    const int e = e0 + tid + j * NT;
    eig[j] = (e < e1) ? pd.ent_ig[e] : -1;
    eloc[j] = (e < e1) ? pd.ent_loc[e] : 0u;
    eg2[j] = (e < e1) ? pd.tpiba2 * pd.hg[eig[j]] : 0.0;
  });
  const int rayB = (slotB < SL) ? pd.slot_ray[tile * SL + slotB] : -1;
  const size_t t1_pair = (size_t)pd.nxt * pd.nrays * B;
  // x-major role: element k of the first (radix-R2) pass is x = pB + R1*k
  cplx nv[R2];
  auto fetch = [&](int pair) {
    const cplx* src = T1 + (size_t)pair * t1_pair + (size_t)(rayB < 0 ? 0 : rayB) * B;
    static_for<0, R2>([&](auto kk) {
      constexpr int k = decltype(kk)::value;
      const int x = pB + R1 * k;
      nv[k] = (rayB >= 0) ? ld_stream(&src[(size_t)(x / B) * pd.nrays * B + (x % B)]) : mk(0.0, 0.0);
    });
  };
  if (slotB < SL && p0 < p1) fetch(p0);
  const double sc = pd.inv_n;
  CPB_SHARED int sst1[kMaxGroup], sst2[kMaxGroup];
  CPB_SHARED double sca[kMaxGroup], scb[kMaxGroup];
  for (int i = tid; i < p1 - p0; i += NT) {
    sst1[i] = pr.st1[p0 + i];
    sst2[i] = pr.st2[p0 + i];
    sca[i] = pr.ca[p0 + i];
    scb[i] = pr.cb[p0 + i];
  }
  __syncthreads();
  for (int pair = p0; pair < p1; ++pair) {
    const int s1 = sst1[pair - p0], s2 = sst2[pair - p0];
    const double fi = sca[pair - p0], fip1 = scb[pair - p0];
    const cplx* c1p = c0 + (size_t)s1 * ldc;
    const cplx* c2p = c0 + (size_t)(s2 < 0 ? 0 : s2) * ldc;
    cplx* o1 = c2 + (size_t)s1 * ldc;
    cplx* o2 = c2 + (size_t)(s2 < 0 ? 0 : s2) * ldc;
    // this pair's c0 values: issued now, used after the transform
    cplx ca[EPT], cb[EPT];
    static_for<0, EPT>([&](auto jj) {
      constexpr int j = decltype(jj)::value;
      if (eig[j] >= 0) {
        ca[j] = c1p[eig[j]];
        cb[j] = (s2 >= 0) ? c2p[eig[j]] : mk(0.0, 0.0);
      }
    });
    if (slotB < SL) {
      cplx v[R2];
      static_for<0, R2>([&](auto kk) { v[decltype(kk)::value] = nv[decltype(kk)::value]; });
      // forward transform uses the mirrored factorisation (R2 first, then R1)
      pass_a_st<R2, R1, false, 0, R2>(v, pB, pd.tw1,
                                      [&](int p, cplx o) { SX[(p * SL + slotB) * P1 + pB] = o; });
      if (pair + 1 < p1) fetch(pair + 1);
    }
    __syncthreads();
    if (rA < R2) {
      cplx u[R1];
      static_for<0, R1>([&](auto aa) {
        constexpr int a = decltype(aa)::value;
        u[a] = SX[(rA * SL + slotA) * P1 + a];
      });
      dft<R1, false>(u);
      static_for<KR::lo, KR::hi>([&](auto tt) {
        constexpr int t = decltype(tt)::value;
        const int x = rA + R2 * t;
        if (x >= pd.xlo && x <= pd.xhi) SB[x * LDB + slotA] = u[t];
      });
    }
    __syncthreads();
    static_for<0, EPT>([&](auto jj) {
      constexpr int j = decltype(jj)::value;
      if (eig[j] >= 0) {
        const int ig = eig[j];
        const cplx psin = SB[eloc[j] & 0xffffu];
        const cplx psii = SB[eloc[j] >> 16];
        const cplx fp = mk((psin.x + psii.x) * sc, (psin.y + psii.y) * sc);
        const cplx fm = mk((psin.x - psii.x) * sc, (psin.y - psii.y) * sc);
        const double g2 = eg2[j];
        const cplx a = ca[j];
        cplx r1 = mk(-fi * (g2 * a.x + fp.x), -fi * (g2 * a.y + fm.y));
        if (ACC) r1 = cadd(r1, o1[ig]);
        o1[ig] = r1;
        if (s2 >= 0) {
          const cplx bq = cb[j];
          cplx r2 = mk(-fip1 * (g2 * bq.x + fp.y), -fip1 * (g2 * bq.y - fm.x));
          if (ACC) r2 = cadd(r2, o2[ig]);
          o2[ig] = r2;
        }
      }
    });
  }
}

// ---------------------------------------------------------------------------------------------
// y and z passes.  Common structure: a block owns B consecutive x (one 128-byte row per (y|z)
// index) and loops over several packed pairs.  The band elements of the NEXT pair are fetched
// into registers right after the first radix pass of the current pair, so their L2/HBM latency is
// hidden behind the second radix pass; the exchange buffer in shared memory is double buffered,
// which leaves one block barrier per transform.
//
// HALF = true: the plan verified that the coefficient band lies inside [RB*KLO, RB*KHI) along this
// axis (always the case for the dual = 4 sphere), so only k in [KLO, KHI) of the decimated index
// is loaded (inverse) or stored (forward) and the first radix pass skips the zero terms (dft_in).
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// y pass, inverse.  Block = (x tile of the chunk, z plane of the band, group of pairs).  Reads the
// rays of the plane (zero outside [ylo,yhi]: unpack_x2y's zero fill, fftutil_utils.mod.F90:413-457),
// writes all n2 rows of the chunk's T2.  grid = (x tiles of the chunk, nzb, pair groups)
// ---------------------------------------------------------------------------------------------
template <int R1, int R2, int B, bool HALF>
CPB_GLOBAL CPB_LAUNCH_BOUNDS(B* MaxOf<R1, R2>::v, YZBlocks<R1, R2>::v)
    k_y_inv(const cplx* CPB_RESTRICT T1, cplx* CPB_RESTRICT T2, PlanDev pd, int xt0, int npair, int ppg) {
  constexpr int N = R1 * R2;
  using KR = KRange<R1, HALF>;
  CPB_DYN_SMEM(cplx, S);
  const int tid = threadIdx.x;
  const int b = tid % B, r = tid / B;
  const int xtc = blockIdx.x, nxc = gridDim.x;
  const int zr = blockIdx.y;
  const int p0 = blockIdx.z * ppg;
  const int p1 = (p0 + ppg < npair) ? p0 + ppg : npair;
  const int ylo = pd.ylo[zr], yhi = pd.yhi[zr];
  const size_t t1_pair = (size_t)pd.nxt * pd.nrays * B;
  const size_t t2_pair = (size_t)nxc * N * pd.nzb * B;
  const cplx* src = T1 + ((size_t)(xt0 + xtc) * pd.nrays + pd.rayoff[zr]) * B + b;
  cplx* dst = T2 + ((size_t)xtc * N * pd.nzb + zr) * B + b;
  cplx nv[KR::cnt];
  auto fetch = [&](int pair) {
    const cplx* s = src + (size_t)pair * t1_pair;
    static_for<0, KR::cnt>([&](auto kk) {
      constexpr int k = KR::lo + decltype(kk)::value;
      const int y = r + R2 * k;
      nv[decltype(kk)::value] = (y >= ylo && y <= yhi) ? s[(y - ylo) * B] : mk(0.0, 0.0);
    });
  };
  if (r < R2 && p0 < p1) fetch(p0);
  int buf = 0;
  for (int pair = p0; pair < p1; ++pair) {
    cplx* Sb = S + buf * (N * B) + b;
    if (r < R2) {
      cplx v[R1];
      static_for<0, R1>([&](auto kk) {
        constexpr int k = decltype(kk)::value;
        if constexpr (k >= KR::lo && k < KR::hi) v[k] = nv[k - KR::lo];
        else v[k] = mk(0.0, 0.0);
      });
      pass_a_in<R1, R2, true, KR::lo, KR::hi>(v, r, pd.tw2, Sb, B);
      if (pair + 1 < p1) fetch(pair + 1);
    }
    __syncthreads();
    if (r < R1) {
      cplx u[R2];
      pass_b<R1, R2, true>(u, r, Sb, B);
      cplx* d = dst + (size_t)pair * t2_pair;
      static_for<0, R2>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        d[(size_t)(r + R1 * q) * pd.nzb * B] = u[q];
      });
    }
    buf ^= 1;
  }
}

// y pass, forward: reads all n2 rows of the chunk's T2, writes only the rays of the plane into T1.
template <int R1, int R2, int B, bool HALF>
CPB_GLOBAL CPB_LAUNCH_BOUNDS(B* MaxOf<R1, R2>::v, YZBlocks<R1, R2>::v)
    k_y_fwd(const cplx* CPB_RESTRICT T2, cplx* CPB_RESTRICT T1, PlanDev pd, int xt0, int npair, int ppg) {
  constexpr int N = R1 * R2;
  using KR = KRange<R1, HALF>;
  CPB_DYN_SMEM(cplx, S);
  const int tid = threadIdx.x;
  const int b = tid % B, r = tid / B;
  const int xtc = blockIdx.x, nxc = gridDim.x;
  const int zr = blockIdx.y;
  const int p0 = blockIdx.z * ppg;
  const int p1 = (p0 + ppg < npair) ? p0 + ppg : npair;
  const int ylo = pd.ylo[zr], yhi = pd.yhi[zr];
  const size_t t1_pair = (size_t)pd.nxt * pd.nrays * B;
  const size_t t2_pair = (size_t)nxc * N * pd.nzb * B;
  const size_t ystride = (size_t)pd.nzb * B;
  const cplx* src = T2 + ((size_t)xtc * N * pd.nzb + zr) * B + b;
  cplx* dst = T1 + ((size_t)(xt0 + xtc) * pd.nrays + pd.rayoff[zr]) * B + b;
  cplx nv[R2];
  auto fetch = [&](int pair) {
    const cplx* s = src + (size_t)pair * t2_pair;
    static_for<0, R2>([&](auto kk) {
      constexpr int k = decltype(kk)::value;
      nv[k] = s[(size_t)(r + R1 * k) * ystride];
    });
  };
  if (r < R1 && p0 < p1) fetch(p0);
  int buf = 0;
  for (int pair = p0; pair < p1; ++pair) {
    cplx* Sb = S + buf * (N * B) + b;
    if (r < R1) {
      cplx v[R2];
      static_for<0, R2>([&](auto kk) { v[decltype(kk)::value] = nv[decltype(kk)::value]; });
      pass_a<R2, R1, false>(v, r, pd.tw2, Sb, B);
      if (pair + 1 < p1) fetch(pair + 1);
    }
    __syncthreads();
    if (r < R2) {
      cplx u[R1];
      pass_b<R2, R1, false>(u, r, Sb, B);
      cplx* d = dst + (size_t)pair * t1_pair;
      static_for<KR::lo, KR::hi>([&](auto tt) {
        constexpr int t = decltype(tt)::value;
        const int y = r + R2 * t;
        if (y >= ylo && y <= yhi) d[(y - ylo) * B] = u[t];
      });
    }
    buf ^= 1;
  }
}

// ---------------------------------------------------------------------------------------------
// z pass of rhoofr: z-inverse FFT (band zero-padded to n3: putz, fftutil_utils.mod.F90:87-104)
// fused with build_density_sum (density_utils.mod.F90:61-83).  The block keeps its rho tile in
// registers over all pairs of the batch and does ONE read-modify-write of rho(r) per batch.
// grid = (x tiles of the chunk, n2)
// ---------------------------------------------------------------------------------------------
template <int R1, int R2, int B, bool HALF>
CPB_GLOBAL CPB_LAUNCH_BOUNDS(B* MaxOf<R1, R2>::v, YZBlocks<R1, R2>::v)
    k_z_rho(const cplx* CPB_RESTRICT T2, double* rho, PlanDev pd, PairDev pr, int npair,
            int xt0) {
  constexpr int N = R1 * R2;
  using KR = KRange<R1, HALF>;
  CPB_DYN_SMEM(cplx, S);
  const int tid = threadIdx.x;
  const int b = tid % B, r = tid / B;
  const int xtc = blockIdx.x, nxc = gridDim.x;
  const int x = (xt0 + xtc) * B + b;
  const int y = blockIdx.y;
  const bool xok = x < pd.n1;
  const size_t pstride = (size_t)nxc * pd.n2 * pd.nzb * B;
  const cplx* tile = T2 + ((size_t)xtc * pd.n2 + y) * pd.nzb * B + b;
  const int zlo = pd.zlo, nzb = pd.nzb;
  cplx nv[KR::cnt];
  auto fetch = [&](int pair) {
    const cplx* s = tile + (size_t)pair * pstride;
    static_for<0, KR::cnt>([&](auto kk) {
      constexpr int k = KR::lo + decltype(kk)::value;
      const int zr = r + R2 * k - zlo;
      nv[decltype(kk)::value] = (zr >= 0 && zr < nzb) ? s[zr * B] : mk(0.0, 0.0);
    });
  };
  // the accumulators start from rho itself: the read-modify-write's read overlaps the first tile
  double acc[R2];
  static_for<0, R2>([&](auto qq) {
    constexpr int q = decltype(qq)::value;
    acc[q] = (r < R1 && xok) ? rho[((size_t)(r + R1 * q) * pd.kr2 + y) * pd.kr1 + x] : 0.0;
  });
  if (r < R2 && npair > 0) fetch(0);
  CPB_SHARED double sca[kMaxGroup], scb[kMaxGroup];
  for (int i = tid; i < npair; i += B * MaxOf<R1, R2>::v) {
    sca[i] = pr.ca[i];
    scb[i] = pr.cb[i];
  }
  __syncthreads();
  int buf = 0;
  for (int pair = 0; pair < npair; ++pair) {
    cplx* Sb = S + buf * (N * B) + b;
    const double ca = sca[pair], cb = scb[pair];
    if (r < R2) {
      cplx v[R1];
      static_for<0, R1>([&](auto kk) {
        constexpr int k = decltype(kk)::value;
        if constexpr (k >= KR::lo && k < KR::hi) v[k] = nv[k - KR::lo];
        else v[k] = mk(0.0, 0.0);
      });
      pass_a_in<R1, R2, true, KR::lo, KR::hi>(v, r, pd.tw3, Sb, B);
      if (pair + 1 < npair) fetch(pair + 1);
    }
    __syncthreads();
    if (r < R1) {
      cplx u[R2];
      pass_b<R1, R2, true>(u, r, Sb, B);
      static_for<0, R2>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        acc[q] += ca * (u[q].x * u[q].x) + cb * (u[q].y * u[q].y);
      });
    }
    buf ^= 1;
  }
  if (r < R1 && xok) {
    static_for<0, R2>([&](auto qq) {
      constexpr int q = decltype(qq)::value;
      const size_t o = ((size_t)(r + R1 * q) * pd.kr2 + y) * pd.kr1 + x;
      rho[o] = acc[q];
    });
  }
}

// ---------------------------------------------------------------------------------------------
// z pass of vpsi: z-inverse FFT, multiply by V(r) (vpsi_utils.mod.F90:487-493), z-forward FFT,
// store only the band back in place (getz, fftutil_utils.mod.F90:106-125).  The real-space
// psi(r) never touches memory.  The forward transform uses the mirrored factorisation so every
// thread stores exactly the elements it loaded.  V tile lives in registers over the pair loop.
// grid = (x tiles of the chunk, n2, pair groups)
// ---------------------------------------------------------------------------------------------
template <int R1, int R2, int B, bool HALF>
CPB_GLOBAL CPB_LAUNCH_BOUNDS(B* MaxOf<R1, R2>::v, YZBlocks<R1, R2>::v)
    k_z_vpsi(cplx* T2, const double* CPB_RESTRICT vpot, PlanDev pd, int xt0, int npair, int ppg) {
  constexpr int N = R1 * R2;
  using KR = KRange<R1, HALF>;
  CPB_DYN_SMEM(cplx, S);
  const int tid = threadIdx.x;
  const int b = tid % B, r = tid / B;
  const int xtc = blockIdx.x, nxc = gridDim.x;
  const int x = (xt0 + xtc) * B + b;
  const int y = blockIdx.y;
  const int p0 = blockIdx.z * ppg;
  const int p1 = (p0 + ppg < npair) ? p0 + ppg : npair;
  const bool xok = x < pd.n1;
  const size_t pstride = (size_t)nxc * pd.n2 * pd.nzb * B;
  cplx* tile = T2 + ((size_t)xtc * pd.n2 + y) * pd.nzb * B + b;
  const int zlo = pd.zlo, nzb = pd.nzb;
  double vv[R2];
  static_for<0, R2>([&](auto qq) {
    constexpr int q = decltype(qq)::value;
    vv[q] = (r < R1 && xok) ? __ldg(&vpot[((size_t)(r + R1 * q) * pd.kr2 + y) * pd.kr1 + x]) : 0.0;
  });
  cplx nv[KR::cnt];
  auto fetch = [&](int pair) {
    const cplx* s = tile + (size_t)pair * pstride;
    static_for<0, KR::cnt>([&](auto kk) {
      constexpr int k = KR::lo + decltype(kk)::value;
      const int zr = r + R2 * k - zlo;
      nv[decltype(kk)::value] = (zr >= 0 && zr < nzb) ? s[zr * B] : mk(0.0, 0.0);
    });
  };
  if (r < R2 && p0 < p1) fetch(p0);
  cplx* Sa = S + b;            // exchange buffer of the inverse transform
  cplx* Sf = S + N * B + b;    // exchange buffer of the forward transform
  for (int pair = p0; pair < p1; ++pair) {
    if (r < R2) {
      cplx v[R1];
      static_for<0, R1>([&](auto kk) {
        constexpr int k = decltype(kk)::value;
        if constexpr (k >= KR::lo && k < KR::hi) v[k] = nv[k - KR::lo];
        else v[k] = mk(0.0, 0.0);
      });
      pass_a_in<R1, R2, true, KR::lo, KR::hi>(v, r, pd.tw3, Sa, B);
      if (pair + 1 < p1) fetch(pair + 1);
    }
    __syncthreads();
    if (r < R1) {
      cplx u[R2];
      pass_b<R1, R2, true>(u, r, Sa, B);
      static_for<0, R2>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        u[q].x *= vv[q];
        u[q].y *= vv[q];
      });
      pass_a<R2, R1, false>(u, r, pd.tw3, Sf, B);
    }
    __syncthreads();
    if (r < R2) {
      cplx w[R1];
      pass_b<R2, R1, false>(w, r, Sf, B);
      cplx* d = tile + (size_t)pair * pstride;
      static_for<KR::lo, KR::hi>([&](auto kk) {
        constexpr int k = decltype(kk)::value;
        const int zr = r + R2 * k - zlo;
        if (zr >= 0 && zr < nzb) d[zr * B] = w[k];
      });
    }
  }
}

}  // namespace cpb
