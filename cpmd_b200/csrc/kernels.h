// sm_100a kernels of the vpsi / rhoofr pipeline (templates; instantiated per mesh length in
// axis_tu.cu).  See DESIGN.md for the data layout and the byte model.
//
// Pipeline per batch of packed state pairs (two real states per complex transform, Gamma point):
//   rhoofr:  k_x_inv (gather + pack + x)  ->  k_y_inv  ->  k_z_rho
//   vpsi:    k_x_inv  ->  k_y_inv  ->  k_z_vpsi (z-inverse * V(r) * z-forward, fused)
//                     ->  k_y_fwd  ->  k_x_fwd -> k_unpack (unpack + kinetic + scale + c2 update)
//
// Replaces (not ports) the reference's per-pair sequence set_psi_2_states_g -> invfftn ->
// {build_density_sum | V*psi -> fwfftn -> unpack}  (rhoofr_utils.mod.F90:306-410,
// vpsi_utils.mod.F90:376-675, fftmain_utils.mod.F90:92-136) and its cuFFT staging
// (fftcu_methods.mod.F90).  Intermediates:
//   T1[pair][xt][ray][B]       after the x pass, only rays inside the cutoff disc (S_x bytes/pair);
//                              xt = x tile of B consecutive x (B*16 = 128-byte rows)
//   T2[pair][xtc][y][zr][B]    after the y pass, only z planes inside the band (S_y bytes/pair).  By
//                              default it holds ALL x tiles of the batch and goes through HBM once in
//                              each direction (measured: ncu DRAM bytes = the algorithmic S_y, see
//                              profiles/*_traffic.json); CPB_CHUNK_XT splits the y/z passes into chunks
//                              of x tiles (xtc = xt - xt0), a tuning hook that did not pay off.
// The full n^3 complex box never exists in memory: the z pass consumes it in registers.
//
// Every 1-D FFT is a two-pass Cooley-Tukey N = RA*RB: each thread owns one radix-RA
// sub-transform in registers (codelets.h), one exchange through shared memory, then one radix-RB
// sub-transform.  Lanes of a warp run along the contiguous batch coordinate (x, or the ray slot
// in the x pass), so global accesses are 128-byte coalesced and shared-memory accesses are
// conflict free (16-byte elements, consecutive lanes -> consecutive elements).
#pragma once
#include "codelets.h"
namespace cpb {

struct PlanDev {
  int n1, n2, n3;     // mesh (spar%nr1s..)
  int kr1, kr2, kr3;  // padded real-space leading dimensions (fpar%kr1, kr2s, kr3s)
  int xlo, nxb;       // 0-based first x of the band that holds G-sphere coefficients, band width
  int zlo, nzb;       // 0-based first z plane of the band, number of planes (kr3min..kr3max)
  int nrays;          // internal ray count (>= msrays; dense in y inside every plane)
  int nyb;            // largest number of rays in one z plane (y band width)
  int nrp;            // row pitch of the band-ray storage (nrays rounded up to the x-pass tile)
  int nxt;            // x tiles of B columns: ceil(n1 / B)
  int ngw;
  const int* ylo;     // [nzb] first y with a ray in plane zr (0-based)
  const int* yhi;     // [nzb] last y (ylo > yhi: plane has no ray)
  const int* rayoff;  // [nzb] ray index of (ylo, zr)
  const uint32_t* gtab;      // [nxb*nrp] plane wave stored at a band-ray position (kNoPW / kNegPW flags)
  const uint32_t* gpos;      // [ngw] band-ray storage position of +G: xb*nrp + ray  (nzhs)
  const uint32_t* gneg;      // [ngw] same for -G                                   (indzs)
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
#ifndef CPB_YZ_BLOCKS_SCALE
#define CPB_YZ_BLOCKS_SCALE 1  // tuning builds with narrower blocks (CPB_B=4) ask for more of them
#endif
template <int R1, int R2>
struct YZBlocks {
  static constexpr int v = ((MaxOf<R1, R2>::v <= 16 && R1 + R2 <= 28) ? 3 : 2) * CPB_YZ_BLOCKS_SCALE;
};

// Thread role a (0 <= a < RB) holds v[k] = x[a + RB*k] (zero outside k in [LO,HI)).  Radix-RA
// transform, twiddle w^(a p), then st(p, value) for every p.  TWS: `tw` is a shared-memory copy
// of the twiddle table (plain loads) instead of the global one (read-only cache loads).
template <int RA, int RB, bool INV, int LO, int HI, bool TWS = false, class ST>
CPB_D void pass_a_st(cplx (&v)[RA], int a, const cplx* CPB_RESTRICT tw, ST&& st) {
  dft_in<RA, INV, LO, HI>(v);
  if constexpr (!TWS) {
    // global twiddle table through the read-only path: the compiler is free to hoist the loads
    static_for<0, RA>([&](auto pp) {
      constexpr int p = decltype(pp)::value;
      cplx o = v[p];
      if constexpr (p != 0) {
        cplx t = __ldg(&tw[a * p]);
        if constexpr (!INV) t.y = -t.y;
        o = cmul(o, t);
      }
      st(p, o);
    });
  } else {
    // Shared-memory twiddle table: the loads are written in chunks of kTwChunk *before* the stores
    // of the previous chunk, because the compiler may not hoist a shared-memory load above a
    // shared-memory store (possible alias) and one load -> multiply -> store chain per output
    // would expose the load latency RA-1 times.
    constexpr int kTwChunk = 8;
    auto load = [&](auto cc, cplx (&t)[kTwChunk]) {
      constexpr int c0 = decltype(cc)::value;
      static_for<0, kTwChunk>([&](auto jj) {
        constexpr int p = c0 + decltype(jj)::value;
        if constexpr (p < RA) {
          t[p - c0] = tw[a * p];
          if constexpr (!INV) t[p - c0].y = -t[p - c0].y;
        }
      });
    };
    cplx t[kTwChunk];
    load(IC<1>{}, t);
    st(0, v[0]);
    static_for<0, (RA - 1 + kTwChunk - 1) / kTwChunk>([&](auto cc) {
      constexpr int c0 = 1 + decltype(cc)::value * kTwChunk;
      static_for<0, kTwChunk>([&](auto jj) {
        constexpr int p = c0 + decltype(jj)::value;
        if constexpr (p < RA) v[p] = cmul(v[p], t[p - c0]);
      });
      if constexpr (c0 + kTwChunk < RA) load(IC<c0 + kTwChunk>{}, t);
      static_for<0, kTwChunk>([&](auto jj) {
        constexpr int p = c0 + decltype(jj)::value;
        if constexpr (p < RA) st(p, v[p]);
      });
    });
  }
}

template <int RA, int RB, bool INV, int LO, int HI, bool TWS = false>
CPB_D void pass_a_in(cplx (&v)[RA], int a, const cplx* CPB_RESTRICT tw, cplx* Sb, int LD) {
  pass_a_st<RA, RB, INV, LO, HI, TWS>(v, a, tw, [&](int p, cplx o) { Sb[(p * RB + a) * LD] = o; });
}

template <int RA, int RB, bool INV, bool TWS = false>
CPB_D void pass_a(cplx (&v)[RA], int a, const cplx* CPB_RESTRICT tw, cplx* Sb, int LD) {
  pass_a_in<RA, RB, INV, 0, RA, TWS>(v, a, tw, Sb, LD);
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

// Variants for kernels whose second-pass role is fixed over the pair loop (the z kernels): the first pass stores
// its outputs untwiddled, the second pass multiplies by the twiddles w^(a p) of its role p, which it holds in
// registers for the whole loop (tb[a], a >= 1) - no twiddle load inside the loop, and the multiplications move
// from the pass that occupies RB of the max(RA,RB) role rows to the one that occupies RA.
template <int RA, int RB, bool INV, int LO, int HI>
CPB_D void pass_a_raw(cplx (&v)[RA], int a, cplx* Sb, int LD) {
  dft_in<RA, INV, LO, HI>(v);
  static_for<0, RA>([&](auto pp) {
    constexpr int p = decltype(pp)::value;
    Sb[(p * RB + a) * LD] = v[p];
  });
}
// tb[a] = exp(+2 pi i a p / N); INV: multiply by tb, else by conj(tb)
template <int RA, int RB, bool INV>
CPB_D void pass_b_tw(cplx (&u)[RB], int p, const cplx* Sb, int LD, const cplx (&tb)[RB]) {
  static_for<0, RB>([&](auto aa) {
    constexpr int a = decltype(aa)::value;
    u[a] = Sb[(p * RB + a) * LD];
    if constexpr (a != 0) u[a] = INV ? cmul(u[a], tb[a]) : cmulc(u[a], tb[a]);
  });
  dft<RB, INV>(u);
}
// forward first pass of role a with the same register twiddles: dft, times conj(tb[p]), store
template <int RA, int RB>
CPB_D void pass_a_fwd_tw(cplx (&v)[RA], int a, cplx* Sb, int LD, const cplx (&tb)[RA]) {
  dft<RA, false>(v);
  static_for<0, RA>([&](auto pp) {
    constexpr int p = decltype(pp)::value;
    cplx o = v[p];
    if constexpr (p != 0) o = cmulc(o, tb[p]);
    Sb[(p * RB + a) * LD] = o;
  });
}
#ifndef CPB_Z_TWREG
#define CPB_Z_TWREG 1  // z kernels: twiddles of the fixed real-space-side role in registers (0: shared-memory table)
#endif

// ---------------------------------------------------------------------------------------------
// x passes.  Positions along a ray are addressed like the reference's compressed ray storage
// psi(kr1s, msrays) (fftprp_utils.mod.F90:269-285) restricted to the x band that holds
// coefficients and stored ray-minor: pos = xb * nrp + ray (xb = x - xlo, ray = internal ray index,
// row pitch nrp).  Three plan tables use it: gtab[pos] = plane wave stored at that position
// (inverse direction: the x kernel gathers c0 itself), gpos/gneg[ig] = positions of +G / -G
// (forward direction: k_x_fwd writes the band-ray storage G, k_unpack reads it).
//
// One block = SL consecutive rays, loops over a group of packed pairs.  Two thread roles:
//   "slot-major" (slot = tid % SL, row = tid / SL): lanes run along rays;
//   "x-major"    (row = tid % R1, slot = tid / R1): lanes run along x -> T1 is touched in
//                128-byte rows (consecutive lanes = consecutive x of one ray).
// The exchange buffer SX between the two radix passes is laid out so that both roles access it
// conflict free (odd row pitch).
// grid = (ray tiles, pair groups), block = SL * max(R1,R2)
// ---------------------------------------------------------------------------------------------
template <int R1, int R2, int SL>
struct XCfg {
  static constexpr int N = R1 * R2;
  static constexpr int RM = MaxOf<R1, R2>::v;
  static constexpr int NT = SL * RM;
  static constexpr int P1 = R1 | 1;  // odd pitch of SX rows
  static constexpr int SX_ELEMS = R2 * SL * P1;
  static constexpr int MINB = (NT <= 128) ? ((RM <= 16) ? 4 : 3) : 2;  // inverse kernel
  static constexpr int MINB_FWD = (NT <= 128) ? 3 : 2;
  // forward kernel: double-buffered exchange
  static constexpr size_t SMEM_FWD = (size_t)(2 * SX_ELEMS) * sizeof(cplx);
  // inverse kernel: one exchange buffer + twiddles + the gather stage [2 states][kcnt][R2*SL threads]
  static constexpr int NA = R2 * SL;  // threads with a slot-major role in the first pass
  static constexpr size_t smem_inv(int kcnt) { return (size_t)(SX_ELEMS + N + 2 * kcnt * NA) * sizeof(cplx); }
};

#ifndef CPB_X_ROT
#define CPB_X_ROT 1
#endif
constexpr uint32_t kNoPW = 0xffffffffu;  // gtab: position holds no plane wave
constexpr uint32_t kNegPW = 0x80000000u; // gtab: position holds the -G partner of plane wave (value & ~kNegPW)

// x pass, inverse: gather of the pair's coefficients + pack + FFT along x.  Fuses zeroing(psi) +
// set_psi_2_states_g / set_psi_1_state_g (state_utils.mod.F90:132-189:  psi(+G) = c1 + i c2,
// psi(-G) = conj(c1) + i conj(c2), G = 0 stored once) with the x mltfft of fftnew
// (fftmain_utils.mod.F90:93-94).  Every slot-major thread owns the band positions
// x = rA + R2*k of its ray: it issues 16-byte cp.async gathers of c1[ig], c2[ig] for them into its
// private slots of the shared-memory stage one pair ahead, and combines them when it starts the
// pair - no other thread touches those slots, so the gather needs no block barrier, holds no
// registers while in flight and the coefficients come straight from c0 (no packed copy in
// memory).  Positions without a plane wave are zeroed once.
// GK = true (tauofr / vtaupsi, dpsisc in tauofr_utils.mod.F90:113-137): every coefficient is multiplied by
// gk[3*ig] (one Cartesian component of G, the caller passes gk + direction) at +G and by -gk[3*ig] at -G.
template <int R1, int R2, int SL, int B, bool HALF, bool GK = false>
CPB_GLOBAL CPB_LAUNCH_BOUNDS((XCfg<R1, R2, SL>::NT), (XCfg<R1, R2, SL>::MINB))
    k_x_inv(const cplx* CPB_RESTRICT c0, long ldc, cplx* CPB_RESTRICT T1, PlanDev pd, PairDev pr, int npair,
            int ppg, const double* CPB_RESTRICT gk = nullptr) {
  using C = XCfg<R1, R2, SL>;
  using KR = KRange<R1, HALF>;
  constexpr int N = C::N, NT = C::NT, P1 = C::P1;
  CPB_DYN_SMEM(cplx, S);
  cplx* SX = S;
  cplx* TW = S + C::SX_ELEMS;
  cplx* ST = TW + N;  // [state][k][tid < NA]
  constexpr int NA = C::NA, KC = KR::cnt;
  const int tid = threadIdx.x;
  const int p0 = blockIdx.y * ppg;
  const int p1 = (p0 + ppg < npair) ? p0 + ppg : npair;
#if CPB_X_ROT
  // The slot-major pass needs R2 of the max(R1,R2) role rows, i.e. one warp of the block idles in it.
  // Which physical warp that is depends on the block index, so that the blocks resident on an SM
  // leave a different sub-partition idle instead of all the same one.
  const int tidA = (NT % 32 == 0) ? (int)((tid + 32 * (blockIdx.x % (NT / 32))) % NT) : tid;
#else
  const int tidA = tid;
#endif
  const int slotA = tidA % SL, rA = tidA / SL;  // slot-major role
  const int pB = tid % R1, slotB = tid / R1;  // x-major role (valid if slotB < SL)
  const int rayA = blockIdx.x * SL + slotA;
  const int rayB = blockIdx.x * SL + slotB;
  const bool okB = slotB < SL && rayB < pd.nrays;
  const size_t t1_pair = (size_t)pd.nxt * pd.nrays * B;
  pdl_trigger();
  for (int i = tid; i < N; i += NT) TW[i] = pd.tw1[i];
  pdl_wait();  // everything below may touch what the preceding kernel produced (or still reads: T1)
  // my band positions -> plane-wave index (bit 31: -G partner) or kNoPW
  uint32_t tab[KR::cnt];
  static_for<0, KR::cnt>([&](auto kk) {
    constexpr int j = decltype(kk)::value;
    const int xb = rA + R2 * (KR::lo + j) - pd.xlo;
    tab[j] = (rA < R2 && xb >= 0 && xb < pd.nxb) ? __ldg(&pd.gtab[(size_t)xb * pd.nrp + rayA]) : kNoPW;
    if (rA < R2 && tab[j] == kNoPW) {
      ST[(0 * KC + j) * NA + tidA] = mk(0.0, 0.0);
      ST[(1 * KC + j) * NA + tidA] = mk(0.0, 0.0);
    }
  });
  auto gather = [&](int pair) {
    const int s1 = __ldg(&pr.st1[pair]), s2 = __ldg(&pr.st2[pair]);
    const cplx* c1p = c0 + (size_t)s1 * ldc;
    const cplx* c2p = c0 + (size_t)(s2 < 0 ? s1 : s2) * ldc;
    static_for<0, KR::cnt>([&](auto kk) {
      constexpr int j = decltype(kk)::value;
      if (tab[j] != kNoPW) {
        const uint32_t ig = tab[j] & ~kNegPW;
        cp_async16(&ST[(0 * KC + j) * NA + tidA], c1p + ig);
        if (s2 >= 0) cp_async16(&ST[(1 * KC + j) * NA + tidA], c2p + ig);
        else ST[(1 * KC + j) * NA + tidA] = mk(0.0, 0.0);  // single-state path: c2 = 0
      }
    });
    cp_async_commit();
  };
  if (rA < R2 && p0 < p1) gather(p0);
  __syncthreads();  // twiddles visible
  for (int pair = p0; pair < p1; ++pair) {
    if (rA < R2) {
      cp_async_wait_all();
      cplx v[R1];
      static_for<0, R1>([&](auto kk) {
        constexpr int k = decltype(kk)::value;
        if constexpr (k >= KR::lo && k < KR::hi) {
          constexpr int j = k - KR::lo;
          const cplx a = ST[(0 * KC + j) * NA + tidA];
          const cplx bq = ST[(1 * KC + j) * NA + tidA];
          const double sg = (tab[j] & kNegPW) ? -1.0 : 1.0;
          // +G: c1 + i c2 = (a.x - b.y, a.y + b.x);  -G: conj(c1) + i conj(c2) = (a.x + b.y, b.x - a.y)
          v[k] = mk(a.x - sg * bq.y, sg * a.y + bq.x);
          if constexpr (GK) {
            const double g = (tab[j] != kNoPW) ? sg * __ldg(&gk[3 * (size_t)(tab[j] & ~kNegPW)]) : 0.0;
            v[k] = cscale(v[k], g);
          }
        } else {
          v[k] = mk(0.0, 0.0);
        }
      });
      if (pair + 1 < p1) gather(pair + 1);  // my slots are free again: next pair's gather
      cplx* dst = SX + (rA * SL + slotA) * P1;
      pass_a_st<R1, R2, true, KR::lo, KR::hi, true>(v, rA, TW, [&](int p, cplx o) { dst[p] = o; });
    }
    __syncthreads();
    if (slotB < SL) {
      cplx u[R2];
      static_for<0, R2>([&](auto aa) {
        constexpr int a = decltype(aa)::value;
        u[a] = SX[(a * SL + slotB) * P1 + pB];
      });
      dft<R2, true>(u);
      if (okB) {
        cplx* dst = T1 + (size_t)pair * t1_pair + (size_t)rayB * B;
        static_for<0, R2>([&](auto qq) {
          constexpr int q = decltype(qq)::value;
          const int x = pB + R1 * q;
          st_stream(&dst[(size_t)(x / B) * pd.nrays * B + (x % B)], u[q]);
        });
      }
    }
    __syncthreads();  // SX is single buffered
  }
}

// x pass, forward: FFT along x of T1 with fwfftn's scale 1/(n1 n2 n3) (the `scale` argument of the
// last mltfft, fftmain_utils.mod.F90:134-136); only the band rows are stored, into the band-ray
// storage that k_unpack reads.
template <int R1, int R2, int SL, int B, bool HALF>
CPB_GLOBAL CPB_LAUNCH_BOUNDS((XCfg<R1, R2, SL>::NT), (XCfg<R1, R2, SL>::MINB_FWD))
    k_x_fwd(const cplx* CPB_RESTRICT T1, cplx* CPB_RESTRICT G, PlanDev pd, int npair, int ppg) {
  using C = XCfg<R1, R2, SL>;
  using KR = KRange<R1, HALF>;
  constexpr int P1 = C::P1;
  CPB_DYN_SMEM(cplx, S);
  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x;
  const int p0 = blockIdx.y * ppg;
  const int p1 = (p0 + ppg < npair) ? p0 + ppg : npair;
#if CPB_X_ROT
  constexpr int NTF = XCfg<R1, R2, SL>::NT;
  const int tidA = (NTF % 32 == 0) ? (int)((tid + 32 * (blockIdx.x % (NTF / 32))) % NTF) : tid;  // see k_x_inv
#else
  const int tidA = tid;
#endif
  const int slotA = tidA % SL, rA = tidA / SL;
  const int pB = tid % R1, slotB = tid / R1;
  const int rayA = blockIdx.x * SL + slotA;
  const int rayB = blockIdx.x * SL + slotB;
  const bool okA = rA < R2 && rayA < pd.nrays;
  const bool okB = slotB < SL && rayB < pd.nrays;
  const size_t g_pair = (size_t)pd.nxb * pd.nrp;
  const size_t t1_pair = (size_t)pd.nxt * pd.nrays * B;
  const double sc = pd.inv_n;
  // x-major role: element k of the first (radix-R2) pass is x = pB + R1*k
  cplx nv[R2];
  auto fetch = [&](int pair) {
    const cplx* s = T1 + (size_t)pair * t1_pair + (size_t)(okB ? rayB : 0) * B;
    static_for<0, R2>([&](auto kk) {
      constexpr int k = decltype(kk)::value;
      const int x = pB + R1 * k;
      nv[k] = okB ? ld_stream(&s[(size_t)(x / B) * pd.nrays * B + (x % B)]) : mk(0.0, 0.0);
    });
  };
  if (slotB < SL && p0 < p1) fetch(p0);
  int buf = 0;
  for (int pair = p0; pair < p1; ++pair) {
    cplx* SX = S + buf * C::SX_ELEMS;
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
      if (okA) {
        cplx* d = G + (size_t)pair * g_pair + rayA;
        static_for<KR::lo, KR::hi>([&](auto tt) {
          constexpr int t = decltype(tt)::value;
          const int xb = rA + R2 * t - pd.xlo;
          if (xb >= 0 && xb < pd.nxb) d[(size_t)xb * pd.nrp] = cscale(u[t], sc);
        });
      }
    }
    buf ^= 1;
  }
}

// ---------------------------------------------------------------------------------------------
// Mirror-pair x passes (the Gamma-point hot path of rhoofr / vpsi).
//
// A plane wave G sits at (x, ray) and its partner -G at (n1 - x, mirror ray): the ray of (n2 - y, n3 - z).
// The internal ray numbering (z-major, dense in y) makes the mirror of ray r simply nrays - 1 - r (the
// plan verifies it).  A block owns H = SL/2 rays AND their mirrors (slots 0..H-1: rays H*bx + s, slots
// H..SL-1: their mirrors), so both halves of every +-G couple live in the same block:
//   k_x_inv_m   each coefficient c(ig) is fetched ONCE (by the thread that owns the +G position, x >=
//               n1/2) and the thread that owns the -G position reads it from the owner's shared-memory
//               stage: half the divergent 16-byte gathers of k_x_inv.  KIN: the owner also accumulates
//               hg |c|^2 and dotp's w |c|^2 (kin_energy_utils.mod.F90:62-110, dotp_utils.mod.F90:26-53),
//               so rhoofr needs no separate pass over c0.
//   k_x_fwd_m   after the forward x pass the block holds FFT[V psi] at +G and at -G: the +-G
//               separation, the kinetic term, the -f/2 scale and the c2 update (vpsi_utils.mod.F90:
//               626-673, add_wfn :717) happen right there - no band-ray storage in HBM, no k_unpack.
// Position classes along a ray (x = rA + R2 k): k < R1/2 lies below the centre plane (only -G partners or
// nothing), k > R1/2 above it (only +G or nothing), k == R1/2 contains the centre plane x = n1/2 (either).
// The staged / exchanged slots are indexed by k - R1/2 of the +G position, NPOS = KR::hi - R1/2 per thread.
// The centre ray (odd nrays) is its own mirror: its mirror slot is switched off and its -G positions
// read the ray's own slots.
// ---------------------------------------------------------------------------------------------
// gathers of the mirror-pair kernels: cp.async with an L2 evict-last hint (CPB_GATHER_KEEP=1) or plain
#ifndef CPB_GATHER_KEEP
#define CPB_GATHER_KEEP 0
#endif
#if CPB_GATHER_KEEP
#define CPB_GATHER_POLICY const uint64_t gather_pol = l2_keep_policy()
#define CPB_GATHER16(dst, src) cp_async16_keep(dst, src, gather_pol)
#else
#define CPB_GATHER_POLICY
#define CPB_GATHER16(dst, src) cp_async16(dst, src)
#endif
template <int R1, int R2, int SL, bool HALF>
struct XMCfg {
  using C = XCfg<R1, R2, SL>;
  using KR = KRange<R1, HALF>;
  static constexpr int H = SL / 2;
  static constexpr int C0 = R1 / 2;              // first decimated index that can hold a +G position
  static constexpr int NPOS = KR::hi - C0;
  static constexpr int NA = C::NA;
  static constexpr int NW = (C::NT + 31) / 32;
  // inverse: exchange + twiddles + 2 gather stages [state][NPOS][NA] (+ reduction scratch with KIN)
  // (+ the block's pair descriptors: 2 ints + 2 doubles per pair)
  static constexpr size_t SMEM_INV = (size_t)(C::SX_ELEMS + C::N + 2 * 2 * NPOS * NA) * sizeof(cplx) +
                                     (size_t)(4 * NA + 2 * 4 * 8 + 3 * kMaxGroup) * sizeof(double);
  // forward: exchange + twiddles + partner exchange [NPOS][NA] + staged c0 / c2 values [4][NPOS][NA]
  static constexpr size_t SMEM_FWD = (size_t)(C::SX_ELEMS + C::N + NPOS * NA + 4 * NPOS * NA) * sizeof(cplx) +
                                     (size_t)(3 * kMaxGroup) * sizeof(double);
  static constexpr int MINB_INV = (C::NT <= 128) ? ((C::RM <= 16) ? 4 : 3) : 2;
  static constexpr int MINB_FWD = (C::NT <= 128) ? 3 : 2;
};

// ray of slot s of block bx; valid = false: the slot does nothing.  self = true: the ray is its own mirror
// (the slot reads its own stage for the -G positions); otherwise the partner is slot (s + H) % SL.
struct XSlot {
  int ray;
  bool valid, self;
};
template <int SL>
CPB_D XSlot x_slot(int bx, int s, int nrays) {
  constexpr int H = SL / 2;
  const int base = bx * H + (s % H);
  const int nhalf = (nrays + 1) / 2;
  XSlot o;
  o.self = (2 * base == nrays - 1);
  o.valid = base < nhalf && !(s >= H && o.self);
  o.ray = (s < H) ? base : nrays - 1 - base;
  return o;
}

template <int R1, int R2, int SL, int B, bool HALF, bool KIN>
CPB_GLOBAL CPB_LAUNCH_BOUNDS((XCfg<R1, R2, SL>::NT), (XMCfg<R1, R2, SL, HALF>::MINB_INV))
    k_x_inv_m(const cplx* CPB_RESTRICT c0, long ldc, cplx* CPB_RESTRICT T1, PlanDev pd, PairDev pr, int npair,
              int ppg, double* CPB_RESTRICT kin_part, int geq0) {
  using C = XCfg<R1, R2, SL>;
  using M = XMCfg<R1, R2, SL, HALF>;
  using KR = KRange<R1, HALF>;
  constexpr int N = C::N, NT = C::NT, P1 = C::P1, NA = C::NA, H = M::H, C0 = M::C0, NPOS = M::NPOS;
  CPB_DYN_SMEM(cplx, S);
  cplx* SX = S;
  cplx* TW = S + C::SX_ELEMS;
  cplx* ST = TW + N;                                     // [buf][state][NPOS][NA]
  double* RED = reinterpret_cast<double*>(ST + 2 * 2 * NPOS * NA);   // [4][NA]
  double* RED2 = RED + 4 * NA;                           // [2][4][8], alternating per pair
  int* PS1 = reinterpret_cast<int*>(RED2 + 2 * 4 * 8);   // the block's pair descriptors [kMaxGroup] each
  int* PS2 = PS1 + kMaxGroup;
  CPB_GATHER_POLICY;
  const int tid = threadIdx.x;
  const int p0 = blockIdx.y * ppg;
  const int p1 = (p0 + ppg < npair) ? p0 + ppg : npair;
  pdl_trigger();
  for (int i = tid; i < N; i += NT) TW[i] = pd.tw1[i];
  pdl_wait();  // everything below may touch what the preceding kernel produced (or still reads: T1)
  for (int i = tid; i < p1 - p0; i += NT) {
    PS1[i] = __ldg(&pr.st1[p0 + i]);
    PS2[i] = __ldg(&pr.st2[p0 + i]);
  }
#if CPB_X_ROT
  const int tidA = (NT % 32 == 0) ? (int)((tid + 32 * (blockIdx.x % (NT / 32))) % NT) : tid;  // see k_x_inv
#else
  const int tidA = tid;
#endif
  const int slotA = tidA % SL, rA = tidA / SL;           // slot-major role
  const int pB = tid % R1, slotB = tid / R1;             // x-major role (valid if slotB < SL)
  const XSlot sa = x_slot<SL>(blockIdx.x, slotA, pd.nrays);
  const XSlot sb = x_slot<SL>(blockIdx.x, slotB < SL ? slotB : 0, pd.nrays);
  const bool actA = rA < R2;
  const bool okB = slotB < SL && sb.valid;
  const int spA = sa.self ? slotA : (slotA + H) % SL;    // slot that stages my -G partners
  const size_t t1_pair = (size_t)pd.nxt * pd.nrays * B;
  // my band positions -> plane-wave index (bit 31: -G partner) or kNoPW
  uint32_t tab[KR::cnt];
  double hgv[NPOS];
  static_for<0, KR::cnt>([&](auto kk) {
    constexpr int j = decltype(kk)::value;
    constexpr int k = KR::lo + j;
    const int xb = rA + R2 * k - pd.xlo;
    tab[j] = (actA && sa.valid && xb >= 0 && xb < pd.nxb) ? __ldg(&pd.gtab[(size_t)xb * pd.nrp + sa.ray]) : kNoPW;
    if constexpr (k < C0) {
      if (!(tab[j] & kNegPW)) tab[j] = kNoPW;            // below the centre plane only -G partners exist
    }
    if constexpr (KIN && k >= C0) {
      hgv[k - C0] = (tab[j] != kNoPW && !(tab[j] & kNegPW)) ? __ldg(&pd.hg[tab[j]]) : 0.0;
    }
  });
  auto gather = [&](int pair, int buf) {
    const int s1 = PS1[pair - p0], s2 = PS2[pair - p0];
    const cplx* c1p = c0 + (size_t)s1 * ldc;
    const cplx* c2p = c0 + (size_t)(s2 < 0 ? s1 : s2) * ldc;
    cplx* st = ST + (size_t)buf * (2 * NPOS * NA);
    static_for<C0, KR::hi>([&](auto kk) {
      constexpr int k = decltype(kk)::value;
      constexpr int j = k - KR::lo;
      if (tab[j] != kNoPW && !(tab[j] & kNegPW)) {
        CPB_GATHER16(&st[(0 * NPOS + (k - C0)) * NA + tidA], c1p + tab[j]);
        if (s2 >= 0) CPB_GATHER16(&st[(1 * NPOS + (k - C0)) * NA + tidA], c2p + tab[j]);
      }
    });
    cp_async_commit();
  };
  __syncthreads();  // pair descriptors visible
  if (actA && p0 < p1) gather(p0, 0);
  for (int pair = p0; pair < p1; ++pair) {
    const int buf = (pair - p0) & 1;
    if (actA) cp_async_wait_all();
    __syncthreads();  // the stage of this pair is complete and visible to the partner threads; SX and TW are free / ready
    if (actA) {
      if (pair + 1 < p1) gather(pair + 1, buf ^ 1);      // the other stage was last read one iteration ago
      const bool two = PS2[pair - p0] >= 0;
      const cplx* st = ST + (size_t)buf * (2 * NPOS * NA);
      double sk1 = 0.0, sd1 = 0.0, sk2 = 0.0, sd2 = 0.0;
      cplx v[R1];
      static_for<0, R1>([&](auto kk) {
        constexpr int k = decltype(kk)::value;
        if constexpr (k >= KR::lo && k < KR::hi) {
          constexpr int j = k - KR::lo;
          const uint32_t e = tab[j];
          v[k] = mk(0.0, 0.0);
          if (e != kNoPW) {
            if (k >= C0 && !(e & kNegPW)) {
              // +G: c1 + i c2
              const cplx a = st[(0 * NPOS + (k - C0)) * NA + tidA];
              const cplx bq = two ? st[(1 * NPOS + (k - C0)) * NA + tidA] : mk(0.0, 0.0);
              v[k] = mk(a.x - bq.y, a.y + bq.x);
              if constexpr (KIN && k >= C0) {
                const double m1 = a.x * a.x + a.y * a.y, m2 = bq.x * bq.x + bq.y * bq.y;
                const bool g0 = (e == 0u) && geq0;       // dotp counts the real part of G = 0 once
                sk1 += hgv[k - C0] * m1;
                sk2 += hgv[k - C0] * m2;
                sd1 += g0 ? a.x * a.x : 2.0 * m1;
                sd2 += g0 ? bq.x * bq.x : 2.0 * m2;
              }
            } else {
              // -G: conj(c1) + i conj(c2) of the coefficient staged by the owner of (n1 - x, mirror ray)
              const int xm = pd.n1 - (rA + R2 * k);
              const int idx = (xm / R2 - C0) * NA + (xm % R2) * SL + spA;
              const cplx a = st[0 * NPOS * NA + idx];
              const cplx bq = two ? st[1 * NPOS * NA + idx] : mk(0.0, 0.0);
              v[k] = mk(a.x + bq.y, bq.x - a.y);
            }
          }
        } else {
          v[k] = mk(0.0, 0.0);
        }
      });
      if constexpr (KIN) {
        RED[0 * NA + tidA] = sk1;
        RED[1 * NA + tidA] = sd1;
        RED[2 * NA + tidA] = sk2;
        RED[3 * NA + tidA] = sd2;
      }
      cplx* dst = SX + (rA * SL + slotA) * P1;
      pass_a_st<R1, R2, true, KR::lo, KR::hi, true>(v, rA, TW, [&](int p, cplx o) { dst[p] = o; });
    }
    __syncthreads();
    if constexpr (KIN) {
      // fixed-order reduction of the NA per-thread partials in three steps (NA -> 8 -> 1 per quantity); the
      // last step of a pair runs one iteration later, so no barrier is added
      if (pair > p0 && tid < 4) {
        const double* r2 = RED2 + ((pair - 1 - p0) & 1) * 32;
        double s = 0.0;
        for (int i = 0; i < 8; ++i) s += r2[tid * 8 + i];
        kin_part[((size_t)(pair - 1) * gridDim.x + blockIdx.x) * 4 + tid] = s;
      }
    }
    if (slotB < SL) {
      cplx u[R2];
      static_for<0, R2>([&](auto aa) {
        constexpr int a = decltype(aa)::value;
        u[a] = SX[(a * SL + slotB) * P1 + pB];
      });
      dft<R2, true>(u);
      if (okB) {
        cplx* dst = T1 + (size_t)pair * t1_pair + (size_t)sb.ray * B;
        static_for<0, R2>([&](auto qq) {
          constexpr int q = decltype(qq)::value;
          const int x = pB + R1 * q;
          st_stream(&dst[(size_t)(x / B) * pd.nrays * B + (x % B)], u[q]);
        });
      }
    }
    if constexpr (KIN) {
      if (tid < 32) {
        const int q = tid / 8, part = tid % 8;
        constexpr int PER = (NA + 7) / 8;
        double s = 0.0;
        for (int i = part * PER; i < (part + 1) * PER && i < NA; ++i) s += RED[q * NA + i];
        RED2[((pair - p0) & 1) * 32 + q * 8 + part] = s;
      }
    }
  }
  if constexpr (KIN) {
    __syncthreads();
    if (p1 > p0 && tid < 4) {
      const double* r2 = RED2 + ((p1 - 1 - p0) & 1) * 32;
      double s = 0.0;
      for (int i = 0; i < 8; ++i) s += r2[tid * 8 + i];
      kin_part[((size_t)(p1 - 1) * gridDim.x + blockIdx.x) * 4 + tid] = s;
    }
  }
}

// forward x pass fused with the unpack of vpsi (see the header comment above).  pr.ca / pr.cb = fi / fip1
// (vpsi_utils.mod.F90:627-633).  ACC: c2 += result (reference semantics), else c2 = result.
template <int R1, int R2, int SL, int B, bool HALF, bool ACC>
CPB_GLOBAL CPB_LAUNCH_BOUNDS((XCfg<R1, R2, SL>::NT), (XMCfg<R1, R2, SL, HALF>::MINB_FWD))
    k_x_fwd_m(const cplx* CPB_RESTRICT T1, const cplx* CPB_RESTRICT c0, cplx* c2, long ldc, PlanDev pd, PairDev pr,
              int npair, int ppg) {
  using C = XCfg<R1, R2, SL>;
  using M = XMCfg<R1, R2, SL, HALF>;
  using KR = KRange<R1, HALF>;
  constexpr int NT = C::NT, P1 = C::P1, NA = C::NA, H = M::H, C0 = M::C0, NPOS = M::NPOS;
  CPB_DYN_SMEM(cplx, S);
  cplx* SX = S;
  cplx* TW = S + C::SX_ELEMS;
  cplx* EX = TW + C::N;               // [NPOS][NA]: FFT[V psi] at the -G partner of my +G positions
  cplx* CS = EX + NPOS * NA;          // [4][NPOS][NA]: c0(ig, s1), c0(ig, s2), c2(ig, s1), c2(ig, s2)
  double* PCA = reinterpret_cast<double*>(CS + 4 * NPOS * NA);   // the block's pair descriptors [kMaxGroup] each
  double* PCB = PCA + kMaxGroup;
  int* PS1 = reinterpret_cast<int*>(PCB + kMaxGroup);
  int* PS2 = PS1 + kMaxGroup;
  CPB_GATHER_POLICY;
  const int tid = threadIdx.x;
  const int p0 = blockIdx.y * ppg;
  const int p1 = (p0 + ppg < npair) ? p0 + ppg : npair;
  pdl_trigger();
  for (int i = tid; i < C::N; i += NT) TW[i] = pd.tw1[i];
  pdl_wait();  // everything below may touch what the preceding kernel produced
  for (int i = tid; i < p1 - p0; i += NT) {
    PS1[i] = __ldg(&pr.st1[p0 + i]);
    PS2[i] = __ldg(&pr.st2[p0 + i]);
    PCA[i] = __ldg(&pr.ca[p0 + i]);
    PCB[i] = __ldg(&pr.cb[p0 + i]);
  }
#if CPB_X_ROT
  const int tidA = (NT % 32 == 0) ? (int)((tid + 32 * (blockIdx.x % (NT / 32))) % NT) : tid;  // see k_x_inv
#else
  const int tidA = tid;
#endif
  const int slotA = tidA % SL, rA = tidA / SL;
  const int pB = tid % R1, slotB = tid / R1;
  const XSlot sa = x_slot<SL>(blockIdx.x, slotA, pd.nrays);
  const XSlot sb = x_slot<SL>(blockIdx.x, slotB < SL ? slotB : 0, pd.nrays);
  const bool actA = rA < R2;
  const bool okB = slotB < SL && sb.valid;
  const int spA = sa.self ? slotA : (slotA + H) % SL;
  #if defined(CPB_DEBUG_KNOBS) && (CPB_DEBUG_KNOBS & 16)
  const size_t t1_pair = 0;  // experiment: every pair reads the same (L2-resident) T1 region
#else
  const size_t t1_pair = (size_t)pd.nxt * pd.nrays * B;
#endif
  const double sc = pd.inv_n;
  uint32_t tab[KR::cnt];
  double g2v[NPOS];
  static_for<0, KR::cnt>([&](auto kk) {
    constexpr int j = decltype(kk)::value;
    constexpr int k = KR::lo + j;
    const int xb = rA + R2 * k - pd.xlo;
    tab[j] = (actA && sa.valid && xb >= 0 && xb < pd.nxb) ? __ldg(&pd.gtab[(size_t)xb * pd.nrp + sa.ray]) : kNoPW;
    if constexpr (k < C0) {
      if (!(tab[j] & kNegPW)) tab[j] = kNoPW;
    }
    if constexpr (k >= C0) {
      g2v[k - C0] = (tab[j] != kNoPW && !(tab[j] & kNegPW)) ? pd.tpiba2 * __ldg(&pd.hg[tab[j]]) : 0.0;
    }
  });
  // x-major role: element k of the first (radix-R2) pass is x = pB + R1*k
  cplx nv[R2];
  auto fetch = [&](int pair) {
    const cplx* s = T1 + (size_t)pair * t1_pair + (size_t)(okB ? sb.ray : 0) * B;
    static_for<0, R2>([&](auto kk) {
      constexpr int k = decltype(kk)::value;
      const int x = pB + R1 * k;
      nv[k] = okB ? ld_stream(&s[(size_t)(x / B) * pd.nrays * B + (x % B)]) : mk(0.0, 0.0);
    });
  };
  if (slotB < SL && p0 < p1) fetch(p0);
  __syncthreads();  // pair descriptors and twiddles visible
  for (int pair = p0; pair < p1; ++pair) {
    const int s1 = PS1[pair - p0], s2 = PS2[pair - p0];
    if (actA) {
      // stage the coefficients the unpack of this pair needs (own slots only: no barrier involved)
      const cplx* a1 = c0 + (size_t)s1 * ldc;
      const cplx* a2 = c0 + (size_t)(s2 < 0 ? s1 : s2) * ldc;
      const cplx* o1 = c2 + (size_t)s1 * ldc;
      const cplx* o2 = c2 + (size_t)(s2 < 0 ? s1 : s2) * ldc;
      static_for<C0, KR::hi>([&](auto kk) {
        constexpr int k = decltype(kk)::value;
        constexpr int j = k - KR::lo;
        if (tab[j] != kNoPW && !(tab[j] & kNegPW)) {
          const int o = (k - C0) * NA + tidA;
#if defined(CPB_DEBUG_KNOBS) && (CPB_DEBUG_KNOBS & 4)
          const uint32_t ig = tab[j] & 0x3ffu;  // experiment: the coefficient gathers read a 16 KB window
#else
          const uint32_t ig = tab[j];
#endif
          CPB_GATHER16(&CS[0 * NPOS * NA + o], a1 + ig);
          if (s2 >= 0) CPB_GATHER16(&CS[1 * NPOS * NA + o], a2 + ig);
          if constexpr (ACC) {
            CPB_GATHER16(&CS[2 * NPOS * NA + o], o1 + ig);
            if (s2 >= 0) CPB_GATHER16(&CS[3 * NPOS * NA + o], o2 + ig);
          }
        }
      });
      cp_async_commit();
    }
    if (slotB < SL) {
      cplx v[R2];
      static_for<0, R2>([&](auto kk) { v[decltype(kk)::value] = nv[decltype(kk)::value]; });
      // forward transform uses the mirrored factorisation (R2 first, then R1)
      pass_a_st<R2, R1, false, 0, R2, true>(v, pB, TW, [&](int p, cplx o) { SX[(p * SL + slotB) * P1 + pB] = o; });
      if (pair + 1 < p1) fetch(pair + 1);
    }
    __syncthreads();
    cplx up[NPOS];  // my outputs at the positions that can hold +G
    if (actA) {
      cplx u[R1];
      static_for<0, R1>([&](auto aa) {
        constexpr int a = decltype(aa)::value;
        u[a] = SX[(rA * SL + slotA) * P1 + a];
      });
      dft<R1, false>(u);
      static_for<KR::lo, KR::hi>([&](auto tt) {
        constexpr int t = decltype(tt)::value;
        constexpr int j = t - KR::lo;
        const cplx w = cscale(u[t], sc);
        if constexpr (t >= C0) up[t - C0] = w;
        if (tab[j] != kNoPW && (tab[j] & kNegPW)) {
          // hand FFT[V psi](-G) to the owner of the +G position (n1 - x, mirror ray)
          const int xm = pd.n1 - (rA + R2 * t);
          EX[(xm / R2 - C0) * NA + (xm % R2) * SL + spA] = w;
        }
      });
    }
    __syncthreads();
    if (actA) {
      cp_async_wait_all();
      const double fi = PCA[pair - p0], fip1 = PCB[pair - p0];
      static_for<C0, KR::hi>([&](auto tt) {
        constexpr int t = decltype(tt)::value;
        constexpr int j = t - KR::lo;
#if defined(CPB_DEBUG_KNOBS) && (CPB_DEBUG_KNOBS & 8)
        const uint32_t ig = (tab[j] == kNoPW || (tab[j] & kNegPW)) ? tab[j] : (tab[j] & 0x3ffu);  // experiment: c2 stores into a 16 KB window
#else
        const uint32_t ig = tab[j];
#endif
        if (ig != kNoPW && !(ig & kNegPW)) {
          const int o = (t - C0) * NA + tidA;
          const cplx psin = up[t - C0];
          // G = 0 is its own partner (vpsi_utils.mod.F90:655-671 reads psi(nzhs) and psi(indzs), the same element)
          const bool selfpos = sa.self && 2 * (rA + R2 * t) == pd.n1;
          const cplx psii = selfpos ? psin : EX[o];
          const cplx a = CS[0 * NPOS * NA + o];
          const cplx fp = cadd(psin, psii);
          const cplx fm = csub(psin, psii);
          const double g2 = g2v[t - C0];
          cplx r1 = mk(-fi * (g2 * a.x + fp.x), -fi * (g2 * a.y + fm.y));
          if constexpr (ACC) r1 = cadd(r1, CS[2 * NPOS * NA + o]);
          c2[(size_t)s1 * ldc + ig] = r1;
          if (s2 >= 0) {
            const cplx bq = CS[1 * NPOS * NA + o];
            cplx r2 = mk(-fip1 * (g2 * bq.x + fp.y), -fip1 * (g2 * bq.y - fm.x));
            if constexpr (ACC) r2 = cadd(r2, CS[3 * NPOS * NA + o]);
            c2[(size_t)s2 * ldc + ig] = r2;
          }
        }
      });
    }
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
// Staged variants of the y/z kernels ("bulk" kernels).  Where a block's input tile of one pair is
// one contiguous run of memory (the band of a column set in T2, the rays of a plane in T1) it is
// brought into a ring of KStages shared-memory stages by the TMA engine (bulk_g2s, one instruction
// per tile issued by one thread) instead of per-thread register prefetches: the prefetch distance
// is KStages pairs, no registers are held by in-flight loads and no per-element address arithmetic
// is executed.  The twiddle table lives in shared memory too, so nothing inside the pair loop
// waits on a global load.  Role rotation: the first radix pass only needs R2 of the max(R1,R2)
// role rows of a block; which warp idles rotates with the pair index so that the four SM
// sub-partitions carry the same FP64 load.
// Shared memory: [XB exchange buffers N*B][twiddles N][KStages tiles][KStages mbarriers].
// XB = 2: one block barrier per transform; XB = 1: two barriers, but the smaller footprint (and a
// 128-register budget) lets a fourth block share the SM.
// ---------------------------------------------------------------------------------------------
constexpr int kStages = 2;
#ifndef CPB_Y_ZFAST
#define CPB_Y_ZFAST 1
#endif
#ifndef CPB_ZRHO_XB
#define CPB_ZRHO_XB 2
#endif
#ifndef CPB_YFWD_ASYNC
#define CPB_YFWD_ASYNC 1  // k_y_fwd: next pair's rows by cp.async into private slots (0: register prefetch)
#endif
#ifndef CPB_YINV_XB
#define CPB_YINV_XB 2
#endif
// resident blocks per SM the staged kernels with XB exchange buffers are compiled for
template <int R1, int R2, int XB>
struct YZBlocksX {
  static constexpr int v = YZBlocks<R1, R2>::v + ((XB == 1 && YZBlocks<R1, R2>::v == 3) ? 1 : 0);
};

template <int R1, int R2, int B>
struct YZCfg {
  static constexpr int N = R1 * R2;
  static constexpr int RM = MaxOf<R1, R2>::v;
  static constexpr int NT = B * RM;
  static constexpr int RPW = 32 / B;  // role rows per warp
  static constexpr bool ROT = (NT % 32 == 0) && (RM % RPW == 0) && (R2 < RM);
  // bytes of dynamic shared memory for tiles of `tile_elems` complex numbers, `xb` exchange buffers
  static constexpr size_t smem(int tile_elems, int xb = 2) {
    return ((size_t)xb * N * B + N + (size_t)kStages * tile_elems) * sizeof(cplx) + kStages * sizeof(uint64_t);
  }
};

// Exchange buffers of the staged y/z kernels: two (one block barrier per transform) unless that leaves room for
// only one block per SM (lengths from 288 on: 2 x 16 N B bytes of exchange + the TMA ring exceed half an SM's
// shared memory); then one buffer and one more barrier per transform, and two blocks per SM again.  The tile is
// taken as the dual = 4 band (half the axis).
template <int R1, int R2, int B>
struct YZXB {
  static constexpr size_t two = YZCfg<R1, R2, B>::smem((R1 * R2 / 2 + 1) * B, 2);
  static constexpr int v = (2 * (two + 1024) <= (size_t)228 * 1024) ? 2 : 1;
};

// ---------------------------------------------------------------------------------------------
// y pass, inverse.  Block = (x tile of the chunk, z plane of the band, group of pairs).  Reads the
// rays of the plane (zero outside [ylo,yhi]: unpack_x2y's zero fill, fftutil_utils.mod.F90:413-457),
// writes all n2 rows of the chunk's T2.  grid = (x tiles of the chunk, nzb, pair groups)
// ---------------------------------------------------------------------------------------------
template <int R1, int R2, int B, bool HALF, int XB>
CPB_GLOBAL CPB_LAUNCH_BOUNDS(B* MaxOf<R1, R2>::v, (YZBlocksX<R1, R2, XB>::v))
    k_y_inv(const cplx* CPB_RESTRICT T1, cplx* CPB_RESTRICT T2, PlanDev pd, int xt0, int npair, int ppg) {
  using C = YZCfg<R1, R2, B>;
  constexpr int N = C::N, RM = C::RM, NT = C::NT;
  using KR = KRange<R1, HALF>;
  CPB_DYN_SMEM(cplx, S);
  cplx* TW = S + XB * N * B;
  cplx* ST = TW + N;
  const int tile_elems = pd.nyb * B;
  uint64_t* bar = reinterpret_cast<uint64_t*>(ST + (size_t)kStages * tile_elems);
  const int tid = threadIdx.x;
  const int b = tid % B, r = tid / B;
#if CPB_Y_ZFAST
  // z plane is the fastest-varying block coordinate: blocks that run at the same time touch
  // adjacent 128-byte rows of T2, which L2 merges into full DRAM pages
  const int zr = blockIdx.x;
  const int xtc = blockIdx.y, nxc = gridDim.y;
#else
  const int xtc = blockIdx.x, nxc = gridDim.x;
  const int zr = blockIdx.y;
#endif
  const int p0 = blockIdx.z * ppg;
  const int p1 = (p0 + ppg < npair) ? p0 + ppg : npair;
  const int ylo = pd.ylo[zr], yhi = pd.yhi[zr];
  const int ny = yhi - ylo + 1;  // rays of this plane (<= 0: none), one contiguous run of T1
  const unsigned tile_bytes = (unsigned)(ny > 0 ? ny : 0) * B * (unsigned)sizeof(cplx);
  const size_t t1_pair = (size_t)pd.nxt * pd.nrays * B;
  const size_t t2_pair = (size_t)nxc * N * pd.nzb * B;
  const cplx* src = T1 + ((size_t)(xt0 + xtc) * pd.nrays + pd.rayoff[zr]) * B;
  cplx* dst = T2 + ((size_t)xtc * N * pd.nzb + zr) * B + b;
  pdl_trigger();
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&bar[s], 1);
    mbar_fence_init();
  }
  for (int i = tid; i < N; i += NT) TW[i] = pd.tw2[i];
  pdl_wait();  // T1 comes from the preceding kernel; T2 may still be read by it
  __syncthreads();
  if (tid == 0 && ny > 0) {
    for (int s = 0; s < kStages && p0 + s < p1; ++s) {
      mbar_expect_tx(&bar[s], tile_bytes);
      bulk_g2s(ST + (size_t)s * tile_elems, src + (size_t)(p0 + s) * t1_pair, tile_bytes, &bar[s]);
    }
  }
  int rA = r;  // role in the first radix pass (rotates by one warp per pair)
  for (int pair = p0; pair < p1; ++pair) {
    const int it = pair - p0;
    const int st = it % kStages;
    cplx* Sb = S + (XB == 2 ? (it & 1) : 0) * (N * B) + b;
    if (rA < R2) {
      if (ny > 0) mbar_wait(&bar[st], (unsigned)((it / kStages) & 1));
      const cplx* in = ST + (size_t)st * tile_elems + b;
      cplx v[R1];
      static_for<0, R1>([&](auto kk) {
        constexpr int k = decltype(kk)::value;
        if constexpr (k >= KR::lo && k < KR::hi) {
          const int y = rA + R2 * k;
          v[k] = (y >= ylo && y <= yhi) ? in[(y - ylo) * B] : mk(0.0, 0.0);
        } else {
          v[k] = mk(0.0, 0.0);
        }
      });
      pass_a_in<R1, R2, true, KR::lo, KR::hi, true>(v, rA, TW, Sb, B);
    }
    __syncthreads();
    if (tid == 0 && ny > 0 && pair + kStages < p1) {
      mbar_expect_tx(&bar[st], tile_bytes);
      bulk_g2s(ST + (size_t)st * tile_elems, src + (size_t)(pair + kStages) * t1_pair, tile_bytes, &bar[st]);
    }
    if (r < R1) {
      cplx u[R2];
      pass_b<R1, R2, true>(u, r, Sb, B);
      cplx* d = dst + (size_t)pair * t2_pair;
      static_for<0, R2>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        d[(size_t)(r + R1 * q) * pd.nzb * B] = u[q];
      });
    }
    if constexpr (XB == 1) __syncthreads();
    if constexpr (C::ROT) {
      rA += C::RPW;
      if (rA >= RM) rA -= RM;
    }
  }
}

// y pass, forward: reads all n2 rows of the chunk's T2, writes only the rays of the plane into T1.
// ASYNC: the next pair's rows travel by cp.async into private shared-memory slots instead of registers
template <int R1, int R2, int B, bool HALF, bool ASYNC>
CPB_GLOBAL CPB_LAUNCH_BOUNDS(B* MaxOf<R1, R2>::v, YZBlocks<R1, R2>::v)
    k_y_fwd(const cplx* CPB_RESTRICT T2, cplx* CPB_RESTRICT T1, PlanDev pd, int xt0, int npair, int ppg) {
  constexpr int N = R1 * R2;
  using KR = KRange<R1, HALF>;
  CPB_DYN_SMEM(cplx, S);
  pdl_trigger();
  const int tid = threadIdx.x;
  const int b = tid % B, r = tid / B;
#if CPB_Y_ZFAST
  // z plane is the fastest-varying block coordinate: blocks that run at the same time touch
  // adjacent 128-byte rows of T2, which L2 merges into full DRAM pages
  const int zr = blockIdx.x;
  const int xtc = blockIdx.y, nxc = gridDim.y;
#else
  const int xtc = blockIdx.x, nxc = gridDim.x;
  const int zr = blockIdx.y;
#endif
  const int p0 = blockIdx.z * ppg;
  const int p1 = (p0 + ppg < npair) ? p0 + ppg : npair;
  const int ylo = pd.ylo[zr], yhi = pd.yhi[zr];
  const size_t t1_pair = (size_t)pd.nxt * pd.nrays * B;
  const size_t t2_pair = (size_t)nxc * N * pd.nzb * B;
  const size_t ystride = (size_t)pd.nzb * B;
  pdl_wait();  // T2 comes from the preceding kernel
  const cplx* src = T2 + ((size_t)xtc * N * pd.nzb + zr) * B + b;
  cplx* dst = T1 + ((size_t)(xt0 + xtc) * pd.nrays + pd.rayoff[zr]) * B + b;
  // ASYNC: the rows of the NEXT pair travel with 16-byte cp.async copies into private shared-memory slots (a warp
  // instruction = 4 whole 128-byte rows): no registers are held by loads in flight
  cplx* NV = S + 2 * (N * B);  // [R2][threads]
  constexpr int nthr = B * MaxOf<R1, R2>::v;
  cplx nv[ASYNC ? 1 : R2];
  auto fetch = [&](int pair) {
    const cplx* s = src + (size_t)pair * t2_pair;
    static_for<0, R2>([&](auto kk) {
      constexpr int k = decltype(kk)::value;
      if constexpr (ASYNC) cp_async16(&NV[k * nthr + tid], &s[(size_t)(r + R1 * k) * ystride]);
      else nv[k] = s[(size_t)(r + R1 * k) * ystride];
    });
    if constexpr (ASYNC) cp_async_commit();
  };
  if (r < R1 && p0 < p1) fetch(p0);
  int buf = 0;
  for (int pair = p0; pair < p1; ++pair) {
    cplx* Sb = S + buf * (N * B) + b;
    if (r < R1) {
      cplx v[R2];
      if constexpr (ASYNC) {
        cp_async_wait_all();
        static_for<0, R2>([&](auto kk) { v[decltype(kk)::value] = NV[decltype(kk)::value * nthr + tid]; });
        if (pair + 1 < p1) fetch(pair + 1);  // my slots are free again
        pass_a<R2, R1, false>(v, r, pd.tw2, Sb, B);
      } else {
        static_for<0, R2>([&](auto kk) { v[decltype(kk)::value] = nv[decltype(kk)::value]; });
        pass_a<R2, R1, false>(v, r, pd.tw2, Sb, B);
        if (pair + 1 < p1) fetch(pair + 1);
      }
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
template <int R1, int R2, int B, bool HALF, int XB>
CPB_GLOBAL CPB_LAUNCH_BOUNDS(B* MaxOf<R1, R2>::v, (YZBlocksX<R1, R2, XB>::v))
    k_z_rho(const cplx* CPB_RESTRICT T2, double* rho, PlanDev pd, PairDev pr, int npair,
            int xt0) {
  using C = YZCfg<R1, R2, B>;
  constexpr int N = C::N, RM = C::RM, NT = C::NT;
  using KR = KRange<R1, HALF>;
  CPB_DYN_SMEM(cplx, S);
  cplx* TW = S + XB * N * B;
  cplx* ST = TW + N;
  const int tile_elems = pd.nzb * B;
  uint64_t* bar = reinterpret_cast<uint64_t*>(ST + (size_t)kStages * tile_elems);
  const unsigned tile_bytes = (unsigned)(tile_elems * sizeof(cplx));
  const int tid = threadIdx.x;
  const int b = tid % B, r = tid / B;
  const int xtc = blockIdx.x, nxc = gridDim.x;
  const int x = (xt0 + xtc) * B + b;
  const int y = blockIdx.y;
  const bool xok = x < pd.n1;
  const size_t pstride = (size_t)nxc * pd.n2 * pd.nzb * B;
  const cplx* tile = T2 + ((size_t)xtc * pd.n2 + y) * pd.nzb * B;
  const int zlo = pd.zlo, nzb = pd.nzb;
  pdl_trigger();
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&bar[s], 1);
    mbar_fence_init();
  }
#if CPB_Z_TWREG
  cplx tb[R2];  // twiddles of my real-space-side role r
  static_for<0, R2>([&](auto aa) {
    constexpr int a = decltype(aa)::value;
    tb[a] = (r < R1) ? __ldg(&pd.tw3[(a * r) % N]) : mk(1.0, 0.0);
  });
#else
  for (int i = tid; i < N; i += NT) TW[i] = pd.tw3[i];
#endif
  pdl_wait();  // T2 and rho come from preceding kernels
  // the accumulators start from rho itself: the read-modify-write's read overlaps the first tile
  double acc[R2];
  static_for<0, R2>([&](auto qq) {
    constexpr int q = decltype(qq)::value;
    acc[q] = (r < R1 && xok) ? rho[((size_t)(r + R1 * q) * pd.kr2 + y) * pd.kr1 + x] : 0.0;
  });
  __syncthreads();
  if (tid == 0) {
    for (int s = 0; s < kStages && s < npair; ++s) {
      mbar_expect_tx(&bar[s], tile_bytes);
      bulk_g2s(ST + (size_t)s * tile_elems, tile + (size_t)s * pstride, tile_bytes, &bar[s]);
    }
  }
  int rA = r;  // role in the first radix pass (rotates by one warp per pair)
  for (int pair = 0; pair < npair; ++pair) {
    const int st = pair % kStages;
    cplx* Sb = S + (XB == 2 ? (pair & 1) : 0) * (N * B) + b;
    const double ca = __ldg(&pr.ca[pair]), cb = __ldg(&pr.cb[pair]);  // used after the barrier
    if (rA < R2) {
      mbar_wait(&bar[st], (unsigned)((pair / kStages) & 1));
      const cplx* in = ST + (size_t)st * tile_elems + b;
      cplx v[R1];
      static_for<0, R1>([&](auto kk) {
        constexpr int k = decltype(kk)::value;
        if constexpr (k >= KR::lo && k < KR::hi) {
          const int zr = rA + R2 * k - zlo;
          v[k] = (zr >= 0 && zr < nzb) ? in[zr * B] : mk(0.0, 0.0);
        } else {
          v[k] = mk(0.0, 0.0);
        }
      });
#if CPB_Z_TWREG
      pass_a_raw<R1, R2, true, KR::lo, KR::hi>(v, rA, Sb, B);
#else
      pass_a_in<R1, R2, true, KR::lo, KR::hi, true>(v, rA, TW, Sb, B);
#endif
    }
    __syncthreads();
    if (tid == 0 && pair + kStages < npair) {
      mbar_expect_tx(&bar[st], tile_bytes);
      bulk_g2s(ST + (size_t)st * tile_elems, tile + (size_t)(pair + kStages) * pstride, tile_bytes, &bar[st]);
    }
    if (r < R1) {
      cplx u[R2];
#if CPB_Z_TWREG
      pass_b_tw<R1, R2, true>(u, r, Sb, B, tb);
#else
      pass_b<R1, R2, true>(u, r, Sb, B);
#endif
      static_for<0, R2>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        acc[q] += ca * (u[q].x * u[q].x) + cb * (u[q].y * u[q].y);
      });
    }
    if constexpr (XB == 1) __syncthreads();
    if constexpr (C::ROT) {
      rA += C::RPW;
      if (rA >= RM) rA -= RM;
    }
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
// XB = 2: separate exchange buffers for the inverse and the forward transform (two block barriers per pair);
// XB = 1: one buffer, two more barriers (the large lengths, see YZXB).
template <int R1, int R2, int B, bool HALF, int XB = 2>
CPB_GLOBAL CPB_LAUNCH_BOUNDS(B* MaxOf<R1, R2>::v, YZBlocks<R1, R2>::v)
    k_z_vpsi(cplx* T2, const double* CPB_RESTRICT vpot, PlanDev pd, int xt0, int npair, int ppg) {
  using C = YZCfg<R1, R2, B>;
  constexpr int N = C::N, RM = C::RM, NT = C::NT;
  using KR = KRange<R1, HALF>;
  CPB_DYN_SMEM(cplx, S);
  cplx* TW = S + XB * N * B;
  cplx* ST = TW + N;
  const int tile_elems = pd.nzb * B;
  uint64_t* bar = reinterpret_cast<uint64_t*>(ST + (size_t)kStages * tile_elems);
  const unsigned tile_bytes = (unsigned)(tile_elems * sizeof(cplx));
  const int tid = threadIdx.x;
  const int b = tid % B, r = tid / B;
  const int xtc = blockIdx.x, nxc = gridDim.x;
  const int x = (xt0 + xtc) * B + b;
  const int y = blockIdx.y;
  const int p0 = blockIdx.z * ppg;
  const int p1 = (p0 + ppg < npair) ? p0 + ppg : npair;
  const bool xok = x < pd.n1;
  const size_t pstride = (size_t)nxc * pd.n2 * pd.nzb * B;
  cplx* tile = T2 + ((size_t)xtc * pd.n2 + y) * pd.nzb * B;
  const int zlo = pd.zlo, nzb = pd.nzb;
  pdl_trigger();
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&bar[s], 1);
    mbar_fence_init();
  }
#if CPB_Z_TWREG
  cplx tb[R2];  // twiddles of my real-space-side role r
  static_for<0, R2>([&](auto aa) {
    constexpr int a = decltype(aa)::value;
    tb[a] = (r < R1) ? __ldg(&pd.tw3[(a * r) % N]) : mk(1.0, 0.0);
  });
#else
  for (int i = tid; i < N; i += NT) TW[i] = pd.tw3[i];
#endif
  pdl_wait();  // T2 (and possibly V) come from preceding kernels
  double vv[R2];
  static_for<0, R2>([&](auto qq) {
    constexpr int q = decltype(qq)::value;
    vv[q] = (r < R1 && xok) ? __ldg(&vpot[((size_t)(r + R1 * q) * pd.kr2 + y) * pd.kr1 + x]) : 0.0;
  });
  __syncthreads();
  if (tid == 0) {
    for (int s = 0; s < kStages && p0 + s < p1; ++s) {
      mbar_expect_tx(&bar[s], tile_bytes);
      bulk_g2s(ST + (size_t)s * tile_elems, tile + (size_t)(p0 + s) * pstride, tile_bytes, &bar[s]);
    }
  }
  cplx* Sa = S + b;                             // exchange buffer of the inverse transform
  cplx* Sf = S + (XB == 2 ? N * B : 0) + b;     // exchange buffer of the forward transform
  int rA = r;  // role in the band-side radix passes (rotates by one warp per pair)
  for (int pair = p0; pair < p1; ++pair) {
    const int it = pair - p0;
    const int st = it % kStages;
    if (rA < R2) {
      mbar_wait(&bar[st], (unsigned)((it / kStages) & 1));
      const cplx* in = ST + (size_t)st * tile_elems + b;
      cplx v[R1];
      static_for<0, R1>([&](auto kk) {
        constexpr int k = decltype(kk)::value;
        if constexpr (k >= KR::lo && k < KR::hi) {
          const int zr = rA + R2 * k - zlo;
          v[k] = (zr >= 0 && zr < nzb) ? in[zr * B] : mk(0.0, 0.0);
        } else {
          v[k] = mk(0.0, 0.0);
        }
      });
#if CPB_Z_TWREG
      pass_a_raw<R1, R2, true, KR::lo, KR::hi>(v, rA, Sa, B);
#else
      pass_a_in<R1, R2, true, KR::lo, KR::hi, true>(v, rA, TW, Sa, B);
#endif
    }
    __syncthreads();
    if (tid == 0 && pair + kStages < p1) {
      mbar_expect_tx(&bar[st], tile_bytes);
      bulk_g2s(ST + (size_t)st * tile_elems, tile + (size_t)(pair + kStages) * pstride, tile_bytes, &bar[st]);
    }
    cplx u[R2];
    if (r < R1) {
#if CPB_Z_TWREG
      pass_b_tw<R1, R2, true>(u, r, Sa, B, tb);
#else
      pass_b<R1, R2, true>(u, r, Sa, B);
#endif
      static_for<0, R2>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        u[q].x *= vv[q];
        u[q].y *= vv[q];
      });
    }
    if constexpr (XB == 1) __syncthreads();  // everybody has read the single buffer before it is rewritten
#if CPB_Z_TWREG
    if (r < R1) pass_a_fwd_tw<R2, R1>(u, r, Sf, B, tb);
#else
    if (r < R1) pass_a<R2, R1, false, true>(u, r, TW, Sf, B);
#endif
    __syncthreads();
    if (rA < R2) {
      cplx w[R1];
      pass_b<R2, R1, false>(w, rA, Sf, B);
      cplx* d = tile + (size_t)pair * pstride + b;
      static_for<KR::lo, KR::hi>([&](auto kk) {
        constexpr int k = decltype(kk)::value;
        const int zr = rA + R2 * k - zlo;
        if (zr >= 0 && zr < nzb) d[zr * B] = w[k];
      });
    }
    if constexpr (XB == 1) __syncthreads();  // ... and before the next pair's inverse pass writes it
    if constexpr (C::ROT) {
      rA += C::RPW;
      if (rA >= RM) rA -= RM;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// z passes of the DENSE transforms (fftnew with sparse = .FALSE., fftmain_utils.mod.F90:105-120,
// 137-153), used for the density / potential on the density-cutoff sphere (vofrho's local part,
// vofrhoa_utils.mod.F90:88-95, vofrhob_utils.mod.F90:155-173).  A plan built from the nhg list of
// the density cutoff runs them with the same x/y kernels as the wavefunction path; what differs
// is the real-space side: REAL fields (one, or two packed as Re / Im of one complex transform),
// and phasen's factor (-1)^(x+y+z) (fftutil_utils.mod.F90:479-503), which moves G = 0 from the box
// centre used by the index maps to the origin.  One transform per call: no pair loop, plain
// global loads.  grid = (x tiles of the chunk, n2), block = B * max(R1,R2)
// ---------------------------------------------------------------------------------------------
// forward: f(r) -> phasen -> z FFT (e^{-i...}) -> band of T2.  mul != nullptr: the field is scale * mul(r) *
// (fre(r) + i fim(r)) - the pair densities psi_a psi_b / omega and the products v(r) (psi_a + i psi_b) of the
// exact-exchange path (hfx_utils.mod.F90:1052-1063, 1085-1095) without a pass of their own
template <int R1, int R2, int B, bool HALF>
CPB_GLOBAL CPB_LAUNCH_BOUNDS(B* MaxOf<R1, R2>::v, 2)
    k_z_fwd_real(const double* CPB_RESTRICT fre, const double* CPB_RESTRICT fim, cplx* CPB_RESTRICT T2, PlanDev pd,
                 int xt0, const double* CPB_RESTRICT mul, double scale) {
  using KR = KRange<R1, HALF>;
  CPB_DYN_SMEM(cplx, S);
  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x;
  const int b = tid % B, r = tid / B;
  const int xtc = blockIdx.x;
  const int x = (xt0 + xtc) * B + b;
  const int y = blockIdx.y;
  const bool xok = x < pd.n1;
  const int zlo = pd.zlo, nzb = pd.nzb;
  cplx* Sf = S + b;
  if (r < R1) {
    cplx u[R2];
    static_for<0, R2>([&](auto qq) {
      constexpr int q = decltype(qq)::value;
      const int z = r + R1 * q;
      const size_t o = ((size_t)z * pd.kr2 + y) * pd.kr1 + x;
      const double sg = ((x + y + z) & 1) ? -1.0 : 1.0;
      if (mul) {
        const double m = xok ? sg * scale * mul[o] : 0.0;
        u[q] = xok ? mk(m * fre[o], fim ? m * fim[o] : 0.0) : mk(0.0, 0.0);
      } else {
        u[q] = xok ? mk(sg * fre[o], fim ? sg * fim[o] : 0.0) : mk(0.0, 0.0);
      }
    });
    pass_a<R2, R1, false>(u, r, pd.tw3, Sf, B);
  }
  __syncthreads();
  if (r < R2) {
    cplx w[R1];
    pass_b<R2, R1, false>(w, r, Sf, B);
    cplx* d = T2 + ((size_t)xtc * pd.n2 + y) * nzb * B + b;
    static_for<KR::lo, KR::hi>([&](auto kk) {
      constexpr int k = decltype(kk)::value;
      const int zr = r + R2 * k - zlo;
      if (zr >= 0 && zr < nzb) d[zr * B] = w[k];
    });
  }
}

// inverse: band of T2 -> z FFT (e^{+i...}) -> phasen -> Re to ore, Im to oim (if non-null);
// acc != 0: added to the arrays instead of stored
template <int R1, int R2, int B, bool HALF>
CPB_GLOBAL CPB_LAUNCH_BOUNDS(B* MaxOf<R1, R2>::v, 2)
    k_z_inv_real(const cplx* CPB_RESTRICT T2, double* ore, double* oim, PlanDev pd, int xt0, int acc) {
  using KR = KRange<R1, HALF>;
  CPB_DYN_SMEM(cplx, S);
  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x;
  const int b = tid % B, r = tid / B;
  const int xtc = blockIdx.x;
  const int x = (xt0 + xtc) * B + b;
  const int y = blockIdx.y;
  const bool xok = x < pd.n1;
  const int zlo = pd.zlo, nzb = pd.nzb;
  cplx* Sb = S + b;
  if (r < R2) {
    const cplx* in = T2 + ((size_t)xtc * pd.n2 + y) * nzb * B + b;
    cplx v[R1];
    static_for<0, R1>([&](auto kk) {
      constexpr int k = decltype(kk)::value;
      if constexpr (k >= KR::lo && k < KR::hi) {
        const int zr = r + R2 * k - zlo;
        v[k] = (zr >= 0 && zr < nzb) ? in[zr * B] : mk(0.0, 0.0);
      } else {
        v[k] = mk(0.0, 0.0);
      }
    });
    pass_a_in<R1, R2, true, KR::lo, KR::hi>(v, r, pd.tw3, Sb, B);
  }
  __syncthreads();
  if (r < R1 && xok) {
    cplx u[R2];
    pass_b<R1, R2, true>(u, r, Sb, B);
    static_for<0, R2>([&](auto qq) {
      constexpr int q = decltype(qq)::value;
      const int z = r + R1 * q;
      const size_t o = ((size_t)z * pd.kr2 + y) * pd.kr1 + x;
      const double sg = ((x + y + z) & 1) ? -1.0 : 1.0;
      ore[o] = (acc ? ore[o] : 0.0) + sg * u[q].x;
      if (oim) oim[o] = (acc ? oim[o] : 0.0) + sg * u[q].y;
    });
  }
}

}  // namespace cpb
