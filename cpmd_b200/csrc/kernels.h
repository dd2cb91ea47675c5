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
//   T1[pair][ray][x]      after the x pass, only rays inside the cutoff disc   (S_x bytes/pair)
//   T2[pair][zr][y][x]    after the y pass, only z planes inside the band       (S_y bytes/pair)
// The full n^3 complex box never exists in HBM: the z pass consumes it in registers.
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
  int xlo, xhi;       // 0-based x band that holds G-sphere coefficients
  int zlo, nzb;       // 0-based first z plane of the band, number of planes (kr3min..kr3max)
  int nrays;          // internal ray count (>= msrays; dense in y inside every plane)
  int ntiles;         // x-pass tiles (mirror-closed groups of rays)
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
struct PairDev {
  const int* st1;     // state index of the real part (column of c0), always valid
  const int* st2;     // state index of the imaginary part, or -1 (single-state path)
  const double* ca;   // rhoofr: f1/omega ; vpsi: fi   (vpsi_utils.mod.F90:627-633)
  const double* cb;   // rhoofr: f2/omega ; vpsi: fip1
};

// ---------------------------------------------------------------------------------------------
// two-pass tile FFT through shared memory; element idx of batch column b lives at S[idx*LD + b]
// ---------------------------------------------------------------------------------------------

// Thread role a (0 <= a < RB) holds v[k] = x[a + RB*k].  Radix-RA transform, twiddle, store.
template <int RA, int RB, bool INV>
CPB_D void pass_a(cplx (&v)[RA], int a, const cplx* CPB_RESTRICT tw, cplx* Sb, int LD) {
  dft<RA, INV>(v);
  static_for<0, RA>([&](auto pp) {
    constexpr int p = decltype(pp)::value;
    cplx o = v[p];
    if constexpr (p != 0) {
      cplx t = __ldg(&tw[a * p]);
      if constexpr (!INV) t.y = -t.y;
      o = cmul(o, t);
    }
    Sb[(p * RB + a) * LD] = o;
  });
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

template <int A, int B>
struct MaxOf {
  static constexpr int v = A > B ? A : B;
};

// ---------------------------------------------------------------------------------------------
// x pass, inverse: scatter G coefficients of a packed pair into rays + FFT along x.
// One block = one mirror-closed tile of up to SL rays (a ray and its (-y,-z) partner are in the
// same tile, so +G and -G of every plane wave are written by the same block and c0 is read once).
// Fuses zeroing(psi) + set_psi_2_states_g / set_psi_1_state_g (state_utils.mod.F90:132-189) +
// the x mltfft of fftnew (fftmain_utils.mod.F90:93-94).
// grid = (ntiles, npair), block = SL * max(R1,R2), smem = N*(SL+1)*16
// ---------------------------------------------------------------------------------------------
template <int R1, int R2, int SL>
CPB_GLOBAL CPB_LAUNCH_BOUNDS(SL* MaxOf<R1, R2>::v, 1)
    k_x_inv(const cplx* CPB_RESTRICT c0, long ldc, cplx* CPB_RESTRICT T1, PlanDev pd, PairDev pr) {
  constexpr int N = R1 * R2;
  constexpr int LD = SL + 1;
  constexpr int NT = SL * MaxOf<R1, R2>::v;
  CPB_DYN_SMEM(cplx, S);
  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int pair = blockIdx.y;
  const int s1 = pr.st1[pair];
  const int s2 = pr.st2[pair];

  // zero the band rows (only they are read before being overwritten)
  {
    const int lo = pd.xlo * LD, hi = (pd.xhi + 1) * LD;
    for (int i = lo + tid; i < hi; i += NT) S[i] = mk(0.0, 0.0);
  }
  __syncthreads();
  {
    const int e0 = pd.ent_off[tile], e1 = pd.ent_off[tile + 1];
    const cplx* c1p = c0 + (size_t)s1 * ldc;
    const cplx* c2p = c0 + (size_t)(s2 < 0 ? 0 : s2) * ldc;
    for (int e = e0 + tid; e < e1; e += NT) {
      const int ig = pd.ent_ig[e];
      const uint32_t loc = pd.ent_loc[e];
      const int lp = loc & 0xffffu, lm = loc >> 16;
      const cplx a = c1p[ig];
      cplx bq = mk(0.0, 0.0);
      if (s2 >= 0) bq = c2p[ig];
      S[lp] = mk(a.x - bq.y, a.y + bq.x);                // c1 + i c2
      if (lm != lp) S[lm] = mk(a.x + bq.y, bq.x - a.y);  // conj(c1) + i conj(c2)
    }
  }
  __syncthreads();
  const int slot = tid % SL, r = tid / SL;
  cplx v[R1];
  if (r < R2) {
    static_for<0, R1>([&](auto kk) {
      constexpr int k = decltype(kk)::value;
      const int x = r + R2 * k;
      v[k] = (x >= pd.xlo && x <= pd.xhi) ? S[x * LD + slot] : mk(0.0, 0.0);
    });
  }
  __syncthreads();
  if (r < R2) pass_a<R1, R2, true>(v, r, pd.tw1, S + slot, LD);
  __syncthreads();
  cplx u[R2];
  if (r < R1) pass_b<R1, R2, true>(u, r, S + slot, LD);
  __syncthreads();
  if (r < R1) {
    static_for<0, R2>([&](auto qq) {
      constexpr int q = decltype(qq)::value;
      S[(r + R1 * q) * LD + slot] = u[q];
    });
  }
  __syncthreads();
  // coalesced row writes: lanes along x
  for (int i = tid; i < SL * N; i += NT) {
    const int s = i / N, x = i - s * N;
    const int ray = pd.slot_ray[tile * SL + s];
    if (ray >= 0) T1[((size_t)pair * pd.nrays + ray) * N + x] = S[x * LD + s];
  }
}

// ---------------------------------------------------------------------------------------------
// x pass, forward: FFT along x (scale 1/N_total) + unpack of the two states + kinetic term +
// occupation scale + accumulation into c2.  Fuses the last mltfft of fwfftn
// (fftmain_utils.mod.F90:134-136) with vpsi_utils.mod.F90:626-673 and add_wfn (:717).
// ACC: c2 += result (reference semantics) ; !ACC: c2 = result.
// ---------------------------------------------------------------------------------------------
template <int R1, int R2, int SL, bool ACC>
CPB_GLOBAL CPB_LAUNCH_BOUNDS(SL* MaxOf<R1, R2>::v, 1)
    k_x_fwd(const cplx* CPB_RESTRICT T1, const cplx* CPB_RESTRICT c0, cplx* CPB_RESTRICT c2, long ldc,
            PlanDev pd, PairDev pr) {
  constexpr int N = R1 * R2;
  constexpr int LD = SL + 1;
  constexpr int NT = SL * MaxOf<R1, R2>::v;
  CPB_DYN_SMEM(cplx, S);
  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int pair = blockIdx.y;

  for (int i = tid; i < SL * N; i += NT) {
    const int s = i / N, x = i - s * N;
    const int ray = pd.slot_ray[tile * SL + s];
    S[x * LD + s] = (ray >= 0) ? T1[((size_t)pair * pd.nrays + ray) * N + x] : mk(0.0, 0.0);
  }
  __syncthreads();
  const int slot = tid % SL, r = tid / SL;
  // forward transform uses the mirrored factorisation (R2 first, then R1)
  cplx v[R2];
  if (r < R1) {
    static_for<0, R2>([&](auto kk) {
      constexpr int k = decltype(kk)::value;
      v[k] = S[(r + R1 * k) * LD + slot];
    });
  }
  __syncthreads();
  if (r < R1) pass_a<R2, R1, false>(v, r, pd.tw1, S + slot, LD);
  __syncthreads();
  cplx u[R1];
  if (r < R2) pass_b<R2, R1, false>(u, r, S + slot, LD);
  __syncthreads();
  if (r < R2) {
    static_for<0, R1>([&](auto tt) {
      constexpr int t = decltype(tt)::value;
      const int x = r + R2 * t;
      if (x >= pd.xlo && x <= pd.xhi) S[x * LD + slot] = u[t];
    });
  }
  __syncthreads();
  {
    const int s1 = pr.st1[pair];
    const int s2 = pr.st2[pair];
    const double fi = pr.ca[pair], fip1 = pr.cb[pair];
    const int e0 = pd.ent_off[tile], e1 = pd.ent_off[tile + 1];
    const cplx* c1p = c0 + (size_t)s1 * ldc;
    const cplx* c2p = c0 + (size_t)(s2 < 0 ? 0 : s2) * ldc;
    cplx* o1 = c2 + (size_t)s1 * ldc;
    cplx* o2 = c2 + (size_t)(s2 < 0 ? 0 : s2) * ldc;
    const double sc = pd.inv_n;
    for (int e = e0 + tid; e < e1; e += NT) {
      const int ig = pd.ent_ig[e];
      const uint32_t loc = pd.ent_loc[e];
      const cplx psin = S[loc & 0xffffu];
      const cplx psii = S[loc >> 16];
      const cplx fp = mk((psin.x + psii.x) * sc, (psin.y + psii.y) * sc);
      const cplx fm = mk((psin.x - psii.x) * sc, (psin.y - psii.y) * sc);
      const double g2 = pd.tpiba2 * pd.hg[ig];
      const cplx a = c1p[ig];
      cplx r1 = mk(-fi * (g2 * a.x + fp.x), -fi * (g2 * a.y + fm.y));
      if (ACC) r1 = cadd(r1, o1[ig]);
      o1[ig] = r1;
      if (s2 >= 0) {
        const cplx bq = c2p[ig];
        cplx r2 = mk(-fip1 * (g2 * bq.x + fp.y), -fip1 * (g2 * bq.y - fm.x));
        if (ACC) r2 = cadd(r2, o2[ig]);
        o2[ig] = r2;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// y pass, inverse.  Block = (x tile of B columns, z plane of the band, pair).  Reads the rays of
// the plane (zero outside [ylo,yhi]: unpack_x2y's zero fill, fftutil_utils.mod.F90:413-457),
// writes all n2 rows of T2.  grid = (ceil(n1/B), nzb, npair)
// ---------------------------------------------------------------------------------------------
template <int R1, int R2, int B>
CPB_GLOBAL CPB_LAUNCH_BOUNDS(B* MaxOf<R1, R2>::v, 1)
    k_y_inv(const cplx* CPB_RESTRICT T1, cplx* CPB_RESTRICT T2, PlanDev pd) {
  constexpr int N = R1 * R2;
  CPB_DYN_SMEM(cplx, S);
  const int tid = threadIdx.x;
  const int b = tid % B, r = tid / B;
  const int x = blockIdx.x * B + b;
  const int zr = blockIdx.y;
  const int pair = blockIdx.z;
  const bool xok = x < pd.n1;
  const int ylo = pd.ylo[zr], yhi = pd.yhi[zr];
  if (r < R2) {
    cplx v[R1];
    const cplx* src = T1 + ((size_t)pair * pd.nrays + pd.rayoff[zr]) * pd.n1 + x;
    static_for<0, R1>([&](auto kk) {
      constexpr int k = decltype(kk)::value;
      const int y = r + R2 * k;
      v[k] = (xok && y >= ylo && y <= yhi) ? src[(size_t)(y - ylo) * pd.n1] : mk(0.0, 0.0);
    });
    pass_a<R1, R2, true>(v, r, pd.tw2, S + b, B);
  }
  __syncthreads();
  if (r < R1) {
    cplx u[R2];
    pass_b<R1, R2, true>(u, r, S + b, B);
    cplx* dst = T2 + ((size_t)pair * pd.nzb + zr) * N * pd.n1 + x;
    if (xok) {
      static_for<0, R2>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        dst[(size_t)(r + R1 * q) * pd.n1] = u[q];
      });
    }
  }
}

// y pass, forward: reads all n2 rows of T2, writes only the rays of the plane into T1.
template <int R1, int R2, int B>
CPB_GLOBAL CPB_LAUNCH_BOUNDS(B* MaxOf<R1, R2>::v, 1)
    k_y_fwd(const cplx* CPB_RESTRICT T2, cplx* CPB_RESTRICT T1, PlanDev pd) {
  constexpr int N = R1 * R2;
  CPB_DYN_SMEM(cplx, S);
  const int tid = threadIdx.x;
  const int b = tid % B, r = tid / B;
  const int x = blockIdx.x * B + b;
  const int zr = blockIdx.y;
  const int pair = blockIdx.z;
  const bool xok = x < pd.n1;
  const int ylo = pd.ylo[zr], yhi = pd.yhi[zr];
  if (r < R1) {
    cplx v[R2];
    const cplx* src = T2 + ((size_t)pair * pd.nzb + zr) * N * pd.n1 + x;
    static_for<0, R2>([&](auto kk) {
      constexpr int k = decltype(kk)::value;
      v[k] = xok ? src[(size_t)(r + R1 * k) * pd.n1] : mk(0.0, 0.0);
    });
    pass_a<R2, R1, false>(v, r, pd.tw2, S + b, B);
  }
  __syncthreads();
  if (r < R2) {
    cplx u[R1];
    pass_b<R2, R1, false>(u, r, S + b, B);
    cplx* dst = T1 + ((size_t)pair * pd.nrays + pd.rayoff[zr]) * pd.n1 + x;
    static_for<0, R1>([&](auto tt) {
      constexpr int t = decltype(tt)::value;
      const int y = r + R2 * t;
      if (xok && y >= ylo && y <= yhi) dst[(size_t)(y - ylo) * pd.n1] = u[t];
    });
  }
}

// ---------------------------------------------------------------------------------------------
// z pass of rhoofr: z-inverse FFT (band zero-padded to n3: putz, fftutil_utils.mod.F90:87-104)
// fused with build_density_sum (density_utils.mod.F90:61-83).  The block keeps its rho tile in
// registers over all pairs of the batch and does ONE read-modify-write of rho(r) per batch.
// grid = (ceil(n1/B), n2)
// ---------------------------------------------------------------------------------------------
template <int R1, int R2, int B>
CPB_GLOBAL CPB_LAUNCH_BOUNDS(B* MaxOf<R1, R2>::v, 1)
    k_z_rho(const cplx* CPB_RESTRICT T2, double* CPB_RESTRICT rho, PlanDev pd, PairDev pr, int npair) {
  constexpr int N = R1 * R2;
  CPB_DYN_SMEM(cplx, S);
  const int tid = threadIdx.x;
  const int b = tid % B, r = tid / B;
  const int x = blockIdx.x * B + b;
  const int y = blockIdx.y;
  const bool xok = x < pd.n1;
  const size_t zstride = (size_t)pd.n2 * pd.n1;
  double acc[R2];
  static_for<0, R2>([&](auto qq) { acc[decltype(qq)::value] = 0.0; });
  for (int pair = 0; pair < npair; ++pair) {
    if (r < R2) {
      cplx v[R1];
      const cplx* src = T2 + ((size_t)pair * pd.nzb * pd.n2 + y) * pd.n1 + x;
      static_for<0, R1>([&](auto kk) {
        constexpr int k = decltype(kk)::value;
        const int zr = r + R2 * k - pd.zlo;
        v[k] = (xok && zr >= 0 && zr < pd.nzb) ? src[(size_t)zr * zstride] : mk(0.0, 0.0);
      });
      pass_a<R1, R2, true>(v, r, pd.tw3, S + b, B);
    }
    __syncthreads();
    if (r < R1) {
      cplx u[R2];
      pass_b<R1, R2, true>(u, r, S + b, B);
      const double ca = pr.ca[pair], cb = pr.cb[pair];
      static_for<0, R2>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        acc[q] += ca * (u[q].x * u[q].x) + cb * (u[q].y * u[q].y);
      });
    }
    __syncthreads();
  }
  if (r < R1 && xok) {
    static_for<0, R2>([&](auto qq) {
      constexpr int q = decltype(qq)::value;
      const size_t o = ((size_t)(r + R1 * q) * pd.kr2 + y) * pd.kr1 + x;
      rho[o] += acc[q];
    });
  }
}

// ---------------------------------------------------------------------------------------------
// z pass of vpsi: z-inverse FFT, multiply by V(r) (vpsi_utils.mod.F90:487-493), z-forward FFT,
// store only the band back in place (getz, fftutil_utils.mod.F90:106-125).  The real-space
// psi(r) never touches HBM.  The forward transform uses the mirrored factorisation so every
// thread stores exactly the elements it loaded.  V tile lives in registers over the batch.
// grid = (ceil(n1/B), n2)
// ---------------------------------------------------------------------------------------------
template <int R1, int R2, int B>
CPB_GLOBAL CPB_LAUNCH_BOUNDS(B* MaxOf<R1, R2>::v, 1)
    k_z_vpsi(cplx* T2, const double* CPB_RESTRICT vpot, PlanDev pd, int npair) {
  constexpr int N = R1 * R2;
  CPB_DYN_SMEM(cplx, S);
  const int tid = threadIdx.x;
  const int b = tid % B, r = tid / B;
  const int x = blockIdx.x * B + b;
  const int y = blockIdx.y;
  const bool xok = x < pd.n1;
  const size_t zstride = (size_t)pd.n2 * pd.n1;
  double vv[R2];
  static_for<0, R2>([&](auto qq) {
    constexpr int q = decltype(qq)::value;
    vv[q] = (r < R1 && xok) ? __ldg(&vpot[((size_t)(r + R1 * q) * pd.kr2 + y) * pd.kr1 + x]) : 0.0;
  });
  for (int pair = 0; pair < npair; ++pair) {
    cplx* base = T2 + ((size_t)pair * pd.nzb * pd.n2 + y) * pd.n1 + x;
    cplx v[R1];
    if (r < R2) {
      static_for<0, R1>([&](auto kk) {
        constexpr int k = decltype(kk)::value;
        const int zr = r + R2 * k - pd.zlo;
        v[k] = (xok && zr >= 0 && zr < pd.nzb) ? base[(size_t)zr * zstride] : mk(0.0, 0.0);
      });
      pass_a<R1, R2, true>(v, r, pd.tw3, S + b, B);
    }
    __syncthreads();
    cplx u[R2];
    if (r < R1) {
      pass_b<R1, R2, true>(u, r, S + b, B);
      static_for<0, R2>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        u[q].x *= vv[q];
        u[q].y *= vv[q];
      });
    }
    __syncthreads();
    if (r < R1) pass_a<R2, R1, false>(u, r, pd.tw3, S + b, B);
    __syncthreads();
    if (r < R2) {
      pass_b<R2, R1, false>(v, r, S + b, B);
      static_for<0, R1>([&](auto kk) {
        constexpr int k = decltype(kk)::value;
        const int zr = r + R2 * k - pd.zlo;
        if (xok && zr >= 0 && zr < pd.nzb) base[(size_t)zr * zstride] = v[k];
      });
    }
    __syncthreads();
  }
  (void)N;
}

}  // namespace cpb
