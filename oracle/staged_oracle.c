/* Staged CPU restatement of CPMD's Gamma-point vpsi + rhoofr (plain C + OpenMP).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/cpmd_oracle.py header): used by tests/ as a second,
 * independent checker and by bench.py's cpu_baseline / --impl reference legs as the timed CPU
 * path ("port": the reference itself is Fortran + FFTW + MPI and cannot be built in this image).
 * PARITY UNPINNED by reference fixtures (none exist); pinned by the known-answer tests and by
 * agreement with the dense-FFT NumPy restatement.
 *
 * Where cpmd_oracle.py uses one dense 3-D FFT, this file follows the reference's *staged sparse*
 * pipeline step by step (paths relative to /root/reference/src):
 *   set_psi_2_states_g / set_psi_1_state_g          state_utils.mod.F90:132-189
 *   fftnew(isign=-1, sparse): x mltfft over msrays rays -> unpack_x2y (zero + scatter through
 *     msp(:,2)) -> y mltfft over zband*kr1 -> putz -> z mltfft      fftmain_utils.mod.F90:92-104,
 *     fftutil_utils.mod.F90:87-104, 394-477
 *   fftnew(isign=+1, sparse): the mirror, 1/(n1 n2 n3) in the last pass   fftmain_utils:122-136
 *   build_density_sum                                density_utils.mod.F90:61-83
 *   V*psi, unpack + kinetic + occupation scale       vpsi_utils.mod.F90:487-493, 626-673
 *   kin_energy, dotp                                 kin_energy_utils:62-110, dotp_utils:26-53
 *   part_1d block partition and pairing              part_1d.mod.F90:22-57, vpsi_utils:376-383
 * The batched 1-D FFT is organised like mltfft_default (mltfft_utils.mod.F90:101-225): OpenMP
 * threads split the m transforms, each thread walks its share in cache-sized batches of `lot`
 * transforms stored z(lot, n) - as separate real and imaginary planes - so the butterflies vectorise
 * across transforms without shuffles (about +35 % per core over interleaved C99 complex); the
 * butterflies themselves are a generic Stockham autosort (radix 4/2/3/5/7), not Goedecker's code.
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cpx;

#define MAXFAC 16
#define LOT 16 /* transforms per cache batch: 2 buffers * 16 * n * 16 B = 96 KB at n=192 (L2 resident) */

typedef struct {
  int n;
  int nfac;
  int fac[MAXFAC];
  cpx* tw; /* exp(sign*2*pi*i*k/n), k<n, built per sign */
  int sign;
} fftplan;

static void plan_init(fftplan* p, int n, int sign) {
  p->n = n;
  p->sign = sign;
  p->nfac = 0;
  int m = n;
  while (m % 4 == 0) { p->fac[p->nfac++] = 4; m /= 4; }
  while (m % 2 == 0) { p->fac[p->nfac++] = 2; m /= 2; }
  while (m % 3 == 0) { p->fac[p->nfac++] = 3; m /= 3; }
  while (m % 5 == 0) { p->fac[p->nfac++] = 5; m /= 5; }
  while (m % 7 == 0) { p->fac[p->nfac++] = 7; m /= 7; }
  if (m != 1) { p->nfac = -1; }
  p->tw = (cpx*)malloc(sizeof(cpx) * (size_t)n);
  for (int k = 0; k < n; ++k) {
    long double a = (long double)sign * 6.283185307179586476925286766559L * k / n;
    p->tw[k] = (double)cosl(a) + I * (double)sinl(a);
  }
}
static void plan_free(fftplan* p) { free(p->tw); }

/* One Stockham stage of radix R on a batch.  The batch is kept as separate real and imaginary planes
 * [n][lot] (the butterflies then vectorise across the lot transforms without shuffles); ns = product of the
 * earlier radices. */
static void stage(const fftplan* p, int R, int ns, const double* restrict ir, const double* restrict ii,
                  double* restrict or_, double* restrict oi, int lot) {
  const int n = p->n, T = n / R;
  const cpx* tw = p->tw;
  const int step = n / (ns * R); /* tw index multiplier: exp(sign 2 pi i r k /(ns R)) = tw[r*k*step] */
  const double sg = (double)p->sign;
  for (int i = 0; i < T; ++i) {
    const int k = i % ns;
    const int j0 = (i / ns) * ns * R + k;
    if (R == 2) {
      const double w1r = creal(tw[(k * step) % n]), w1i = cimag(tw[(k * step) % n]);
      const double *ar = ir + (size_t)i * lot, *ai = ii + (size_t)i * lot;
      const double *br = ir + (size_t)(i + T) * lot, *bi = ii + (size_t)(i + T) * lot;
      double *o0r = or_ + (size_t)j0 * lot, *o0i = oi + (size_t)j0 * lot;
      double *o1r = or_ + (size_t)(j0 + ns) * lot, *o1i = oi + (size_t)(j0 + ns) * lot;
#pragma omp simd
      for (int l = 0; l < lot; ++l) {
        const double v1r = br[l] * w1r - bi[l] * w1i, v1i = br[l] * w1i + bi[l] * w1r;
        o0r[l] = ar[l] + v1r; o0i[l] = ai[l] + v1i;
        o1r[l] = ar[l] - v1r; o1i[l] = ai[l] - v1i;
      }
    } else if (R == 4) {
      const cpx w1 = tw[(k * step) % n], w2 = tw[(2 * k * step) % n], w3 = tw[(3 * k * step) % n];
      const double w1r = creal(w1), w1i = cimag(w1), w2r = creal(w2), w2i = cimag(w2), w3r = creal(w3), w3i = cimag(w3);
      const double *ar = ir + (size_t)i * lot, *ai = ii + (size_t)i * lot;
      const double *br = ir + (size_t)(i + T) * lot, *bi = ii + (size_t)(i + T) * lot;
      const double *cr = ir + (size_t)(i + 2 * T) * lot, *ci = ii + (size_t)(i + 2 * T) * lot;
      const double *dr = ir + (size_t)(i + 3 * T) * lot, *di = ii + (size_t)(i + 3 * T) * lot;
      double *o0r = or_ + (size_t)j0 * lot, *o0i = oi + (size_t)j0 * lot;
      double *o1r = or_ + (size_t)(j0 + ns) * lot, *o1i = oi + (size_t)(j0 + ns) * lot;
      double *o2r = or_ + (size_t)(j0 + 2 * ns) * lot, *o2i = oi + (size_t)(j0 + 2 * ns) * lot;
      double *o3r = or_ + (size_t)(j0 + 3 * ns) * lot, *o3i = oi + (size_t)(j0 + 3 * ns) * lot;
#pragma omp simd
      for (int l = 0; l < lot; ++l) {
        const double v0r = ar[l], v0i = ai[l];
        const double v1r = br[l] * w1r - bi[l] * w1i, v1i = br[l] * w1i + bi[l] * w1r;
        const double v2r = cr[l] * w2r - ci[l] * w2i, v2i = cr[l] * w2i + ci[l] * w2r;
        const double v3r = dr[l] * w3r - di[l] * w3i, v3i = dr[l] * w3i + di[l] * w3r;
        const double t0r = v0r + v2r, t0i = v0i + v2i, t1r = v0r - v2r, t1i = v0i - v2i;
        const double t2r = v1r + v3r, t2i = v1i + v3i;
        const double dqr = v1r - v3r, dqi = v1i - v3i;
        const double t3r = -sg * dqi, t3i = sg * dqr; /* (sign*i) * (v1-v3) */
        o0r[l] = t0r + t2r; o0i[l] = t0i + t2i;
        o1r[l] = t1r + t3r; o1i[l] = t1i + t3i;
        o2r[l] = t0r - t2r; o2i[l] = t0i - t2i;
        o3r[l] = t1r - t3r; o3i[l] = t1i - t3i;
      }
    } else {
      /* generic small radix (3,5,7): O(R^2) with table roots */
      double wr[8], wi[8], rr[8], ri[8];
      for (int r = 0; r < R; ++r) {
        const cpx w = tw[(int)(((long)r * k * step) % n)], q = tw[(r * (n / R)) % n];
        wr[r] = creal(w); wi[r] = cimag(w); rr[r] = creal(q); ri[r] = cimag(q);
      }
      double vr[8][LOT], vi[8][LOT];
      for (int r = 0; r < R; ++r) {
        const double *xr = ir + (size_t)(i + r * T) * lot, *xi = ii + (size_t)(i + r * T) * lot;
#pragma omp simd
        for (int l = 0; l < lot; ++l) {
          vr[r][l] = xr[l] * wr[r] - xi[l] * wi[r];
          vi[r][l] = xr[l] * wi[r] + xi[l] * wr[r];
        }
      }
      for (int q = 0; q < R; ++q) {
        double *yr = or_ + (size_t)(j0 + q * ns) * lot, *yi = oi + (size_t)(j0 + q * ns) * lot;
#pragma omp simd
        for (int l = 0; l < lot; ++l) { yr[l] = vr[0][l]; yi[l] = vi[0][l]; }
        for (int r = 1; r < R; ++r) {
          const double cr_ = rr[(r * q) % R], ci_ = ri[(r * q) % R];
#pragma omp simd
          for (int l = 0; l < lot; ++l) {
            yr[l] += vr[r][l] * cr_ - vi[r][l] * ci_;
            yi[l] += vr[r][l] * ci_ + vi[r][l] * cr_;
          }
        }
      }
    }
  }
}

/* m transforms of length n: element e of transform b is at in[e*ies + b*ibs]; output likewise.
 * Output is multiplied by scale.  (mltfft with 'N','T' / 'T','N' is a choice of strides.) */
static void mltfft(const fftplan* p, const cpx* in, long ies, long ibs, cpx* out, long oes, long obs, long m,
                   double scale) {
  const int n = p->n;
#pragma omp parallel
  {
    double* buf = (double*)aligned_alloc(64, sizeof(double) * 4 * (size_t)n * LOT);
    double *zar = buf, *zai = buf + (size_t)n * LOT, *zbr = buf + 2 * (size_t)n * LOT, *zbi = buf + 3 * (size_t)n * LOT;
#pragma omp for schedule(static)
    for (long b0 = 0; b0 < m; b0 += LOT) {
      const int lot = (int)((m - b0) < LOT ? (m - b0) : LOT);
      if (ies == 1) {
        /* a transform is contiguous in memory: walk it in that order */
        for (int l = 0; l < lot; ++l) {
          const cpx* s = in + (b0 + l) * ibs;
          for (int e = 0; e < n; ++e) { zar[(size_t)e * lot + l] = creal(s[e]); zai[(size_t)e * lot + l] = cimag(s[e]); }
        }
      } else {
        for (int e = 0; e < n; ++e) {
          const cpx* s = in + e * ies + b0 * ibs;
          for (int l = 0; l < lot; ++l) { zar[(size_t)e * lot + l] = creal(s[l * ibs]); zai[(size_t)e * lot + l] = cimag(s[l * ibs]); }
        }
      }
      double *sr = zar, *si = zai, *dr = zbr, *di = zbi;
      int ns = 1;
      for (int s = 0; s < p->nfac; ++s) {
        stage(p, p->fac[s], ns, sr, si, dr, di, lot);
        ns *= p->fac[s];
        double* t = sr; sr = dr; dr = t;
        t = si; si = di; di = t;
      }
      if (oes == 1) {
        for (int l = 0; l < lot; ++l) {
          cpx* d = out + (b0 + l) * obs;
          for (int e = 0; e < n; ++e) d[e] = scale * sr[(size_t)e * lot + l] + I * (scale * si[(size_t)e * lot + l]);
        }
      } else {
        for (int e = 0; e < n; ++e) {
          cpx* d = out + e * oes + b0 * obs;
          for (int l = 0; l < lot; ++l) d[l * obs] = scale * sr[(size_t)e * lot + l] + I * (scale * si[(size_t)e * lot + l]);
        }
      }
    }
    free(buf);
  }
}

typedef struct {
  int n1, n2, n3, kr1, kr2, kr3;
  int ngw, nrays, kr3min, kr3max, zband;
  const int *nzhs, *indzs, *msp2;
  fftplan px[2], py[2], pz[2]; /* [0]: inverse (e^{+i}), [1]: forward (e^{-i}) */
  cpx *psi, *xf, *yf;          /* psi: max(kr1*nrays, nnr1); xf/yf scratch */
} ctx_t;

static size_t maxsz(size_t a, size_t b) { return a > b ? a : b; }

static int ctx_init(ctx_t* c, const int* nr, const int* kr, int ngw, int nrays, const int* nzhs,
                    const int* indzs, const int* msp2, int kr3min, int kr3max) {
  c->n1 = nr[0]; c->n2 = nr[1]; c->n3 = nr[2];
  c->kr1 = kr[0]; c->kr2 = kr[1]; c->kr3 = kr[2];
  c->ngw = ngw; c->nrays = nrays; c->kr3min = kr3min; c->kr3max = kr3max;
  c->zband = kr3max - kr3min + 1;
  c->nzhs = nzhs; c->indzs = indzs; c->msp2 = msp2;
  plan_init(&c->px[0], c->n1, +1); plan_init(&c->px[1], c->n1, -1);
  plan_init(&c->py[0], c->n2, +1); plan_init(&c->py[1], c->n2, -1);
  plan_init(&c->pz[0], c->n3, +1); plan_init(&c->pz[1], c->n3, -1);
  if (c->px[0].nfac < 0 || c->py[0].nfac < 0 || c->pz[0].nfac < 0) return -1;
  size_t nnr1 = (size_t)c->kr1 * c->kr2 * c->kr3;
  size_t rays = (size_t)c->kr1 * nrays;
  size_t planes = (size_t)c->kr2 * c->zband * c->kr1;
  size_t sz = maxsz(maxsz(nnr1, rays), planes);
  c->psi = (cpx*)malloc(sizeof(cpx) * sz);
  c->xf = (cpx*)malloc(sizeof(cpx) * sz);
  c->yf = (cpx*)malloc(sizeof(cpx) * sz);
  return (c->psi && c->xf && c->yf) ? 0 : -2;
}
static void ctx_free(ctx_t* c) {
  for (int s = 0; s < 2; ++s) { plan_free(&c->px[s]); plan_free(&c->py[s]); plan_free(&c->pz[s]); }
  free(c->psi); free(c->xf); free(c->yf);
}

static void zero(cpx* a, size_t n) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)n; ++i) a[i] = 0.0;
}

/* psi ray storage (kr1, nrays) from one or two states; state_utils.mod.F90:132-189 */
static void set_psi(ctx_t* c, const cpx* c1, const cpx* c2, int geq0) {
  zero(c->psi, (size_t)c->kr1 * c->nrays); /* zeroing(psi) rhoofr_utils:328 / vpsi_utils:428 */
  if (c2) {
#pragma omp parallel for schedule(static)
    for (int ig = 0; ig < c->ngw; ++ig) {
      c->psi[c->nzhs[ig] - 1] = c1[ig] + I * c2[ig];
      c->psi[c->indzs[ig] - 1] = conj(c1[ig]) + I * conj(c2[ig]);
    }
    if (geq0) c->psi[c->nzhs[0] - 1] = c1[0] + I * c2[0];
  } else {
#pragma omp parallel for schedule(static)
    for (int ig = 0; ig < c->ngw; ++ig) {
      c->psi[c->nzhs[ig] - 1] = c1[ig];
      c->psi[c->indzs[ig] - 1] = conj(c1[ig]);
    }
    if (geq0) c->psi[c->nzhs[0] - 1] = c1[0];
  }
}

/* fftnew(isign=-1, sparse=.TRUE.): psi rays -> psi(kr1,kr2,kr3) real space */
static void invfft_sparse(ctx_t* c) {
  const int n1 = c->n1, n2 = c->n2, n3 = c->n3, kr1 = c->kr1, kr2 = c->kr2, kr3 = c->kr3;
  const int m = c->nrays, zb = c->zband;
  /* x pass 'N','T': f(kr1, m) -> xf(m, kr1) */
  mltfft(&c->px[0], c->psi, 1, kr1, c->xf, m, 1, m, 1.0);
  /* pack_x2y + fft_comm are copies at one rank; unpack_x2y: zero + scatter (fftutil:413,448-457) */
  const size_t mm = (size_t)kr2 * zb;
  zero(c->yf, mm * kr1);
#pragma omp parallel for schedule(static)
  for (int x = 0; x < n1; ++x)
    for (int r = 0; r < m; ++r) c->yf[(size_t)(c->msp2[r] - 1) + x * mm] = c->xf[r + (size_t)x * m];
  /* y pass: yf(kr2, zb*kr1) -> xf(zb*kr1, kr2) */
  const long my = (long)zb * kr1;
  zero(c->xf, (size_t)my * kr2);
  mltfft(&c->py[0], c->yf, 1, kr2, c->xf, my, 1, (long)zb * n1 /* x < n1 only: pads stay zero */, 1.0);
  /* the batch index is zr + zb*x, so restricting to zb*n1 covers exactly x < n1 */
  /* putz: xf(zb, kr1*kr2) -> yf(kr3, kr1*kr2), zero outside the band (fftutil:87-104) */
  const long mz = (long)kr1 * kr2;
  zero(c->yf, (size_t)kr3 * mz);
#pragma omp parallel for schedule(static)
  for (long t = 0; t < mz; ++t)
    for (int zr = 0; zr < zb; ++zr) c->yf[(size_t)(c->kr3min - 1 + zr) + t * kr3] = c->xf[zr + t * zb];
  /* z pass: yf(kr3, mz) -> psi(mz, kr3) */
  /* (pad columns are all-zero input, so their output is zero; the pad plane z = n3 is zeroed
   * explicitly like mltfft's padding loop, mltfft_utils.mod.F90:227-253) */
  zero(c->psi + (size_t)n3 * mz, (size_t)(kr3 - n3) * mz);
  mltfft(&c->pz[0], c->yf, 1, kr3, c->psi, mz, 1, mz, 1.0);
  (void)n2;
  (void)n3;
}

/* fftnew(isign=+1, sparse=.TRUE.): psi(kr1,kr2,kr3) -> psi rays, scaled by 1/(n1 n2 n3) */
static void fwfft_sparse(ctx_t* c) {
  const int n1 = c->n1, n2 = c->n2, n3 = c->n3, kr1 = c->kr1, kr2 = c->kr2, kr3 = c->kr3;
  const int m = c->nrays, zb = c->zband;
  const long mz = (long)kr1 * kr2;
  /* z pass 'T','N': f(mz, kr3) -> xf(kr3, mz) */
  mltfft(&c->pz[1], c->psi, mz, 1, c->xf, 1, kr3, mz, 1.0);
  /* getz: xf(kr3, mz) -> f(zb, mz) (fftutil:106-125) */
#pragma omp parallel for schedule(static)
  for (long t = 0; t < mz; ++t)
    for (int zr = 0; zr < zb; ++zr) c->psi[zr + t * zb] = c->xf[(size_t)(c->kr3min - 1 + zr) + t * kr3];
  /* y pass 'T','N': f(zb*kr1, kr2) -> yf(kr2, zb*kr1) */
  const long my = (long)zb * kr1;
  mltfft(&c->py[1], c->psi, my, 1, c->yf, 1, kr2, (long)zb * n1, 1.0);
  /* pack_y2x: gather rays through msp (fftutil:206-290) -> xf(m, kr1) */
  const size_t mm = (size_t)kr2 * zb;
#pragma omp parallel for schedule(static)
  for (int x = 0; x < n1; ++x)
    for (int r = 0; r < m; ++r) c->xf[r + (size_t)x * m] = c->yf[(size_t)(c->msp2[r] - 1) + x * mm];
  /* x pass 'T','N' with scale: xf(m, kr1) -> f(kr1, m) */
  const double scale = 1.0 / ((double)n1 * n2 * n3);
  mltfft(&c->px[1], c->xf, m, 1, c->psi, 1, kr1, m, scale);
}

static int nbr_el_in_blk(int n, int proc, int nproc) {
  int res = n % nproc, nbr = (n - res) / nproc;
  return proc < res ? nbr + 1 : nbr;
}
static int get_el_in_blk(int i, int n, int proc, int nproc) {
  int res = n % nproc, nbr = (n - res) / nproc;
  return i + nbr * proc + (proc < res ? proc : res);
}

static double dotp(int n, const cpx* a, const cpx* b, int geq0) {
  double d;
  if (geq0) d = creal(a[0]) * creal(b[0]);
  else d = 2.0 * (creal(a[0]) * creal(b[0]) + cimag(a[0]) * cimag(b[0]));
  double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int i = 1; i < n; ++i) s += creal(a[i]) * creal(b[i]) + cimag(a[i]) * cimag(b[i]);
  return d + 2.0 * s;
}

int orc_set_threads(int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
  return omp_get_max_threads();
#else
  (void)nthreads;
  return 1;
#endif
}

/* rhoofr (rhoofr_utils.mod.F90:122-644), Gamma, no LSD.  c0 is column-major (ld, nstate). */
int orc_rhoofr(const int* nr, const int* kr, int ngw, int nrays, const int* nzhs, const int* indzs,
               const int* msp2, int kr3min, int kr3max, const double* hg, int geq0, double tpiba2,
               double omega, const cpx* c0, long ld, int nstate, const double* f, int ngroups, int my_group,
               double* rhoe, double* ekin, double* rsum_g, double* rsum_r) {
  ctx_t c;
  int rc = ctx_init(&c, nr, kr, ngw, nrays, nzhs, indzs, msp2, kr3min, kr3max);
  if (rc) return rc;
  const size_t nnr1 = (size_t)c.kr1 * c.kr2 * c.kr3;
  /* kin_energy (:178) over all states */
  double rsum = 0.0, xkin = 0.0;
  for (int i = 0; i < nstate; ++i) {
    if (f[i] != 0.0) {
      const cpx* ci = c0 + (size_t)i * ld;
      rsum += f[i] * dotp(ngw, ci, ci, geq0);
      double sk = 0.0;
#pragma omp parallel for reduction(+ : sk) schedule(static)
      for (int ig = 0; ig < ngw; ++ig) sk += hg[ig] * (creal(ci[ig]) * creal(ci[ig]) + cimag(ci[ig]) * cimag(ci[ig]));
      xkin += f[i] * sk;
    }
  }
  memset(rhoe, 0, sizeof(double) * nnr1); /* :198 */
  const int nblk = nbr_el_in_blk(nstate, my_group, ngroups);
  for (int i = 1; i <= nblk; i += 2) { /* :306-310 */
    const int is1 = get_el_in_blk(i, nstate, my_group, ngroups) - 1;
    const int is2 = (i + 1 <= nblk) ? get_el_in_blk(i + 1, nstate, my_group, ngroups) - 1 : -1;
    int tfcal = f[is1] != 0.0;
    if (is2 >= 0) tfcal = tfcal || (f[is2] != 0.0); /* :312-316 */
    if (!tfcal) continue;
    set_psi(&c, c0 + (size_t)is1 * ld, is2 >= 0 ? c0 + (size_t)is2 * ld : NULL, geq0);
    invfft_sparse(&c); /* :346 */
    const double coef3 = f[is1] / omega;
    const double coef4 = is2 >= 0 ? f[is2] / omega : 0.0;
#pragma omp parallel for schedule(static)
    for (long ir = 0; ir < (long)nnr1; ++ir) { /* build_density_sum */
      const double re = creal(c.psi[ir]), im = cimag(c.psi[ir]);
      rhoe[ir] += coef3 * re * re + coef4 * im * im;
    }
  }
  double rsum1 = 0.0;
#pragma omp parallel for reduction(+ : rsum1) schedule(static)
  for (long ir = 0; ir < (long)nnr1; ++ir) rsum1 += rhoe[ir];
  *ekin = xkin * tpiba2;
  *rsum_g = rsum;
  *rsum_r = rsum1 * omega / ((double)c.n1 * c.n2 * c.n3); /* :607-619 */
  ctx_free(&c);
  return 0;
}

/* vpsi (vpsi_utils.mod.F90:120-732), Gamma, RKS, akin = 0.  c2 += C2_vpsi for the group's block. */
int orc_vpsi(const int* nr, const int* kr, int ngw, int nrays, const int* nzhs, const int* indzs,
             const int* msp2, int kr3min, int kr3max, const double* hg, int geq0, double tpiba2,
             const cpx* c0, cpx* c2, long ld, int nstate, const double* f, const double* vpot, int ngroups,
             int my_group, int tksham) {
  ctx_t c;
  int rc = ctx_init(&c, nr, kr, ngw, nrays, nzhs, indzs, msp2, kr3min, kr3max);
  if (rc) return rc;
  const size_t nnr1 = (size_t)c.kr1 * c.kr2 * c.kr3;
  const int nblk = nbr_el_in_blk(nstate, my_group, ngroups);
  for (int i = 1; i <= nblk; i += 2) { /* :376-383 */
    const int is1 = get_el_in_blk(i, nstate, my_group, ngroups) - 1;
    const int is2 = (i + 1 <= nblk) ? get_el_in_blk(i + 1, nstate, my_group, ngroups) - 1 : -1;
    const cpx* c1 = c0 + (size_t)is1 * ld;
    const cpx* cc2 = is2 >= 0 ? c0 + (size_t)is2 * ld : NULL;
    set_psi(&c, c1, cc2, geq0);
    invfft_sparse(&c); /* :443 */
#pragma omp parallel for schedule(static)
    for (long ir = 0; ir < (long)nnr1; ++ir) c.psi[ir] = vpot[ir] * c.psi[ir]; /* :487-493 */
    fwfft_sparse(&c); /* :552 */
    double fi = f[is1] * 0.5; /* :627-633 */
    if (fi == 0.0) fi = tksham ? 0.5 : 1.0;
    double fip1 = 0.0;
    if (is2 >= 0) fip1 = f[is2] * 0.5;
    if (fip1 == 0.0) fip1 = tksham ? 0.5 : 1.0;
    cpx* o1 = c2 + (size_t)is1 * ld;
    cpx* o2 = is2 >= 0 ? c2 + (size_t)is2 * ld : NULL;
#pragma omp parallel for schedule(static)
    for (int ig = 0; ig < ngw; ++ig) { /* :655-671 */
      const cpx psin = c.psi[nzhs[ig] - 1];
      const cpx psii = c.psi[indzs[ig] - 1];
      const cpx fp = psin + psii, fm = psin - psii;
      const double g2 = tpiba2 * hg[ig];
      o1[ig] += -fi * (g2 * c1[ig] + (creal(fp) + I * cimag(fm)));
      if (o2) o2[ig] += -fip1 * (g2 * cc2[ig] + (cimag(fp) - I * creal(fm)));
    }
  }
  ctx_free(&c);
  return 0;
}
