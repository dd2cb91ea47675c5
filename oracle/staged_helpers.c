/* Threaded data-movement helpers for oracle/staged_pocketfft.py (TEST INFRASTRUCTURE / CPU baseline
 * only): the copy, scatter and pointwise steps that sit between the library FFT calls of the staged
 * sparse pipeline, i.e. what the reference does in set_psi_2_states_g (state_utils.mod.F90:171-189),
 * unpack_x2y / pack_y2x (fftutil_utils.mod.F90:394-477, 206-290), putz / getz (:87-125),
 * build_density_sum (density_utils.mod.F90:61-83) and the V psi loop (vpsi_utils.mod.F90:487-493),
 * each with an OpenMP loop like the reference's.  All arrays are C-contiguous, pair-major. */
#include <complex.h>
#include <stddef.h>
#include <string.h>

typedef double complex cpx;

/* rays(np, nrays*n1) = 0; rays(izc) = conj(a) + i conj(b); rays(nzc) = a + i b (written last: G = 0) */
void orc_h_set_psi(int np, long ngw, long raylen, const cpx* a, const cpx* b, const long* nzc, const long* izc,
                   cpx* rays) {
#pragma omp parallel for schedule(static)
  for (int p = 0; p < np; ++p) {
    cpx* r = rays + (size_t)p * raylen;
    const cpx* pa = a + (size_t)p * ngw;
    const cpx* pb = b + (size_t)p * ngw;
    memset(r, 0, sizeof(cpx) * (size_t)raylen);
    for (long g = 0; g < ngw; ++g) r[izc[g]] = conj(pa[g]) + I * conj(pb[g]);
    for (long g = 0; g < ngw; ++g) r[nzc[g]] = pa[g] + I * pb[g];
  }
}

/* yf(np, nrows, n1) = 0; yf(:, ms(ray), :) = x(:, ray, :)   (nrows = zband * kr2s) */
void orc_h_unpack_x2y(int np, long nrays, long nrows, long n1, const long* ms, const cpx* x, cpx* yf) {
#pragma omp parallel
  {
#pragma omp for schedule(static)
    for (long i = 0; i < (long)np * nrows; ++i) memset(yf + (size_t)i * n1, 0, sizeof(cpx) * (size_t)n1);
#pragma omp for schedule(static) collapse(2)
    for (int p = 0; p < np; ++p)
      for (long r = 0; r < nrays; ++r)
        memcpy(yf + ((size_t)p * nrows + ms[r]) * n1, x + ((size_t)p * nrays + r) * n1, sizeof(cpx) * (size_t)n1);
  }
}

/* x(:, ray, :) = yf(:, ms(ray), :) */
void orc_h_pack_y2x(int np, long nrays, long nrows, long n1, const long* ms, const cpx* yf, cpx* x) {
#pragma omp parallel for schedule(static) collapse(2)
  for (int p = 0; p < np; ++p)
    for (long r = 0; r < nrays; ++r)
      memcpy(x + ((size_t)p * nrays + r) * n1, yf + ((size_t)p * nrows + ms[r]) * n1, sizeof(cpx) * (size_t)n1);
}

/* full(np, n3, plane) = 0 outside the band, = y(np, nzb, plane) inside [zlo, zlo + nzb) */
void orc_h_putz(int np, long n3, long zlo, long nzb, long plane, const cpx* y, cpx* full) {
#pragma omp parallel for schedule(static) collapse(2)
  for (int p = 0; p < np; ++p)
    for (long z = 0; z < n3; ++z) {
      cpx* d = full + ((size_t)p * n3 + z) * plane;
      if (z >= zlo && z < zlo + nzb)
        memcpy(d, y + ((size_t)p * nzb + (z - zlo)) * plane, sizeof(cpx) * (size_t)plane);
      else
        memset(d, 0, sizeof(cpx) * (size_t)plane);
    }
}

/* rho(n) += sum_p ca(p) Re psi(p, n)^2 + cb(p) Im psi(p, n)^2 */
void orc_h_density_sum(int np, long n, const cpx* psi, const double* ca, const double* cb, double* rho) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) {
    double s = rho[i];
    for (int p = 0; p < np; ++p) {
      const cpx v = psi[(size_t)p * n + i];
      s += ca[p] * (creal(v) * creal(v)) + cb[p] * (cimag(v) * cimag(v));
    }
    rho[i] = s;
  }
}

/* psi(p, n) *= v(n) */
void orc_h_vmul(int np, long n, cpx* psi, const double* v) {
#pragma omp parallel for schedule(static) collapse(2)
  for (int p = 0; p < np; ++p)
    for (long i = 0; i < n; ++i) psi[(size_t)p * n + i] *= v[i];
}

/* pp(p, g) = rays(p, nzc(g)) * scale ; pm(p, g) = rays(p, izc(g)) * scale */
void orc_h_gather_g(int np, long ngw, long raylen, const long* nzc, const long* izc, const cpx* rays, double scale,
                    cpx* pp, cpx* pm) {
#pragma omp parallel for schedule(static) collapse(2)
  for (int p = 0; p < np; ++p)
    for (long g = 0; g < ngw; ++g) {
      pp[(size_t)p * ngw + g] = rays[(size_t)p * raylen + nzc[g]] * scale;
      pm[(size_t)p * ngw + g] = rays[(size_t)p * raylen + izc[g]] * scale;
    }
}
