// Mesh-length independent kernels: G-space pack / unpack and the reductions (included by
// cpb200.cu only).
#pragma once
#include "kernels.h"

namespace cpb {

// ---------------------------------------------------------------------------------------------
// G-space reductions of kin_energy (kin_energy_utils.mod.F90:62-110) and dotp
// (dotp_utils.mod.F90:26-53).  grid = (kKinChunks, states): block (c, st) reduces its contiguous
// slice of the plane waves of one state in a fixed order (bit-stable); the host adds the
// kKinChunks partials of a state in order.
// out[(st*kKinChunks + c)*2] = sum_G hg |c|^2 ; out[.. + 1] = dotp(c,c).   block = 256
// ---------------------------------------------------------------------------------------------
constexpr int kKinChunks = 8;

CPB_GLOBAL k_kin_energy(const cplx* CPB_RESTRICT c0, long ldc, int first_state, int ngw, int geq0,
                        const double* CPB_RESTRICT hg, double* CPB_RESTRICT out) {
  CPB_DYN_SMEM(double, red);  // 2*256
  const int tid = threadIdx.x;
  const int st = blockIdx.y;
  const int per = (ngw + kKinChunks - 1) / kKinChunks;
  const int g0 = blockIdx.x * per;
  const int g1 = (g0 + per < ngw) ? g0 + per : ngw;
  const cplx* c = c0 + (size_t)(first_state + st) * ldc;
  double sk = 0.0, sd = 0.0;
  for (int ig = g0 + tid; ig < g1; ig += 256) {
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
    out[((size_t)st * kKinChunks + blockIdx.x) * 2] = red[0];
    out[((size_t)st * kKinChunks + blockIdx.x) * 2 + 1] = red[256];
  }
}

// Last step of the kin_energy / dotp sums that k_x_inv_m accumulates while it gathers the coefficients:
// block p adds the nbx block partials part[(p*nbx + bx)*4 + q] of pair p in a fixed order (bit-stable)
// and writes them where the host expects k_kin_energy's chunk 0 of the pair's states:
// out[((st - first)*kKinChunks + 0)*2 + {0: sum hg |c|^2, 1: dotp}], q = 0,1 -> st1, q = 2,3 -> st2.
// The other chunks of `out` were zeroed by the caller.  grid = pairs, block = 128
CPB_GLOBAL k_kin_reduce(const double* CPB_RESTRICT part, int nbx, PairDev pr, int first_state,
                        double* CPB_RESTRICT out) {
  CPB_DYN_SMEM(double, red);  // 4*128
  const int tid = threadIdx.x;
  const int pair = blockIdx.x;
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  for (int bx = tid; bx < nbx; bx += 128)
    for (int q = 0; q < 4; ++q) s[q] += part[((size_t)pair * nbx + bx) * 4 + q];
  for (int q = 0; q < 4; ++q) red[q * 128 + tid] = s[q];
  __syncthreads();
  for (int k = 64; k > 0; k >>= 1) {
    if (tid < k)
      for (int q = 0; q < 4; ++q) red[q * 128 + tid] += red[q * 128 + tid + k];
    __syncthreads();
  }
  if (tid < 4) {
    const int st = tid < 2 ? pr.st1[pair] : pr.st2[pair];
    if (st >= 0) out[((size_t)(st - first_state) * kKinChunks) * 2 + (tid & 1)] = red[tid * 128];
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

// LSD post-processing (rhoofr_utils.mod.F90:543-559): per-block partial sums of alpha, beta and
// |alpha - beta| (fixed order); finalize != 0: column 1 becomes alpha + beta.  block = 256,
// partial[0..g) alpha, [g..2g) beta, [2g..3g) |alpha - beta| with g = gridDim.x
CPB_GLOBAL k_lsd_sums(double* a, const double* CPB_RESTRICT b, size_t n, double* CPB_RESTRICT partial, int finalize) {
  CPB_DYN_SMEM(double, red);  // 3*256
  const int tid = threadIdx.x;
  double sa = 0.0, sb = 0.0, sd = 0.0;
  for (size_t i = (size_t)blockIdx.x * 256 + tid; i < n; i += (size_t)gridDim.x * 256) {
    const double x = a[i], y = b[i];
    sa += x;
    sb += y;
    sd += (x - y < 0.0) ? (y - x) : (x - y);
    if (finalize) a[i] = x + y;
  }
  red[tid] = sa;
  red[256 + tid] = sb;
  red[512 + tid] = sd;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (tid < k) {
      red[tid] += red[tid + k];
      red[256 + tid] += red[256 + tid + k];
      red[512 + tid] += red[512 + tid + k];
    }
    __syncthreads();
  }
  if (tid == 0) {
    partial[blockIdx.x] = red[0];
    partial[gridDim.x + blockIdx.x] = red[256];
    partial[2 * gridDim.x + blockIdx.x] = red[512];
  }
}

// ---------------------------------------------------------------------------------------------
// k_unpack: vpsi_utils.mod.F90:626-673 + add_wfn (:717).  Reads FFT[V psi] at +G and -G from the
// band-ray storage G (kernels.h; already scaled by 1/N in k_x_fwd), separates the two states, adds the
// kinetic term, scales by -f/2 and updates c2.  ACC: c2 += result (reference semantics);
// !ACC: c2 = result.  grid = (ceil(ngw/256), pair groups), block = 256
// ---------------------------------------------------------------------------------------------
template <bool ACC>
CPB_GLOBAL k_unpack(const cplx* CPB_RESTRICT G, const cplx* CPB_RESTRICT c0, cplx* CPB_RESTRICT c2, long ldc,
                    PlanDev pd, PairDev pr, int npair, int ppg) {
  const int ig = blockIdx.x * 256 + threadIdx.x;
  if (ig >= pd.ngw) return;
  const int p0 = blockIdx.y * ppg;
  const int p1 = (p0 + ppg < npair) ? p0 + ppg : npair;
  const uint32_t lp = pd.gpos[ig], lm = pd.gneg[ig];
  const double g2 = pd.tpiba2 * pd.hg[ig];
  const size_t g_pair = (size_t)pd.nxb * pd.nrp;
  for (int pair = p0; pair < p1; ++pair) {
    const int s1 = pr.st1[pair], s2 = pr.st2[pair];
    const double fi = pr.ca[pair], fip1 = pr.cb[pair];
    const cplx* g = G + (size_t)pair * g_pair;
    const cplx psin = g[lp];
    const cplx psii = g[lm];
    cplx* o1 = c2 + (size_t)s1 * ldc + ig;
    cplx* o2 = c2 + (size_t)(s2 < 0 ? s1 : s2) * ldc + ig;
    const cplx a = c0[(size_t)s1 * ldc + ig];
    const cplx bq = (s2 >= 0) ? c0[(size_t)s2 * ldc + ig] : mk(0.0, 0.0);
    cplx old1 = mk(0.0, 0.0), old2 = mk(0.0, 0.0);
    if (ACC) {
      old1 = *o1;
      if (s2 >= 0) old2 = *o2;
    }
    const cplx fp = cadd(psin, psii);
    const cplx fm = csub(psin, psii);
    cplx r1 = mk(-fi * (g2 * a.x + fp.x), -fi * (g2 * a.y + fm.y));
    if (ACC) r1 = cadd(r1, old1);
    *o1 = r1;
    if (s2 >= 0) {
      cplx r2 = mk(-fip1 * (g2 * bq.x + fp.y), -fip1 * (g2 * bq.y - fm.x));
      if (ACC) r2 = cadd(r2, old2);
      *o2 = r2;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Dense transforms, G side.  k_gather_g: the counterpart of zgthr(nhg, v, vtemp, nzh)
// (vofrhob_utils.mod.F90:240,245; rhog(ig) = v(nzh(ig)) in ppener_utils.mod.F90:91) on the band-ray
// storage written by k_x_fwd.  One real field: a(G) = P(+G), exactly what the reference reads.  Two
// real fields packed as Re / Im of one transform (p = a + i b): a(G) = [P(G) + conj P(-G)] / 2,
// b(G) = [P(G) - conj P(-G)] / (2i).   grid = ceil(ngw/256), block = 256
// ---------------------------------------------------------------------------------------------
CPB_GLOBAL k_gather_g(const cplx* CPB_RESTRICT G, PlanDev pd, cplx* CPB_RESTRICT ga, cplx* CPB_RESTRICT gb) {
  const int ig = blockIdx.x * 256 + threadIdx.x;
  if (ig >= pd.ngw) return;
  const cplx pp = G[pd.gpos[ig]];
  if (!gb) {
    ga[ig] = pp;
    return;
  }
  const cplx pm = G[pd.gneg[ig]];
  ga[ig] = mk(0.5 * (pp.x + pm.x), 0.5 * (pp.y - pm.y));
  gb[ig] = mk(0.5 * (pp.y + pm.y), 0.5 * (pm.x - pp.x));
}

// ---------------------------------------------------------------------------------------------
// k_ppener: ppener (ppener_utils.mod.F90:23-108).  vtemp(ig) = scg(ig) * (rhog(ig) + eirop(ig)) +
// eivps(ig) and the four complex sums eh, ei, ee, eps; the G = 0 entry (geq0, ig = 0) follows the
// reference's special case (:58-70): half weights, vtemp(1) = scg(1) * rhog without the
// pseudopotential term.  Per-block partials in a fixed order (bit-stable):
// partial[(j*gridDim.x + block)] for j = 0..7 = Re eh, Im eh, Re ei, Im ei, Re ee, Im ee, Re eps,
// Im eps.  block = 256
// ---------------------------------------------------------------------------------------------
CPB_GLOBAL k_ppener(const cplx* CPB_RESTRICT rhog, const double* CPB_RESTRICT scg, const cplx* CPB_RESTRICT eivps,
                    const cplx* CPB_RESTRICT eirop, cplx* CPB_RESTRICT vtemp, int nhg, int geq0,
                    double* CPB_RESTRICT partial) {
  CPB_DYN_SMEM(double, red);  // 8*256
  const int tid = threadIdx.x;
  double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int ig = blockIdx.x * 256 + tid; ig < nhg; ig += gridDim.x * 256) {
    const cplx vp = eivps[ig], rp = eirop[ig], rhet = rhog[ig];
    const cplx rg = cadd(rhet, rp);
    const double sc = scg[ig];
    if (ig == 0 && geq0) {
      const cplx e = cscale(cmulc(vp, rhet), 0.5);  // 0.5 * vp * conj(rhet)
      s[6] += e.x;
      s[7] += e.y;
      s[0] += 0.5 * sc * rg.x * rg.x;               // 0.5 * scg * Re(rhog)^2
      const cplx i2 = cscale(cmul(rp, rp), 0.5 * sc);
      s[2] += i2.x;
      s[3] += i2.y;
      const cplx e2 = cscale(cmul(rhet, rhet), 0.5 * sc);
      s[4] += e2.x;
      s[5] += e2.y;
      vtemp[ig] = cscale(rg, sc);
    } else {
      const cplx vcg = cscale(rg, sc);
      vtemp[ig] = cadd(vcg, vp);
      const cplx h = cmulc(vcg, rg);  // vcg * conj(rhog)
      s[0] += h.x;
      s[1] += h.y;
      const cplx i2 = cscale(cmulc(rp, rp), sc);
      s[2] += i2.x;
      s[3] += i2.y;
      const cplx e2 = cscale(cmulc(rhet, rhet), sc);
      s[4] += e2.x;
      s[5] += e2.y;
      const cplx e = cmulc(vp, rhet);  // conj(rhet) * vp
      s[6] += e.x;
      s[7] += e.y;
    }
  }
  for (int j = 0; j < 8; ++j) red[j * 256 + tid] = s[j];
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (tid < k)
      for (int j = 0; j < 8; ++j) red[j * 256 + tid] += red[j * 256 + tid + k];
    __syncthreads();
  }
  if (tid < 8) partial[(size_t)tid * gridDim.x + blockIdx.x] = red[tid * 256];
}

// ---------------------------------------------------------------------------------------------
// k-points (tkpts%tkpnt): one complex state per transform, c0(2 ngw, nstate) = +G components, then
// -G components.  The gather of k_x_inv runs unchanged on a second position table (gtab_k: the -G
// positions hold ig + ngw without the conjugation flag = set_psi_1_state_g_kpts,
// state_utils.mod.F90:192-224); these are the two G-space kernels that differ.
// ---------------------------------------------------------------------------------------------
// kinetic energy and norm of rhoofr_c (rhoofr_c_utils.mod.F90:117-140): out[(st*kKinChunks+c)*2] =
// sum_G hgkp |c(G)|^2 + hgkm |c(G+ngw)|^2, out[.. + 1] = sum over all 2 ngw components |c|^2
CPB_GLOBAL k_kin_energy_kpt(const cplx* CPB_RESTRICT c0, long ldc, int first_state, int ngw,
                            const double* CPB_RESTRICT hgkp, const double* CPB_RESTRICT hgkm,
                            double* CPB_RESTRICT out) {
  CPB_DYN_SMEM(double, red);  // 2*256
  const int tid = threadIdx.x;
  const int st = blockIdx.y;
  const int per = (ngw + kKinChunks - 1) / kKinChunks;
  const int g0 = blockIdx.x * per;
  const int g1 = (g0 + per < ngw) ? g0 + per : ngw;
  const cplx* c = c0 + (size_t)(first_state + st) * ldc;
  double sk = 0.0, sd = 0.0;
  for (int ig = g0 + tid; ig < g1; ig += 256) {
    const cplx a = c[ig], b = c[ig + ngw];
    const double ma = a.x * a.x + a.y * a.y, mb = b.x * b.x + b.y * b.y;
    sk += hgkp[ig] * ma + hgkm[ig] * mb;
    sd += ma + mb;
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
    out[((size_t)st * kKinChunks + blockIdx.x) * 2] = red[0];
    out[((size_t)st * kKinChunks + blockIdx.x) * 2 + 1] = red[256];
  }
}

// vpsi_utils.mod.F90:614-625 + add_wfn (:717): C2(ig) = -fi (tpiba2/2 hgkp c0(ig) + psi(nzhs(ig))),
// C2(ig+ngw) = -fi (tpiba2/2 hgkm c0(ig+ngw) + psi(indzs(ig))), C2(1+ngw) = 0 if geq0.
// pr.st1 = state, pr.ca = fi.  grid = (ceil(ngw/256), pair groups), block = 256
template <bool ACC>
CPB_GLOBAL k_unpack_kpt(const cplx* CPB_RESTRICT G, const cplx* CPB_RESTRICT c0, cplx* CPB_RESTRICT c2, long ldc,
                        PlanDev pd, PairDev pr, const double* CPB_RESTRICT hgkp, const double* CPB_RESTRICT hgkm,
                        int geq0, int npair, int ppg) {
  const int ig = blockIdx.x * 256 + threadIdx.x;
  if (ig >= pd.ngw) return;
  const int p0 = blockIdx.y * ppg;
  const int p1 = (p0 + ppg < npair) ? p0 + ppg : npair;
  const uint32_t lp = pd.gpos[ig], lm = pd.gneg[ig];
  const double kp = 0.5 * pd.tpiba2 * hgkp[ig], km = 0.5 * pd.tpiba2 * hgkm[ig];
  const size_t g_pair = (size_t)pd.nxb * pd.nrp;
  const bool zero_m = geq0 && ig == 0;
  for (int pair = p0; pair < p1; ++pair) {
    const int s1 = pr.st1[pair];
    const double fi = pr.ca[pair];
    const cplx* g = G + (size_t)pair * g_pair;
    const cplx fp = g[lp], fm = g[lm];
    const cplx a = c0[(size_t)s1 * ldc + ig], b = c0[(size_t)s1 * ldc + ig + pd.ngw];
    cplx* o1 = c2 + (size_t)s1 * ldc + ig;
    cplx* o2 = o1 + pd.ngw;
    cplx r1 = mk(-fi * (kp * a.x + fp.x), -fi * (kp * a.y + fp.y));
    cplx r2 = zero_m ? mk(0.0, 0.0) : mk(-fi * (km * b.x + fm.x), -fi * (km * b.y + fm.y));
    if (ACC) {
      r1 = cadd(r1, *o1);
      r2 = cadd(r2, *o2);
    }
    *o1 = r1;
    *o2 = r2;
  }
}

// ---------------------------------------------------------------------------------------------
// meta-GGA: ftauadd (vtaupsi_utils.mod.F90:131-165).  psi = FFT[vtau * d_k psi] at +-G from the band-ray
// storage; c2(ig,is1) -= fi1 gk (Re fm, Im fp), c2(ig,is2) -= fi2 gk (Im fm, -Re fp), fp/fm = psi(+G) +- psi(-G),
// gk = gk[3*ig] (caller passes gk + direction).  pr.ca/cb = fi1/fi2 = f tpiba2 / 4.
// grid = (ceil(ngw/256), pair groups), block = 256
// ---------------------------------------------------------------------------------------------
CPB_GLOBAL k_unpack_tau(const cplx* CPB_RESTRICT G, cplx* CPB_RESTRICT c2, long ldc, PlanDev pd, PairDev pr,
                        const double* CPB_RESTRICT gk, int npair, int ppg) {
  const int ig = blockIdx.x * 256 + threadIdx.x;
  if (ig >= pd.ngw) return;
  const int p0 = blockIdx.y * ppg;
  const int p1 = (p0 + ppg < npair) ? p0 + ppg : npair;
  const uint32_t lp = pd.gpos[ig], lm = pd.gneg[ig];
  const double g = gk[3 * (size_t)ig];
  const size_t g_pair = (size_t)pd.nxb * pd.nrp;
  for (int pair = p0; pair < p1; ++pair) {
    const int s1 = pr.st1[pair], s2 = pr.st2[pair];
    const double f1 = pr.ca[pair] * g, f2 = pr.cb[pair] * g;
    const cplx* gp = G + (size_t)pair * g_pair;
    const cplx psin = gp[lp], psii = gp[lm];
    const cplx fp = cadd(psin, psii), fm = csub(psin, psii);
    cplx* o1 = c2 + (size_t)s1 * ldc + ig;
    const cplx a = *o1;
    *o1 = mk(a.x - f1 * fm.x, a.y - f1 * fp.y);
    if (s2 >= 0) {
      cplx* o2 = c2 + (size_t)s2 * ldc + ig;
      const cplx b = *o2;
      *o2 = mk(b.x - f2 * fm.y, b.y + f2 * fp.x);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Hartree-Fock exchange (hfx_utils.mod.F90:1034-1260), G side on the pair-density FFT set.
// k_hfx_coulomb: the pair density (densities) from the band-ray storage of the dense forward transform -
// one real field: rho(G) = P(+G); two packed as Re / Im: the +-G separation of k_gather_g - then
// vpotg = -pf scgx rho (:1068) and the pair energy sum_G 4 Re(vpotg conj rho), G = 0 counted half (:1069-1071).
// eacc[block] += the block's partial sum of both fields: a fixed launch order makes the total bit-stable.
// block = 256
// ---------------------------------------------------------------------------------------------
CPB_GLOBAL k_hfx_coulomb(const cplx* CPB_RESTRICT G, PlanDev pd, const double* CPB_RESTRICT scgx, double pf1, double pf2,
                         int two, int geq0, cplx* CPB_RESTRICT v1, cplx* CPB_RESTRICT v2, double* eacc) {
  CPB_DYN_SMEM(double, red);  // 256
  const int tid = threadIdx.x;
  double e = 0.0;
  for (int ig = blockIdx.x * 256 + tid; ig < pd.ngw; ig += gridDim.x * 256) {
    const cplx pp = G[pd.gpos[ig]];
    const double sc = scgx[ig];
    const double w = (ig == 0 && geq0) ? 2.0 : 4.0;
    if (!two) {
      const cplx vg = cscale(pp, -pf1 * sc);
      v1[ig] = vg;
      e += w * (vg.x * pp.x + vg.y * pp.y);
    } else {
      const cplx pm = G[pd.gneg[ig]];
      const cplx ra = mk(0.5 * (pp.x + pm.x), 0.5 * (pp.y - pm.y));
      const cplx rb = mk(0.5 * (pp.y + pm.y), 0.5 * (pm.x - pp.x));
      const cplx va = cscale(ra, -pf1 * sc), vb = cscale(rb, -pf2 * sc);
      v1[ig] = va;
      v2[ig] = vb;
      e += w * (va.x * ra.x + va.y * ra.y) + w * (vb.x * rb.x + vb.y * rb.y);
    }
  }
  red[tid] = e;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (tid < k) red[tid] += red[tid + k];
    __syncthreads();
  }
  if (tid == 0) eacc[blockIdx.x] += red[0];
}

// k_hfx_acc: the decode of hfxab (hfx_utils.mod.F90:1097-1107) from the band-ray storage of the sparse forward
// transform of v(r) (psi_a + i psi_b): c2b -= (Re fp, Im fm), c2a -= (Im fp, -Re fm), fp / fm = P(+G) +- P(-G).
// same != 0 (the diagonal term, a == b): both updates go to the one column.  grid = ceil(ngw/256)
CPB_GLOBAL k_hfx_acc(const cplx* CPB_RESTRICT G, PlanDev pd, cplx* c2a, cplx* c2b, int same) {
  const int ig = blockIdx.x * 256 + threadIdx.x;
  if (ig >= pd.ngw) return;
  const cplx pp = G[pd.gpos[ig]], pm = G[pd.gneg[ig]];
  const cplx fp = cadd(pp, pm), fm = csub(pp, pm);
  if (same) {
    const cplx a = c2a[ig];
    c2a[ig] = mk(a.x - fp.x - fp.y, a.y - fm.y + fm.x);
  } else {
    const cplx b = c2b[ig], a = c2a[ig];
    c2b[ig] = mk(b.x - fp.x, b.y - fm.y);
    c2a[ig] = mk(a.x - fp.y, a.y + fm.x);
  }
}

// dotp(c0_i, c2_i) per state (dotp_utils.mod.F90:26-53: weight 2, the real parts once at G = 0): per-chunk partials
// out[st*kKinChunks + c], fixed order.  grid = (kKinChunks, states), block = 256
CPB_GLOBAL k_dotp(const cplx* CPB_RESTRICT a, const cplx* CPB_RESTRICT b, long ld, int ngw, int geq0,
                  double* CPB_RESTRICT out) {
  CPB_DYN_SMEM(double, red);  // 256
  const int tid = threadIdx.x;
  const int st = blockIdx.y;
  const int per = (ngw + kKinChunks - 1) / kKinChunks;
  const int g0 = blockIdx.x * per;
  const int g1 = (g0 + per < ngw) ? g0 + per : ngw;
  const cplx* pa = a + (size_t)st * ld;
  const cplx* pb = b + (size_t)st * ld;
  double s = 0.0;
  for (int ig = g0 + tid; ig < g1; ig += 256) {
    const cplx x = pa[ig], y = pb[ig];
    s += (ig == 0 && geq0) ? x.x * y.x : 2.0 * (x.x * y.x + x.y * y.y);
  }
  red[tid] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (tid < k) red[tid] += red[tid + k];
    __syncthreads();
  }
  if (tid == 0) out[(size_t)st * kKinChunks + blockIdx.x] = red[0];
}

}  // namespace cpb
