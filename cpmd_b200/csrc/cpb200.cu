// cpb200: plan construction (host side of fftprp/loadpa-derived index maps) and the C ABI.
// See include/cpb200.h for the contract and the reference file:line each entry point replaces.
#include "../../include/cpb200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "axis.h"
#include "misc_kernels.h"
#include "rt.h"

using namespace cpb;

#if !defined(CPB_EMULATE)
namespace cpb {
constexpr unsigned kPdlDefaultMask = 0x2fu;  // not y_fwd: starting it under the tail of k_z_vpsi costs 0.09 ms per launch (profiles/r02p_probe_pdl_mask.txt)
// programmatic dependent launch of the hot-path kernels (cpb_defs.h): bit c of CPB_PDL allows kernels of class c
// (0 x_inv, 1 y_inv, 2 z_rho, 3 z_vpsi, 4 y_fwd, 5 x_fwd, 6 dense z) to become resident before their predecessor
// on the stream has finished
unsigned pdl_mask() {
  static const unsigned mask = [] {
    const char* e = std::getenv("CPB_PDL");
    return e ? (unsigned)std::strtoul(e, nullptr, 0) : kPdlDefaultMask;
  }();
  return mask;
}
}  // namespace cpb
#endif

namespace {
thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

template <class T>
T* upload(const std::vector<T>& h) {
  T* d = (T*)rt::dmalloc(h.size() * sizeof(T));
  if (!h.empty()) {
    rt::h2d(d, h.data(), h.size() * sizeof(T), 0);
    rt::sync(0);
  }
  return d;
}

struct PairHost {
  int s1, s2;    // 0-based global state indices, s2 = -1 for a single-state transform
  int chan = 0;  // LSD: spin channel (0 alpha, 1 beta) of the pair's real-space arrays
};

// part_1d.mod.F90:22-57
int nbr_el_in_blk(int n_elem, int proc, int nproc) {
  const int res = n_elem % nproc;
  int nbr = (n_elem - res) / nproc;
  if (proc < res) nbr += 1;
  return nbr;
}
int get_el_in_blk(int i_elem, int n_elem, int proc, int nproc) {
  const int res = n_elem % nproc;
  const int nbr = (n_elem - res) / nproc;
  return i_elem + nbr * proc + std::min(proc, res);
}

// pairs formed inside the group's block: vpsi_utils.mod.F90:376-383, rhoofr_utils.mod.F90:306-310.
// nsup >= 0 selects LSD (cntl%tlsd): states [0, nsup) are alpha, the rest beta (spin_mod%nsup).
// The reference sends the Re part of a pair to the channel of is1 and the Im part to that of is2
// (rhoofr_utils.mod.F90:375-385, vpsi_utils.mod.F90:450-482).  Here every launch works on one
// channel, so the one pair that straddles the spin boundary (is1 == nsup, 1-based) is transformed
// as two single states - the same linear map, the packing of two states into one transform being
// an optimisation, not part of the result - and all alpha pairs precede all beta pairs.
std::vector<PairHost> block_pairs(int nstate, int my_group, int ngroups, int nsup = -1) {
  std::vector<PairHost> out;
  const int nblk = nbr_el_in_blk(nstate, my_group, ngroups);
  for (int i = 1; i <= nblk; i += 2) {
    PairHost p;
    p.s1 = get_el_in_blk(i, nstate, my_group, ngroups) - 1;
    p.s2 = (i + 1 <= nblk) ? get_el_in_blk(i + 1, nstate, my_group, ngroups) - 1 : -1;
    p.chan = (nsup >= 0 && p.s1 >= nsup) ? 1 : 0;
    if (nsup >= 0 && p.s2 >= 0 && p.s1 < nsup && p.s2 >= nsup) {
      PairHost q;
      q.s1 = p.s2;
      q.s2 = -1;
      q.chan = 1;
      p.s2 = -1;
      out.push_back(p);
      out.push_back(q);
    } else {
      out.push_back(p);
    }
  }
  return out;
}
// number of leading pairs that belong to channel 0
int count_chan0(const std::vector<PairHost>& pairs) {
  int n = 0;
  while (n < (int)pairs.size() && pairs[n].chan == 0) ++n;
  return n;
}
}  // namespace

struct cpb_plan {
  int nr[3], kr[3];
  int ngw = 0;
  int geq0 = 0;
  double tpiba2 = 0, omega = 0;
  int device = 0;
  int max_batch = 32;
  int host_batch = 8;      // pairs per batch of the host-pointer entry points (finer copy / compute pipeline)
  int batch_override = 0;  // > 0: batch size of the call in progress
  bool auto_batch = false;  // max_batch_pairs <= 0: sized from the work space of one pair once the ray table is known
  int nxt = 0;       // x tiles of B columns
  int chunk_xt = 1;  // x tiles per y/z chunk (T2 holds one chunk of the batch)
  int n_sm = 148;
  // CPB_PSI_KEEP / CPB_PSI_REUSE: y-pass output of every pair of the last rhoofr call
  cplx* T2keep = nullptr;
  size_t t2keep_pairs = 0;   // capacity
  size_t t2_pair = 0;        // elements of T2 per pair (all x tiles)
  bool psi_valid = false;
  int psi_npairs = 0;
  const void* psi_key_ptr = nullptr;
  long psi_key[5] = {0, 0, 0, 0, 0};  // ld, nstate, ngroups, my_group, nsup
  double prologue_pairs = 0.25;  // block prologue cost in pair-times (pairs_per_group model)
  double prologue_pairs_x = 1.0; // same for the mirror-pair forward x kernel (position tables + hg per block; CPB_PROLOGUE_X)
  double prologue_pairs_xinv = 1.0;  // ... and for the inverse one (CPB_PROLOGUE_XINV; CPB_PROLOGUE_X sets both)
  int x_sub = 32;         // pairs per forward x-pass sub-batch (measured: fewer launches beat L2 residency of G, profiles/r01g_notes.txt)
  size_t t1_pair = 0;     // elements of T1 per pair
  size_t g_pair = 0;      // elements of the band-ray storage per pair (nxb * nrp)
  bool half_x = false, half_y = false, half_z = false;  // band-pruned kernel variants usable
  bool xw_inv = false, xw_fwd = false;  // x passes run the warp-autonomous mirror-pair kernels (kernels_xw.h)
  int nux_w = 0;                        // their units (warps) per pair group
  bool zw = false;  // z passes of the wavefunction path run the warp-autonomous kernels (kernels_zw.h)
  bool mirror = false;        // mirror-pair x kernels usable (ray numbering mirror-symmetric; CPB_X_MIRROR=0 disables)
  int nbx_m = 0;              // their blocks per pair group
  double* d_kinpart = nullptr;  // kin_energy / dotp block partials of k_x_inv_m, [pair][nbx_m][4]
  size_t kinpart_cap = 0;       // pairs
  double* kin_cur = nullptr;    // set for the duration of a rhoofr call: partials of the call's pair 0
  const AxisKernels *kx = nullptr, *ky = nullptr, *kz = nullptr;
  // geometry (host copies)
  int xlo = 0, xhi = -1, zlo = 0, nzb = 0, nrays = 0, ref_nrays = 0, nrp = 0;
  std::vector<int32_t> nzhs, indzs;
  // device data
  PlanDev pd;
  // k-points: same plan data with the k-point gather table; set for the duration of a *_kpt call
  PlanDev pdk;
  uint32_t* d_gtab_k = nullptr;
  bool kpt_mode = false;
  const double *kpt_hgkp = nullptr, *kpt_hgkm = nullptr;
  int *d_ylo = nullptr, *d_yhi = nullptr, *d_rayoff = nullptr;
  uint32_t *d_gpos = nullptr, *d_gneg = nullptr, *d_gtab = nullptr;
  double* d_hg = nullptr;
  cplx *d_tw1 = nullptr, *d_tw2 = nullptr, *d_tw3 = nullptr;
  // Work spaces.  Consecutive batches of a call alternate between kNumWS work spaces, each with
  // its own stream: the HBM-bound x/y kernels of one batch overlap the FP64-bound z kernel of the
  // other, and the tail of every launch is filled by the next one.
  struct WorkSpace {
    cplx *T1 = nullptr, *T2 = nullptr;
    cplx* G = nullptr;  // band-ray storage of one x_sub sub-batch (written by k_x_fwd, read by k_unpack)
    cudaStream_t s = nullptr;
    rt::event_t ev_join = nullptr, ev_rho = nullptr;
  };
  static constexpr int kNumWS = 4;
  WorkSpace ws[kNumWS];
  rt::event_t vpot_event = nullptr;  // one-shot: the next vpsi waits for it before its first z pass
  // work spaces in use.  Two: consecutive batches alternate between two streams, so the tail of one batch's
  // kernels (last, partly filled wave; FP64-bound z pass) runs beside the head of the next batch (gather-bound x
  // pass): 192^3 29.3 -> 28.5 ms, 256^3 9.64 -> 9.47 ms per step (profiles/r03j/r03k_probe_streams.txt).  With the
  // round-1 kernels the overlap cost vpsi more than it gained; CPB_STREAMS=1 serialises the batches again.
  int nws = 2;
  rt::event_t ev_fork = nullptr;
  size_t workspace_bytes = 0;
  // per-call pair descriptors
  int pair_cap = 0;
  int *d_st1 = nullptr, *d_st2 = nullptr;
  double *d_ca = nullptr, *d_cb = nullptr;
  void* h_pairs = nullptr;  // pinned staging for the four arrays
  rt::event_t ev_pairs = nullptr;  // recorded behind the staging buffer's copies (CPB_ASYNC calls return before them)
  bool pairs_pending = false;
  // CPB_ASYNC rhoofr: what cpb_rhoofr_finish needs once the partial sums have arrived in h_red
  struct PendingRho {
    bool active = false;
    std::vector<double> f;
    int first = 0, nblk = 0, ngroups = 1;
    bool lsd = false;
    unsigned flags = 0;
    rt::event_t ev = nullptr;
  } pending_rho;
  // reductions
  int red_cap = 0;
  double* d_red = nullptr;
  double* h_red = nullptr;  // pinned
  // host-pointer API staging
  cplx* d_c0 = nullptr;
  size_t d_c0_cap = 0;
  cplx* d_c2 = nullptr;
  size_t d_c2_cap = 0;
  double* d_real = nullptr;  // rho or V, d_real_cols * nnr1 (2 columns with LSD)
  int d_real_cols = 0;
  cudaStream_t s_main = nullptr, s_in = nullptr, s_out = nullptr;
  std::vector<rt::event_t> ev_in, ev_done;
  // dense-transform staging (cpb_vofrho_local, host-pointer dense entry points)
  cplx* d_gbuf = nullptr;  // G-space scratch, d_gbuf_cap elements
  size_t d_gbuf_cap = 0;
  double* d_scg = nullptr;  // [ngw]
  // c0 cache key
  const void* c0_key_ptr = nullptr;
  long c0_key_ld = 0;
  int c0_key_nstate = 0, c0_key_ngroups = 0, c0_key_group = 0;
  bool c0_valid = false;
  long launches = 0;
  // optional per-kernel-class timing (cpb_plan_set_profiling): CUDA events around every launch
  bool profiling = false;
  struct Span {
    rt::event_t a, b;
    int kind;
  };
  std::vector<Span> spans;     // recorded in the current call
  std::vector<rt::event_t> ev_pool;
  double kind_ms[CPB_NKINDS] = {0};
  long kind_count[CPB_NKINDS] = {0};

  size_t nnr1() const { return (size_t)kr[0] * kr[1] * kr[2]; }
};

namespace {

void free_plan(cpb_plan* p) {
  if (!p) return;
  rt::dfree(p->d_ylo);
  rt::dfree(p->d_yhi);
  rt::dfree(p->d_rayoff);
  rt::dfree(p->d_gpos);
  rt::dfree(p->d_gneg);
  rt::dfree(p->d_gtab);
  rt::dfree(p->d_gtab_k);
  rt::dfree(p->d_hg);
  rt::dfree(p->d_tw1);
  rt::dfree(p->d_tw2);
  rt::dfree(p->d_tw3);
  for (auto& w : p->ws) {
    rt::dfree(w.T1);
    rt::dfree(w.T2);
    rt::dfree(w.G);
    rt::stream_destroy(w.s);
    rt::event_destroy(w.ev_join);
    rt::event_destroy(w.ev_rho);
  }
  rt::event_destroy(p->ev_fork);
  rt::event_destroy(p->ev_pairs);
  rt::event_destroy(p->pending_rho.ev);
  rt::dfree(p->T2keep);
  rt::dfree(p->d_st1);
  rt::dfree(p->d_st2);
  rt::dfree(p->d_ca);
  rt::dfree(p->d_cb);
  rt::hfree_pinned(p->h_pairs);
  rt::dfree(p->d_red);
  rt::dfree(p->d_kinpart);
  rt::hfree_pinned(p->h_red);
  rt::dfree(p->d_c0);
  rt::dfree(p->d_c2);
  rt::dfree(p->d_real);
  rt::dfree(p->d_gbuf);
  rt::dfree(p->d_scg);
  for (auto& sp : p->spans) {
    rt::event_destroy(sp.a);
    rt::event_destroy(sp.b);
  }
  for (auto e : p->ev_pool) rt::event_destroy(e);
  for (auto e : p->ev_in) rt::event_destroy(e);
  for (auto e : p->ev_done) rt::event_destroy(e);
  rt::stream_destroy(p->s_main);
  rt::stream_destroy(p->s_in);
  rt::stream_destroy(p->s_out);
  delete p;
}

std::vector<cplx> make_twiddles(int n) {
  std::vector<cplx> tw(n);
  const long double twopi = 6.283185307179586476925286766559005768L;
  for (int m = 0; m < n; ++m) {
    // exact octant values first, so trivial twiddles are exact
    if ((8 * m) % n == 0) {
      static const double c8[8] = {1, M_SQRT1_2, 0, -M_SQRT1_2, -1, -M_SQRT1_2, 0, M_SQRT1_2};
      static const double s8[8] = {0, M_SQRT1_2, 1, M_SQRT1_2, 0, -M_SQRT1_2, -1, -M_SQRT1_2};
      const int o = (8 * m) / n;
      tw[m] = mk(c8[o], s8[o]);
    } else {
      const long double a = twopi * (long double)m / (long double)n;
      tw[m] = mk((double)cosl(a), (double)sinl(a));
    }
  }
  return tw;
}

void ensure_pairs(cpb_plan* p, int npairs) {
  if (npairs <= p->pair_cap) return;
  const int cap = std::max(npairs, 2 * p->pair_cap);
  rt::dfree(p->d_st1);
  rt::dfree(p->d_st2);
  rt::dfree(p->d_ca);
  rt::dfree(p->d_cb);
  rt::hfree_pinned(p->h_pairs);
  p->d_st1 = (int*)rt::dmalloc(cap * sizeof(int));
  p->d_st2 = (int*)rt::dmalloc(cap * sizeof(int));
  p->d_ca = (double*)rt::dmalloc(cap * sizeof(double));
  p->d_cb = (double*)rt::dmalloc(cap * sizeof(double));
  p->h_pairs = rt::hmalloc_pinned((size_t)cap * (2 * sizeof(int) + 2 * sizeof(double)));
  p->pair_cap = cap;
}

void ensure_red(cpb_plan* p, int n) {
  if (n <= p->red_cap) return;
  if (p->pending_rho.active) throw Error(-1, "a CPB_ASYNC rhoofr is pending on this plan: call cpb_rhoofr_finish first");
  rt::dfree(p->d_red);
  rt::hfree_pinned(p->h_red);
  p->d_red = (double*)rt::dmalloc((size_t)n * sizeof(double));
  p->h_red = (double*)rt::hmalloc_pinned((size_t)n * sizeof(double));
  p->red_cap = n;
}

// stage the per-pair descriptors of one call; returns base PairDev
PairDev upload_pairs(cpb_plan* p, const std::vector<PairHost>& pairs, const std::vector<double>& ca,
                     const std::vector<double>& cb, cudaStream_t st) {
  const int np = (int)pairs.size();
  if (p->pairs_pending) {  // an enqueue-only call may still be copying out of the staging buffer
    rt::event_sync(p->ev_pairs);
    p->pairs_pending = false;
  }
  ensure_pairs(p, np);
  double* hca = (double*)p->h_pairs;
  double* hcb = hca + p->pair_cap;
  int* hs1 = (int*)(hcb + p->pair_cap);
  int* hs2 = hs1 + p->pair_cap;
  for (int i = 0; i < np; ++i) {
    hs1[i] = pairs[i].s1;
    hs2[i] = pairs[i].s2;
    hca[i] = ca[i];
    hcb[i] = cb[i];
  }
  if (np) {
    rt::h2d(p->d_st1, hs1, np * sizeof(int), st);
    rt::h2d(p->d_st2, hs2, np * sizeof(int), st);
    rt::h2d(p->d_ca, hca, np * sizeof(double), st);
    rt::h2d(p->d_cb, hcb, np * sizeof(double), st);
    if (!p->ev_pairs) p->ev_pairs = rt::event_create();
    rt::event_record(p->ev_pairs, st);
    p->pairs_pending = true;
  }
  PairDev pr;
  pr.st1 = p->d_st1;
  pr.st2 = p->d_st2;
  pr.ca = p->d_ca;
  pr.cb = p->d_cb;
  return pr;
}

PairDev offset_pairs(const PairDev& b, int off) {
  PairDev r;
  r.st1 = b.st1 + off;
  r.st2 = b.st2 + off;
  r.ca = b.ca + off;
  r.cb = b.cb + off;
  return r;
}

constexpr int kSumBlocks = 592;  // 4 x 148 SMs
constexpr int kRedPerState = 2 * kKinChunks;  // k_kin_energy partials per state

rt::event_t pool_event(cpb_plan* p) {
  if (!p->ev_pool.empty()) {
    rt::event_t e = p->ev_pool.back();
    p->ev_pool.pop_back();
    return e;
  }
  return rt::timing_event_create();
}

// RAII: brackets one kernel launch with events when profiling is on
struct Timed {
  cpb_plan* p;
  cudaStream_t st;
  int kind;
  rt::event_t a = nullptr;
  Timed(cpb_plan* p_, cudaStream_t st_, int kind_) : p(p_), st(st_), kind(kind_) {
    p->launches += 1;
    if (p->profiling) {
      a = pool_event(p);
      rt::event_record(a, st);
    }
  }
  ~Timed() {
    if (p->profiling) {
      rt::event_t b = pool_event(p);
      rt::event_record(b, st);
      p->spans.push_back({a, b, kind});
    }
  }
};

// after the stream has been synchronised: fold the recorded spans into the per-kind totals
void resolve_spans(cpb_plan* p) {
  for (auto& s : p->spans) {
    p->kind_ms[s.kind] += rt::event_elapsed_ms(s.a, s.b);
    p->kind_count[s.kind] += 1;
    p->ev_pool.push_back(s.a);
    p->ev_pool.push_back(s.b);
  }
  p->spans.clear();
}

// Pairs each block of an FFT kernel loops over.  A launch has `blocks_per_pair_group` blocks per
// group of pairs and the SM array holds `slots` blocks at a time, so the launch runs in
// ceil(blocks / slots) waves, each lasting (pairs per block + block prologue) pair-times: pick the
// group count that minimises that product.  Longer loops amortise the prologue and keep the
// prefetch pipeline full; more groups cut the cost of the last, partially filled wave.
int pairs_per_group(const cpb_plan* p, int npair, int blocks_per_pair_group, int blocks_per_sm, double prologue = -1.0) {
  if (prologue < 0.0) prologue = p->prologue_pairs;
  const long slots = (long)p->n_sm * std::max(blocks_per_sm, 1);
  int best = npair;
  double best_cost = 1e300;
  for (int groups = 1; groups <= npair; ++groups) {
    const int ppg = (npair + groups - 1) / groups;
    const int g = (npair + ppg - 1) / ppg;  // groups actually launched
    const long blocks = (long)blocks_per_pair_group * g;
    const long waves = (blocks + slots - 1) / slots;
    const double cost = (double)waves * (ppg + prologue);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = ppg;
    }
  }
  return best;
}

// pair groups for the elementwise G-space kernels: enough blocks to fill the machine a few times
int ew_ppg(const cpb_plan* p, int npair, int waves) {
  const int bx = (p->ngw + 255) / 256;
  int groups = (waves * p->n_sm + bx - 1) / bx;
  groups = std::max(1, std::min(groups, npair));
  return (npair + groups - 1) / groups;
}

// x pass, inverse: one launch per batch; the kernel gathers the coefficients from c0 itself.  Gamma point:
// the mirror-pair kernel (each coefficient fetched once; with p->kin_cur set it also accumulates the
// kin_energy / dotp partials of the batch's pairs, `off` = index of the batch's first pair in the call)
void run_x_inv(cpb_plan* p, cpb_plan::WorkSpace& w, const cplx* c0, long ldc, const PairDev& prb, int nb, int off = 0) {
  cudaStream_t st = w.s;
  Timed t(p, st, CPB_K_X_INV);
  if (p->xw_inv && !p->kpt_mode) {
    double* kin = p->kin_cur ? p->kin_cur + (size_t)off * p->nux_w * 4 : nullptr;
    const int blocks = (p->nux_w + p->kx->xw_warps - 1) / p->kx->xw_warps;
    p->kx->x_inv_w(st, c0, ldc, w.T1, p->pd, prb, nb,
                   pairs_per_group(p, nb, blocks, p->kx->xw_inv_blocks, p->prologue_pairs_x), kin, p->geq0);
    return;
  }
  if (p->mirror && !p->kpt_mode) {
    double* kin = p->kin_cur ? p->kin_cur + (size_t)off * p->nbx_m * 4 : nullptr;
    p->kx->x_inv_m(st, c0, ldc, w.T1, p->pd, prb, nb,
                   pairs_per_group(p, nb, p->nbx_m, p->kx->x_inv_m_blocks, p->prologue_pairs_xinv), p->half_x, kin, p->geq0);
    return;
  }
  p->kx->x_inv(st, c0, ldc, w.T1, p->kpt_mode ? p->pdk : p->pd, prb, nb,
               pairs_per_group(p, nb, p->nrp / p->kx->sl, p->kx->x_inv_blocks), p->half_x);
}

// room for the kin_energy / dotp block partials of `npairs` pairs; arms run_x_inv to produce them
void kin_begin(cpb_plan* p, int npairs, int nblk, cudaStream_t st) {
  p->kin_cur = nullptr;
  if (!p->mirror || p->kpt_mode || npairs <= 0) return;
  if ((size_t)npairs > p->kinpart_cap) {
    rt::dfree(p->d_kinpart);
    p->d_kinpart = nullptr;
    p->kinpart_cap = 0;
    p->d_kinpart = (double*)rt::dmalloc((size_t)npairs * std::max(p->nbx_m, p->nux_w) * 4 * sizeof(double));
    p->kinpart_cap = (size_t)npairs;
  }
  rt::dzero(p->d_red, (size_t)kRedPerState * nblk * sizeof(double), st);
  p->kin_cur = p->d_kinpart;
}
// fold the partials into d_red (k_kin_energy's layout, chunk 0); false if the separate pass is needed
bool kin_end(cpb_plan* p, const PairDev& pr, int npairs, int first, cudaStream_t st) {
  if (!p->kin_cur) return false;
  p->kin_cur = nullptr;
  auto k = k_kin_reduce;
  Timed t(p, st, CPB_K_KIN);
  CPB_LAUNCH(k, dim3(npairs), dim3(128), 4 * 128 * sizeof(double), st, (const double*)p->d_kinpart,
             p->xw_inv ? p->nux_w : p->nbx_m, pr, first, p->d_red);
  return true;
}

// x pass, forward, in sub-batches of x_sub pairs: k_x_fwd writes the sub-batch's band-ray storage
// (written to and re-read from HBM: it does not stay in L2 at useful sub-batch sizes), k_unpack gathers +G / -G
// from it and updates c2.
void run_x_fwd(cpb_plan* p, cpb_plan::WorkSpace& w, const cplx* c0, cplx* c2, long ldc, const PairDev& prb, int nb,
               bool accumulate) {
  cudaStream_t st = w.s;
  if (p->xw_fwd && !p->kpt_mode) {
    Timed t(p, st, CPB_K_X_FWD);
    const int blocks = (p->nux_w + p->kx->xw_warps - 1) / p->kx->xw_warps;
    p->kx->x_fwd_w(st, w.T1, c0, c2, ldc, p->pd, prb, nb,
                   pairs_per_group(p, nb, blocks, p->kx->xw_fwd_blocks, p->prologue_pairs_x), accumulate);
    return;
  }
  if (p->mirror && !p->kpt_mode) {
    // Gamma point: forward x pass fused with the unpack (no band-ray storage, one launch per batch)
    Timed t(p, st, CPB_K_X_FWD);
    p->kx->x_fwd_m(st, w.T1, c0, c2, ldc, p->pd, prb, nb,
                   pairs_per_group(p, nb, p->nbx_m, p->kx->x_fwd_m_blocks, p->prologue_pairs_x), p->half_x, accumulate);
    return;
  }
  for (int o = 0; o < nb; o += p->x_sub) {
    const int ns = std::min(p->x_sub, nb - o);
    PairDev prs = offset_pairs(prb, o);
    {
      Timed t(p, st, CPB_K_X_FWD);
      p->kx->x_fwd(st, w.T1 + (size_t)o * p->t1_pair, w.G, p->pd, ns,
                   pairs_per_group(p, ns, p->nrp / p->kx->sl, p->kx->x_fwd_blocks), p->half_x);
    }
    Timed t(p, st, CPB_K_UNPACK);
    const int ppg = ew_ppg(p, ns, 8);
    const dim3 grid((p->ngw + 255) / 256, (ns + ppg - 1) / ppg);
    if (p->kpt_mode) {
      if (accumulate) {
        auto k = k_unpack_kpt<true>;
        CPB_LAUNCH(k, grid, dim3(256), 0, st, (const cplx*)w.G, c0, c2, ldc, p->pd, prs, p->kpt_hgkp, p->kpt_hgkm,
                   p->geq0, ns, ppg);
      } else {
        auto k = k_unpack_kpt<false>;
        CPB_LAUNCH(k, grid, dim3(256), 0, st, (const cplx*)w.G, c0, c2, ldc, p->pd, prs, p->kpt_hgkp, p->kpt_hgkm,
                   p->geq0, ns, ppg);
      }
    } else if (accumulate) {
      auto k = k_unpack<true>;
      CPB_LAUNCH(k, grid, dim3(256), 0, st, (const cplx*)w.G, c0, c2, ldc, p->pd, prs, ns, ppg);
    } else {
      auto k = k_unpack<false>;
      CPB_LAUNCH(k, grid, dim3(256), 0, st, (const cplx*)w.G, c0, c2, ldc, p->pd, prs, ns, ppg);
    }
  }
}

// ------------------------------------------------------------------------------------------
// device-resident rhoofr over an explicit pair list (states are columns of c0 with stride ldc)
// `gate`, if non-null, is called before the kernels of batch b are enqueued (host API: wait for
// that batch's H2D).
// ------------------------------------------------------------------------------------------
struct BatchHooks {
  // called before / after the kernels of batch b are enqueued on stream `st`
  virtual void before_batch(int /*b*/, int /*pair0*/, int /*np*/, cudaStream_t /*st*/) {}
  virtual void after_batch(int /*b*/, int /*pair0*/, int /*np*/, cudaStream_t /*st*/) {}
  virtual ~BatchHooks() {}
};

// fork the work-space streams off the caller's stream / join them back
void fork_streams(cpb_plan* p, cudaStream_t st, int nbatches) {
  rt::event_record(p->ev_fork, st);
  for (int i = 0; i < std::min(p->nws, nbatches); ++i) rt::stream_wait(p->ws[i].s, p->ev_fork);
}
void join_streams(cpb_plan* p, cudaStream_t st, int nbatches) {
  for (int i = 0; i < std::min(p->nws, nbatches); ++i) {
    rt::event_record(p->ws[i].ev_join, p->ws[i].s);
    rt::stream_wait(st, p->ws[i].ev_join);
  }
}

// `pr`: the call's pair descriptors, already uploaded (upload_pairs) - the host-pointer entry points
// do that BEFORE they enqueue their bulk H2D copies, because the copy engine serves all streams in
// FIFO order and the first kernel would otherwise wait behind the whole upload.
// ---- CPB_PSI_KEEP / CPB_PSI_REUSE -------------------------------------------------------------
bool psi_key_match(const cpb_plan* p, const void* c0, long ld, int nstate, int ngroups, int my_group, int nsup,
                   int npairs) {
  return p->psi_valid && p->psi_key_ptr == c0 && p->psi_key[0] == ld && p->psi_key[1] == nstate &&
         p->psi_key[2] == ngroups && p->psi_key[3] == my_group && p->psi_key[4] == nsup && p->psi_npairs == npairs;
}
void psi_key_set(cpb_plan* p, const void* c0, long ld, int nstate, int ngroups, int my_group, int nsup, int npairs) {
  p->psi_key_ptr = c0;
  p->psi_key[0] = ld;
  p->psi_key[1] = nstate;
  p->psi_key[2] = ngroups;
  p->psi_key[3] = my_group;
  p->psi_key[4] = nsup;
  p->psi_npairs = npairs;
  p->psi_valid = true;
}
// room for the y-pass output of `npairs` pairs; false (and no cache) if the device has no room
bool psi_reserve(cpb_plan* p, int npairs) {
  p->psi_valid = false;
  if (p->chunk_xt != p->nxt) return false;  // the cache holds whole pairs
  if ((size_t)npairs <= p->t2keep_pairs) return true;
  rt::dfree(p->T2keep);
  p->T2keep = nullptr;
  p->t2keep_pairs = 0;
  try {
    p->T2keep = (cplx*)rt::dmalloc((size_t)npairs * p->t2_pair * sizeof(cplx));
  } catch (const Error&) {
    rt::check_last_clear();
    return false;
  }
  p->t2keep_pairs = (size_t)npairs;
  return true;
}

// z passes of the wavefunction path: warp-autonomous kernels where the plan can use them
void launch_z_rho(cpb_plan* p, cudaStream_t st, const cplx* T2, double* rho, const PairDev& pr, int nb, int xt0, int nxc) {
  if (p->zw) p->kz->z_rho_w(st, T2, rho, p->pd, pr, nb, xt0, nxc);
  else p->kz->z_rho(st, T2, rho, p->pd, pr, nb, xt0, nxc, p->half_z);
}
void launch_z_vpsi(cpb_plan* p, cudaStream_t st, cplx* T2, const double* vpot, int nb, int xt0, int nxc) {
  if (p->zw) {
    const AxisKernels* k = p->kz;
    const int gy = (p->nr[1] * k->zw_units_per_row + k->zw_warps - 1) / k->zw_warps;
    p->kz->z_vpsi_w(st, T2, vpot, p->pd, nb, xt0, nxc, pairs_per_group(p, nb, nxc * gy, k->zw_blocks_per_sm));
  } else {
    p->kz->z_vpsi(st, T2, vpot, p->pd, nb, xt0, nxc, pairs_per_group(p, nb, nxc * p->nr[1], p->kz->yz_blocks_per_sm),
                  p->half_z);
  }
}

// batches never straddle the channel boundary: [0, n0) work on channel 0, [n0, np) on channel 1
struct BatchSpan {
  int off, n, chan;
};
std::vector<BatchSpan> make_batches(const cpb_plan* p, int np, int n0) {
  std::vector<BatchSpan> out;
  const int mb = p->batch_override > 0 ? std::min(p->batch_override, p->max_batch) : p->max_batch;
  for (int off = 0; off < n0; off += mb) out.push_back({off, std::min(mb, n0 - off), 0});
  for (int off = n0; off < np; off += mb) out.push_back({off, std::min(mb, np - off), 1});
  return out;
}

// rho0 / rho1: the density arrays of channel 0 / 1 (no LSD: n0 == np, rho1 unused)
// keep: if non-null, pair i's y-pass output goes to keep + i * t2_pair and stays there (CPB_PSI_KEEP)
void run_rhoofr(cpb_plan* p, const cplx* c0, long ldc, const PairDev& pr, int np, int n0, double* rho0, double* rho1,
                cplx* keep, cudaStream_t st, BatchHooks* hooks) {
  const std::vector<BatchSpan> batches = make_batches(p, np, n0);
  const int nbatches = (int)batches.size();
  fork_streams(p, st, nbatches);
  for (int b = 0; b < nbatches; ++b) {
    const int off = batches[b].off, nb = batches[b].n;
    double* rho = batches[b].chan ? rho1 : rho0;
    cpb_plan::WorkSpace& w = p->ws[b % p->nws];
    if (hooks) hooks->before_batch(b, off, nb, w.s);
    PairDev prb = offset_pairs(pr, off);
    run_x_inv(p, w, c0, ldc, prb, nb, off);
    for (int xt0 = 0; xt0 < p->nxt; xt0 += p->chunk_xt) {
      const int nxc = std::min(p->chunk_xt, p->nxt - xt0);
      cplx* T2 = keep ? keep + (size_t)off * p->t2_pair : w.T2;
      { Timed t(p, w.s, CPB_K_Y_INV); p->ky->y_inv(w.s, w.T1, T2, p->pd, nb, xt0, nxc, pairs_per_group(p, nb, nxc * p->nzb, p->ky->yz_blocks_per_sm), p->half_y); }
      // rho is read-modify-written batch after batch: keep the order of the single-stream run
      if (b > 0 && p->nws > 1) rt::stream_wait(w.s, p->ws[(b - 1) % p->nws].ev_rho);
      { Timed t(p, w.s, CPB_K_Z_RHO); launch_z_rho(p, w.s, T2, rho, prb, nb, xt0, nxc); }
      rt::event_record(w.ev_rho, w.s);
    }
    if (hooks) hooks->after_batch(b, off, nb, w.s);
  }
  join_streams(p, st, nbatches);
  rt::check_last("rhoofr kernels");
}

// v0 / v1: the potentials of channel 0 / 1 (no LSD: n0 == np, v1 unused)
// reuse: if non-null, pair i starts from the kept y-pass output reuse + i * t2_pair (CPB_PSI_REUSE):
// no gather, no x and y inverse passes; the z pass consumes the cache in place
void run_vpsi(cpb_plan* p, const cplx* c0, cplx* c2, long ldc, const PairDev& pr, int np, int n0, const double* v0,
              const double* v1, bool accumulate, cplx* reuse, cudaStream_t st, BatchHooks* hooks) {
  const std::vector<BatchSpan> batches = make_batches(p, np, n0);
  const int nbatches = (int)batches.size();
  fork_streams(p, st, nbatches);
  rt::event_t vready = p->vpot_event;  // cpb_plan_set_vpot_event: only the z passes read vpot
  p->vpot_event = nullptr;
  for (int b = 0; b < nbatches; ++b) {
    const int off = batches[b].off, nb = batches[b].n;
    const double* vpot = batches[b].chan ? v1 : v0;
    cpb_plan::WorkSpace& w = p->ws[b % p->nws];
    if (hooks) hooks->before_batch(b, off, nb, w.s);
    PairDev prb = offset_pairs(pr, off);
    if (!reuse) run_x_inv(p, w, c0, ldc, prb, nb);
    for (int xt0 = 0; xt0 < p->nxt; xt0 += p->chunk_xt) {
      const int nxc = std::min(p->chunk_xt, p->nxt - xt0);
      const int ppg_y = pairs_per_group(p, nb, nxc * p->nzb, p->ky->yz_blocks_per_sm);
      cplx* T2 = reuse ? reuse + (size_t)off * p->t2_pair : w.T2;
      if (!reuse) { Timed t(p, w.s, CPB_K_Y_INV); p->ky->y_inv(w.s, w.T1, T2, p->pd, nb, xt0, nxc, ppg_y, p->half_y); }
      if (vready && b < p->nws && xt0 == 0) rt::stream_wait(w.s, vready);  // first z pass of every work-space stream
      { Timed t(p, w.s, CPB_K_Z_VPSI); launch_z_vpsi(p, w.s, T2, vpot, nb, xt0, nxc); }
      { Timed t(p, w.s, CPB_K_Y_FWD); p->ky->y_fwd(w.s, T2, w.T1, p->pd, nb, xt0, nxc, ppg_y, p->half_y); }
    }
    run_x_fwd(p, w, c0, c2, ldc, prb, nb, accumulate);
    if (hooks) hooks->after_batch(b, off, nb, w.s);
  }
  join_streams(p, st, nbatches);
  rt::check_last("vpsi kernels");
}

// kin_energy + dotp partial sums for states [first, first+count) of c0 -> d_red[0 .. kRedPerState*count)
void launch_kin(cpb_plan* p, const cplx* c0, long ldc, int first, int count, cudaStream_t st) {
  if (count <= 0) return;
  auto k = k_kin_energy;
  Timed t(p, st, CPB_K_KIN);
  CPB_LAUNCH(k, dim3(kKinChunks, count), dim3(256), 2 * 256 * sizeof(double), st, c0, ldc, first, p->ngw,
             p->geq0, (const double*)p->d_hg, p->d_red);
}

void launch_sum(cpb_plan* p, const double* a, size_t n, double* out, cudaStream_t st) {
  auto k = k_sum;
  Timed t(p, st, CPB_K_SUM);
  CPB_LAUNCH(k, dim3(kSumBlocks), dim3(256), 256 * sizeof(double), st, a, n, out);
}

// LSD: partial sums of alpha, beta and |alpha - beta| -> out[3 * kSumBlocks]; finalize: alpha += beta
void launch_lsd_sums(cpb_plan* p, double* a, const double* b, size_t n, double* out, bool finalize, cudaStream_t st) {
  auto k = k_lsd_sums;
  Timed t(p, st, CPB_K_SUM);
  CPB_LAUNCH(k, dim3(kSumBlocks), dim3(256), 3 * 256 * sizeof(double), st, a, b, n, out, finalize ? 1 : 0);
}

void vpsi_coefs(const std::vector<PairHost>& pairs, const double* f, bool tksham, std::vector<double>& fi,
                std::vector<double>& fip1) {
  // vpsi_utils.mod.F90:627-633
  fi.resize(pairs.size());
  fip1.resize(pairs.size());
  for (size_t i = 0; i < pairs.size(); ++i) {
    double a = f[pairs[i].s1] * 0.5;
    if (a == 0.0) a = tksham ? 0.5 : 1.0;
    double b = 0.0;
    if (pairs[i].s2 >= 0) b = f[pairs[i].s2] * 0.5;
    if (b == 0.0) b = tksham ? 0.5 : 1.0;
    fi[i] = a;
    fip1[i] = b;
  }
}

// keep_all: transform the unoccupied pairs too (they add nothing to rho, but CPB_PSI_KEEP wants the
// y-pass output of every pair, like rsactive forces tfcal, rhoofr_utils.mod.F90:312)
void rho_coefs(cpb_plan* p, const std::vector<PairHost>& all, const double* f, bool keep_all,
               std::vector<PairHost>& pairs, std::vector<double>& ca, std::vector<double>& cb) {
  // rhoofr_utils.mod.F90:312-316 (skip a pair only if both occupations vanish), :369-374
  for (const PairHost& q : all) {
    const double f1 = f[q.s1];
    const double f2 = q.s2 >= 0 ? f[q.s2] : 0.0;
    if (f1 == 0.0 && f2 == 0.0 && !keep_all) continue;
    pairs.push_back(q);
    ca.push_back(f1 / p->omega);
    cb.push_back(f2 / p->omega);
  }
}

// finish rhoofr: scalars from d_red (layout: [kRedPerState*count kin/dotp partials][rho partials:
// kSumBlocks sums, or with LSD 3*kSumBlocks: alpha, beta, |alpha-beta|])
void finish_rho_scalars(cpb_plan* p, const double* f, int first, int count, bool lsd, double* ekin, double* rsum_g,
                        double* rsum_r, double* csums, double* csumsabs) {
  double xkin = 0.0, rsum = 0.0;
  for (int i = 0; i < count; ++i) {
    const double fi = f[first + i];
    if (fi != 0.0) {  // kin_energy_utils.mod.F90:66
      double sk = 0.0, sd = 0.0;
      for (int c = 0; c < kKinChunks; ++c) {
        sk += p->h_red[(size_t)kRedPerState * i + 2 * c];
        sd += p->h_red[(size_t)kRedPerState * i + 2 * c + 1];
      }
      rsum += fi * sd;
      xkin += fi * sk;
    }
  }
  const double* hs = p->h_red + (size_t)kRedPerState * count;
  const double w = p->omega / ((double)p->nr[0] * p->nr[1] * p->nr[2]);
  double sa = 0.0, sb = 0.0, sabs = 0.0;
  for (int i = 0; i < kSumBlocks; ++i) sa += hs[i];
  if (lsd) {
    for (int i = 0; i < kSumBlocks; ++i) sb += hs[kSumBlocks + i];
    for (int i = 0; i < kSumBlocks; ++i) sabs += hs[2 * kSumBlocks + i];
  }
  if (ekin) *ekin = xkin * p->tpiba2;
  if (rsum_g) *rsum_g = rsum;
  if (rsum_r) *rsum_r = (sa + sb) * w;           // rhoofr_utils.mod.F90:607-619 (column 1 = alpha + beta)
  if (csums) *csums = lsd ? (sa - sb) * w : 0.0;  // :557
  if (csumsabs) *csumsabs = lsd ? sabs * w : 0.0; // :558 (meaningful only on the group-summed density)
}

int check_common(cpb_plan* p, const void* c0, long ld, int nstate, const double* f, int ngroups,
                 int my_group) {
  if (!p) return fail(CPB_ERR_INVALID, "null plan");
  if (!c0 || !f) return fail(CPB_ERR_INVALID, "null c0 or f");
  if (ld < p->ngw) return fail(CPB_ERR_INVALID, "leading dimension of c0 smaller than ngw");
  if (nstate < 0) return fail(CPB_ERR_INVALID, "negative nstate");
  if (ngroups < 1 || my_group < 0 || my_group >= ngroups)
    return fail(CPB_ERR_INVALID, "bad (ngroups, my_group)");
  return 0;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

const char* cpb_last_error(void) { return g_last_error.c_str(); }

const char* cpb_version(void) {
#if defined(CPB_EMULATE)
  return "cpb200 0.1 (CPU kernel simulator build - tests only)";
#else
  return "cpb200 0.1 (sm_100a)";
#endif
}

int cpb_length_supported(int n) { return find_axis_kernels(n) ? 1 : 0; }

int cpb_part_1d_nbr_el_in_blk(int n_elem, int proc, int nproc) { return nbr_el_in_blk(n_elem, proc, nproc); }
int cpb_part_1d_get_el_in_blk(int i_elem, int n_elem, int proc, int nproc) {
  return get_el_in_blk(i_elem, n_elem, proc, nproc);
}

int cpb_plan_create(cpb_plan** out, const int* nr, const int* kr, int ngw, const int32_t* inyh,
                    const double* hg, double tpiba2, double omega, int device, int max_batch_pairs) {
  if (!out || !nr || !kr || !inyh || !hg) return fail(CPB_ERR_INVALID, "null argument");
  *out = nullptr;
  if (ngw <= 0) return fail(CPB_ERR_INVALID, "ngw must be positive");
  if (omega <= 0.0) return fail(CPB_ERR_INVALID, "omega must be positive");
  for (int d = 0; d < 3; ++d) {
    if (nr[d] < 2 || kr[d] < nr[d]) return fail(CPB_ERR_INVALID, "bad mesh / leading dimensions");
  }
  cpb_plan* p = nullptr;
  try {
    p = new cpb_plan();
    for (int d = 0; d < 3; ++d) {
      p->nr[d] = nr[d];
      p->kr[d] = kr[d];
    }
    p->ngw = ngw;
    p->tpiba2 = tpiba2;
    p->omega = omega;
    p->device = device;
    p->max_batch = std::min(max_batch_pairs > 0 ? max_batch_pairs : 32, (int)kMaxGroup);
    p->auto_batch = max_batch_pairs <= 0;
    p->kx = find_axis_kernels(nr[0]);
    p->ky = find_axis_kernels(nr[1]);
    p->kz = find_axis_kernels(nr[2]);
    if (!p->kx || !p->ky || !p->kz) {
      char buf[160];
      std::snprintf(buf, sizeof buf, "mesh %dx%dx%d: a length has no kernel instantiation (see sizes.def)",
                    nr[0], nr[1], nr[2]);
      throw Error(CPB_ERR_UNSUPPORTED, buf);
    }
    const int n1 = nr[0], n2 = nr[1], n3 = nr[2];
    const int SL = p->kx->sl;

    // ---- 0-based box positions; the mirror of g is n-g (inyh -> 2*nh-inyh, fftprp :272-277)
    std::vector<int> gx(ngw), gy(ngw), gz(ngw);
    for (int i = 0; i < ngw; ++i) {
      gx[i] = inyh[3 * (size_t)i + 0] - 1;
      gy[i] = inyh[3 * (size_t)i + 1] - 1;
      gz[i] = inyh[3 * (size_t)i + 2] - 1;
      if (gx[i] < 1 || gx[i] > n1 - 1 || gy[i] < 1 || gy[i] > n2 - 1 || gz[i] < 1 || gz[i] > n3 - 1)
        throw Error(CPB_ERR_INVALID, "inyh entry (or its -G mirror) falls outside the mesh");
    }
    p->geq0 = (gx[0] == n1 / 2 && gy[0] == n2 / 2 && gz[0] == n3 / 2) ? 1 : 0;
    for (int i = 1; i < ngw; ++i) {
      if (gx[i] == n1 / 2 && gy[i] == n2 / 2 && gz[i] == n3 / 2)
        throw Error(CPB_ERR_INVALID, "G=0 must be the first plane wave (loadpa sort order)");
    }

    // ---- ray marks (fftprp_utils.mod.F90:145-156)
    std::vector<unsigned char> mark((size_t)n2 * n3, 0);
    int xlo = n1, xhi = -1;
    for (int i = 0; i < ngw; ++i) {
      mark[(size_t)gz[i] * n2 + gy[i]] = 1;
      mark[(size_t)(n3 - gz[i]) * n2 + (n2 - gy[i])] = 1;
      xlo = std::min(xlo, std::min(gx[i], n1 - gx[i]));
      xhi = std::max(xhi, std::max(gx[i], n1 - gx[i]));
    }
    int zlo = n3, zhi = -1;
    for (int z = 0; z < n3; ++z)
      for (int y = 0; y < n2; ++y)
        if (mark[(size_t)z * n2 + y]) {
          zlo = std::min(zlo, z);
          zhi = std::max(zhi, z);
        }
    const int nzb = zhi - zlo + 1;
    std::vector<int> ylo(nzb), yhi(nzb), rayoff(nzb);
    // reference ray numbering (z outer, y inner, marked rays only; fftprp :209-217)
    std::vector<int> refray((size_t)n2 * n3, -1);
    int nref = 0, nrays = 0;
    for (int z = zlo; z <= zhi; ++z) {
      int lo = n2, hi = -1;
      for (int y = 0; y < n2; ++y)
        if (mark[(size_t)z * n2 + y]) {
          refray[(size_t)z * n2 + y] = nref++;
          lo = std::min(lo, y);
          hi = std::max(hi, y);
        }
      if (hi < 0) {
        lo = 1;
        hi = 0;
      }
      ylo[z - zlo] = lo;
      yhi[z - zlo] = hi;
      rayoff[z - zlo] = nrays;
      nrays += hi - lo + 1;
    }
    int nyb = 1;
    for (int zr = 0; zr < nzb; ++zr) nyb = std::max(nyb, yhi[zr] - ylo[zr] + 1);
    auto ray_of = [&](int y, int z) -> int {
      const int zr = z - zlo;
      if (zr < 0 || zr >= nzb || y < ylo[zr] || y > yhi[zr]) return -1;
      return rayoff[zr] + (y - ylo[zr]);
    };
    p->nzhs.resize(ngw);
    p->indzs.resize(ngw);
    for (int i = 0; i < ngw; ++i) {
      p->nzhs[i] = gx[i] + 1 + refray[(size_t)gz[i] * n2 + gy[i]] * kr[0];
      p->indzs[i] = (n1 - gx[i]) + 1 + refray[(size_t)(n3 - gz[i]) * n2 + (n2 - gy[i])] * kr[0];
    }

    // ---- band-ray storage positions of +G / -G (k_x_inv's gtab, k_unpack's gpos/gneg): xb * nrp + ray
    const int nxb = xhi - xlo + 1;
    const int nrp = (nrays + SL - 1) / SL * SL;
    if ((double)nxb * nrp > 4.0e9) throw Error(CPB_ERR_UNSUPPORTED, "band-ray storage exceeds 32-bit positions");
    std::vector<uint32_t> gpos(ngw), gneg(ngw), gtab((size_t)nxb * nrp, kNoPW);
    if (2.0 * ngw >= (double)kNegPW) throw Error(CPB_ERR_UNSUPPORTED, "ngw too large");
    {
      std::vector<unsigned char> occ((size_t)nrays * n1, 0);
      for (int i = 0; i < ngw; ++i) {
        const int rp = ray_of(gy[i], gz[i]);
        const int rm = ray_of(n2 - gy[i], n3 - gz[i]);
        if (rp < 0 || rm < 0) throw Error(CPB_ERR_INVALID, "internal: ray set is not mirror symmetric");
        const uint32_t lp = (uint32_t)(gx[i] - xlo) * (uint32_t)nrp + (uint32_t)rp;
        const uint32_t lm = (uint32_t)(n1 - gx[i] - xlo) * (uint32_t)nrp + (uint32_t)rm;
        if (occ[(size_t)rp * n1 + gx[i]]++) throw Error(CPB_ERR_INVALID, "duplicate plane wave in inyh");
        if (lm != lp) {
          if (occ[(size_t)rm * n1 + (n1 - gx[i])]++)
            throw Error(CPB_ERR_INVALID, "inyh contains both G and -G (half-sphere list expected)");
        } else if (i != 0 || !p->geq0) {
          throw Error(CPB_ERR_INVALID, "self-mirrored plane wave that is not G=0");
        }
        gpos[i] = lp;
        gneg[i] = lm;
        if (lm != lp) gtab[lm] = (uint32_t)i | kNegPW;
        gtab[lp] = (uint32_t)i;  // G = 0: the +G form is the one stored (state_utils.mod.F90:187)
      }
    }

    {
      // band-pruned kernel variants (KRange): index band inside [r2*klo, r2*khi) of the axis
      int ymin = n2, ymax = -1;
      for (int zr = 0; zr < nzb; ++zr)
        if (yhi[zr] >= ylo[zr]) {
          ymin = std::min(ymin, ylo[zr]);
          ymax = std::max(ymax, yhi[zr]);
        }
      auto fits = [](const AxisKernels* k, int lo, int hi) {
        return lo >= k->r2 * k->klo && hi < k->r2 * k->khi;
      };
      p->half_x = fits(p->kx, xlo, xhi);
      p->half_y = fits(p->ky, ymin, ymax);
      p->half_z = fits(p->kz, zlo, zhi);
      if (const char* e = std::getenv("CPB_NO_HALF")) {
        if (std::atoi(e)) p->half_x = p->half_y = p->half_z = false;
      }
      // warp-autonomous z kernels: band-pruned instantiation only, own factorisation of n3
      // (opt-in with CPB_ZW=1 while they are slower than the block kernels on the 192^3 case: 2.13 vs 1.99 ms
      // z_rho, 4.17 vs 3.53 ms z_vpsi per 128 pairs, profiles/r02k_full_zw_192x128.txt)
      const bool zw_ok = p->half_z && p->kz->z_rho_w && zlo >= p->kz->zw_rb * p->kz->zw_klo && zhi < p->kz->zw_rb * p->kz->zw_khi;
      p->zw = false;
      if (const char* e = std::getenv("CPB_ZW")) p->zw = zw_ok && std::atoi(e) != 0;
    }
    {
      // mirror-pair x kernels: the mirror of internal ray r must be ray nrays - 1 - r (true whenever the ray
      // set is symmetric under (y,z) -> (n2-y, n3-z), which the -G marks above guarantee) and the centre of
      // the x axis must lie in the decimated index R1/2 of the kernels' factorisation (n1 even)
      bool ok = (n1 % 2 == 0);
      for (int zr = 0; zr < nzb && ok; ++zr)
        for (int y = ylo[zr]; y <= yhi[zr] && ok; ++y) {
          const int r = rayoff[zr] + (y - ylo[zr]);
          ok = ray_of(n2 - y, n3 - (zlo + zr)) == nrays - 1 - r;
        }
      p->mirror = ok;
      if (const char* e = std::getenv("CPB_X_MIRROR")) {
        if (!std::atoi(e)) p->mirror = false;
      }
      p->nbx_m = ((nrays + 1) / 2 + SL / 2 - 1) / (SL / 2);
      // warp-autonomous variants: band-pruned only, own factorisation of n1 (RB = 8)
      const bool xw_ok = p->mirror && p->half_x && p->kx->x_inv_w && xlo >= 8 * p->kx->xw_klo && xhi < 8 * p->kx->xw_khi;
      if (xw_ok) p->nux_w = ((nrays + 1) / 2 + p->kx->xw_rays - 1) / p->kx->xw_rays;
      if (const char* e = std::getenv("CPB_XW")) {
        const int v = std::atoi(e);  // bit 0: inverse, bit 1: forward
        p->xw_inv = xw_ok && (v & 1);
        p->xw_fwd = xw_ok && (v & 2) && p->kx->x_fwd_w;
      }
    }
    p->xlo = xlo;
    p->xhi = xhi;
    p->zlo = zlo;
    p->nzb = nzb;
    p->nrays = nrays;
    p->ref_nrays = nref;
    p->nrp = nrp;

    // ---- device side
    rt::set_device(device);
    p->n_sm = rt::sm_count(device);
    p->d_ylo = upload(ylo);
    p->d_yhi = upload(yhi);
    p->d_rayoff = upload(rayoff);
    p->d_gpos = upload(gpos);
    p->d_gneg = upload(gneg);
    p->d_gtab = upload(gtab);
    {
      // k-point gather table: the -G position of plane wave ig takes c0(ig + ngw), unconjugated
      std::vector<uint32_t> gtab_k(gtab);
      for (int i = 0; i < ngw; ++i)
        if (gneg[i] != gpos[i]) gtab_k[gneg[i]] = (uint32_t)i + (uint32_t)ngw;
      p->d_gtab_k = upload(gtab_k);
    }
    p->d_hg = upload(std::vector<double>(hg, hg + ngw));
    p->d_tw1 = upload(make_twiddles(n1));
    p->d_tw2 = upload(make_twiddles(n2));
    p->d_tw3 = upload(make_twiddles(n3));
    // T1[pair][xt][ray][B]; T2[pair][xtc][y][zr][B].  By default T2 holds all x tiles of the batch
    // (one y launch + one z launch per batch).  Measured on B200 (profiles/r01b_sweep_chunk.txt):
    // splitting the y/z passes into L2-sized chunks of x tiles loses more to small grids and launch
    // gaps than the L2 residency of T2 saves, so chunking is a tuning hook only (CPB_CHUNK_XT).
    const int Bx = p->kx->b;
    p->nxt = (n1 + Bx - 1) / Bx;
    p->chunk_xt = p->nxt;
    if (const char* e = std::getenv("CPB_CHUNK_XT")) p->chunk_xt = std::max(1, std::min(p->nxt, std::atoi(e)));
    p->t1_pair = (size_t)p->nxt * nrays * Bx;
    if (const char* e = std::getenv("CPB_X_SUB")) p->x_sub = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("CPB_PROLOGUE")) p->prologue_pairs = std::max(0.0, std::atof(e));
    if (const char* e = std::getenv("CPB_PROLOGUE_X")) p->prologue_pairs_x = p->prologue_pairs_xinv = std::max(0.0, std::atof(e));
    if (const char* e = std::getenv("CPB_PROLOGUE_XINV")) p->prologue_pairs_xinv = std::max(0.0, std::atof(e));
    p->g_pair = (size_t)nxb * nrp;
    p->t2_pair = (size_t)p->nxt * n2 * nzb * Bx;
    if (p->auto_batch) {
      // default batch: about 3 GB of T1 + T2, a multiple of 8 pairs in [8, 64].  Measured on B200
      // (profiles/r02n_probe_batch.txt): small meshes gain from long batches (96^3: 1.12 -> 1.02 ms per step with
      // 64 instead of 32 pairs, 120^3: 2.19 -> 2.04), 192^3 is flat between 32 and 64
      const double per_pair = (double)(p->t1_pair + p->t2_pair) * sizeof(cplx);
      const int fit = (int)(3.0e9 / std::max(per_pair, 1.0));
      p->max_batch = std::max(8, std::min((int)kMaxGroup, fit / 8 * 8));
    }
    p->x_sub = std::min(p->x_sub, p->max_batch);
    const size_t gb = (size_t)p->x_sub * p->g_pair * sizeof(cplx);
    const size_t t1 = (size_t)p->max_batch * p->nxt * nrays * Bx * sizeof(cplx);
    const size_t t2 = (size_t)p->max_batch * p->chunk_xt * n2 * nzb * Bx * sizeof(cplx);
    // host-pointer calls: batches of at least 8 pairs and about 24 MB of coefficients (16 ngw bytes per state)
    p->host_batch = std::max(8, (int)std::ceil(24.0e6 / (32.0 * std::max(p->ngw, 1))));
    if (const char* e = std::getenv("CPB_HOST_BATCH")) p->host_batch = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("CPB_STREAMS")) p->nws = std::max(1, std::min((int)cpb_plan::kNumWS, std::atoi(e)));
    for (int i = 0; i < p->nws; ++i) {
      cpb_plan::WorkSpace& w = p->ws[i];
      w.T1 = (cplx*)rt::dmalloc(t1);
      w.T2 = (cplx*)rt::dmalloc(t2);
      w.G = (cplx*)rt::dmalloc(gb);
      // pad columns (x >= n1 in the last x tile) are never written by the x pass: keep them finite.
      rt::dzero(w.T1, t1, 0);
      rt::dzero(w.T2, t2, 0);
      rt::dzero(w.G, gb, 0);
      w.s = rt::stream_create();
      w.ev_join = rt::event_create();
      w.ev_rho = rt::event_create();
    }
    p->ev_fork = rt::event_create();
    rt::sync(0);
    p->workspace_bytes = (size_t)p->nws * (t1 + t2 + gb);
    p->s_main = rt::stream_create();
    p->s_in = rt::stream_create();
    p->s_out = rt::stream_create();

    PlanDev& pd = p->pd;
    pd.n1 = n1;
    pd.n2 = n2;
    pd.n3 = n3;
    pd.kr1 = kr[0];
    pd.kr2 = kr[1];
    pd.kr3 = kr[2];
    pd.xlo = xlo;
    pd.nxb = nxb;
    pd.zlo = zlo;
    pd.nzb = nzb;
    pd.nrays = nrays;
    pd.nrp = nrp;
    pd.nyb = nyb;
    pd.nxt = p->nxt;
    pd.ngw = ngw;
    pd.ylo = p->d_ylo;
    pd.yhi = p->d_yhi;
    pd.rayoff = p->d_rayoff;
    pd.gtab = p->d_gtab;
    pd.gpos = p->d_gpos;
    pd.gneg = p->d_gneg;
    pd.hg = p->d_hg;
    pd.tw1 = p->d_tw1;
    pd.tw2 = p->d_tw2;
    pd.tw3 = p->d_tw3;
    pd.tpiba2 = tpiba2;
    pd.inv_n = 1.0 / ((double)n1 * n2 * n3);
    p->pdk = pd;
    p->pdk.gtab = p->d_gtab_k;
    *out = p;
    return CPB_OK;
  } catch (const Error& e) {
    free_plan(p);
    return fail(e.code, e.what());
  } catch (const std::bad_alloc&) {
    free_plan(p);
    return fail(CPB_ERR_NOMEM, "out of host memory");
  }
}

int cpb_plan_destroy(cpb_plan* plan) {
  if (!plan) return CPB_OK;
  try {
    rt::set_device(plan->device);
  } catch (...) {
  }
  free_plan(plan);
  return CPB_OK;
}

int cpb_plan_get_info(const cpb_plan* p, cpb_plan_info* info) {
  if (!p || !info) return fail(CPB_ERR_INVALID, "null argument");
  for (int d = 0; d < 3; ++d) {
    info->nr[d] = p->nr[d];
    info->kr[d] = p->kr[d];
  }
  info->ngw = p->ngw;
  info->geq0 = p->geq0;
  info->nrays = p->ref_nrays;
  info->zband = p->nzb;
  info->xband = p->xhi - p->xlo + 1;
  info->max_batch = p->max_batch;
  info->device = p->device;
  const AxisKernels* ks[3] = {p->kx, p->ky, p->kz};
  for (int d = 0; d < 3; ++d) {
    info->radix[d][0] = ks[d]->r1;
    info->radix[d][1] = ks[d]->r2;
  }
  info->workspace_bytes = p->workspace_bytes;
  info->band_pruned[0] = p->half_x;
  info->band_pruned[1] = p->half_y;
  info->band_pruned[2] = p->half_z;
  info->chunk_xtiles = p->chunk_xt;
  info->streams = p->nws;
  info->z_warp_kernels = p->zw ? 1 : 0;
  info->z_warp_radix[0] = p->zw ? p->kz->zw_ra : 0;
  info->z_warp_radix[1] = p->zw ? p->kz->zw_rb : 0;
  info->x_warp_kernels = (p->xw_inv ? 1 : 0) | (p->xw_fwd ? 2 : 0);
  info->x_warp_radix = (p->xw_inv || p->xw_fwd) ? p->kx->xw_ra : 0;
  return CPB_OK;
}

int cpb_plan_get_maps(const cpb_plan* p, int32_t* nzhs, int32_t* indzs) {
  if (!p || !nzhs || !indzs) return fail(CPB_ERR_INVALID, "null argument");
  std::copy(p->nzhs.begin(), p->nzhs.end(), nzhs);
  std::copy(p->indzs.begin(), p->indzs.end(), indzs);
  return CPB_OK;
}

long cpb_plan_launch_count(const cpb_plan* p) { return p ? p->launches : 0; }

int cpb_plan_set_streams(cpb_plan* p, int n) {
  if (!p) return fail(CPB_ERR_INVALID, "null plan");
  int have = 0;
  for (const auto& w : p->ws) have += (w.T1 != nullptr);
  if (n < 1 || n > have) return fail(CPB_ERR_INVALID, "stream count outside 1..allocated work spaces");
  p->nws = n;
  return CPB_OK;
}

int cpb_plan_set_vpot_event(cpb_plan* p, void* event) {
  if (!p) return fail(CPB_ERR_INVALID, "null plan");
  p->vpot_event = (rt::event_t)event;
  return CPB_OK;
}

int cpb_plan_set_profiling(cpb_plan* p, int on) {
  if (!p) return fail(CPB_ERR_INVALID, "null plan");
  p->profiling = on != 0;
  return CPB_OK;
}

int cpb_plan_get_kernel_times(cpb_plan* p, double* ms, long* counts, int reset) {
  if (!p || !ms || !counts) return fail(CPB_ERR_INVALID, "null argument");
  for (int k = 0; k < CPB_NKINDS; ++k) {
    ms[k] = p->kind_ms[k];
    counts[k] = p->kind_count[k];
    if (reset) {
      p->kind_ms[k] = 0.0;
      p->kind_count[k] = 0;
    }
  }
  return CPB_OK;
}

// ---------------------------------------------------------------------------------------------
// device-pointer entry points
// ---------------------------------------------------------------------------------------------
static int rhoofr_dev_impl(cpb_plan* p, const void* c0_dev, long ld_c0, int nstate, const double* f, int nsup,
                           int ngroups, int my_group, double* rhoe_dev, double* ekin, double* rsum_g,
                           double* rsum_r, double* csums, double* csumsabs, unsigned flags, void* stream) {
  if (int e = check_common(p, c0_dev, ld_c0, nstate, f, ngroups, my_group)) return e;
  if (!rhoe_dev) return fail(CPB_ERR_INVALID, "null rhoe");
  if (nsup > nstate) return fail(CPB_ERR_INVALID, "nsup larger than nstate");
  if (p->pending_rho.active)
    return fail(CPB_ERR_INVALID, "a CPB_ASYNC rhoofr is pending on this plan: call cpb_rhoofr_finish first");
  try {
    rt::set_device(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const cplx* c0 = (const cplx*)c0_dev;
    const int nblk = nbr_el_in_blk(nstate, my_group, ngroups);
    const int first = nblk > 0 ? get_el_in_blk(1, nstate, my_group, ngroups) - 1 : 0;
    std::vector<PairHost> pairs;
    std::vector<double> ca, cb;
    const bool lsd = nsup >= 0;
    const size_t nnr1 = p->nnr1();
    const std::vector<PairHost> all = block_pairs(nstate, my_group, ngroups, nsup);
    const bool keep = (flags & CPB_PSI_KEEP) && psi_reserve(p, (int)all.size());
    if (!keep) p->psi_valid = false;
    rho_coefs(p, all, f, keep, pairs, ca, cb);
    ensure_red(p, kRedPerState * nblk + 3 * kSumBlocks);
    rt::dzero(rhoe_dev, (lsd ? 2 : 1) * nnr1 * sizeof(double), st);  // rhoofr_utils.mod.F90:198
    const PairDev prd = upload_pairs(p, pairs, ca, cb, st);
    kin_begin(p, (int)pairs.size(), nblk, st);                        // :178 kin_energy rides on the gather ...
    run_rhoofr(p, c0, ld_c0, prd, (int)pairs.size(), count_chan0(pairs), rhoe_dev, rhoe_dev + nnr1,
               keep ? p->T2keep : nullptr, st, nullptr);
    if (!kin_end(p, prd, (int)pairs.size(), first, st)) launch_kin(p, c0, ld_c0, first, nblk, st);  // ... or runs alone
    if (keep) psi_key_set(p, c0_dev, ld_c0, nstate, ngroups, my_group, nsup, (int)pairs.size());
    double* d_sums = p->d_red + kRedPerState * nblk;
    if (lsd) launch_lsd_sums(p, rhoe_dev, rhoe_dev + nnr1, nnr1, d_sums, ngroups == 1, st);  // :543-559
    else launch_sum(p, rhoe_dev, nnr1, d_sums, st);                                            // :607-619
    // only what the kernels above wrote travels (LSD: three sums per block, otherwise one)
    rt::d2h(p->h_red, p->d_red, (size_t)(kRedPerState * nblk + (lsd ? 3 : 1) * kSumBlocks) * sizeof(double), st);
    if ((flags & CPB_ASYNC) && !p->profiling) {
      // enqueue-only: the partial sums are on their way to h_red; cpb_rhoofr_finish waits for them
      cpb_plan::PendingRho& pr = p->pending_rho;
      if (!pr.ev) pr.ev = rt::event_create();
      rt::event_record(pr.ev, st);
      pr.active = true;
      pr.f.assign(f, f + nstate);
      pr.first = first;
      pr.nblk = nblk;
      pr.ngroups = ngroups;
      pr.lsd = lsd;
      pr.flags = flags;
      return CPB_OK;
    }
    rt::sync(st);
    resolve_spans(p);
    double rg = 0, rr = 0;
    finish_rho_scalars(p, f, first, nblk, lsd, ekin, &rg, &rr, csums, csumsabs);
    if (rsum_g) *rsum_g = rg;
    if (rsum_r) *rsum_r = rr;
    if ((flags & CPB_RHO_CHECK_CHARGE) && ngroups == 1 && std::fabs(rr - rg) > 1.0e-6)
      return fail(CPB_ERR_CHARGE, "TOTAL DENSITY SUMS ARE NOT EQUAL");  // :625-635
    return CPB_OK;
  } catch (const Error& e) {
    return fail(e.code, e.what());
  } catch (const std::bad_alloc&) {
    return fail(CPB_ERR_NOMEM, "out of host memory");
  }
}

int cpb_rhoofr_dev(cpb_plan* p, const void* c0_dev, long ld_c0, int nstate, const double* f, int ngroups,
                   int my_group, double* rhoe_dev, double* ekin, double* rsum_g, double* rsum_r,
                   unsigned flags, void* stream) {
  return rhoofr_dev_impl(p, c0_dev, ld_c0, nstate, f, -1, ngroups, my_group, rhoe_dev, ekin, rsum_g, rsum_r, nullptr,
                         nullptr, flags, stream);
}

int cpb_rhoofr_lsd_dev(cpb_plan* p, const void* c0_dev, long ld_c0, int nstate, const double* f, int nsup,
                       int ngroups, int my_group, double* rhoe_dev, double* ekin, double* rsum_g, double* rsum_r,
                       double* csums, double* csumsabs, unsigned flags, void* stream) {
  if (nsup < 0) return fail(CPB_ERR_INVALID, "negative nsup");
  return rhoofr_dev_impl(p, c0_dev, ld_c0, nstate, f, nsup, ngroups, my_group, rhoe_dev, ekin, rsum_g, rsum_r, csums,
                         csumsabs, flags, stream);
}

int cpb_rhoofr_pending(cpb_plan* p) { return (p && p->pending_rho.active) ? 1 : 0; }

int cpb_rhoofr_finish(cpb_plan* p, double* ekin, double* rsum_g, double* rsum_r, double* csums, double* csumsabs) {
  if (!p) return fail(CPB_ERR_INVALID, "null plan");
  cpb_plan::PendingRho& pr = p->pending_rho;
  if (!pr.active) return fail(CPB_ERR_INVALID, "no CPB_ASYNC rhoofr is pending on this plan");
  try {
    rt::set_device(p->device);
    rt::event_sync(pr.ev);
    pr.active = false;
    rt::check_last("rhoofr kernels");
    double rg = 0, rr = 0;
    finish_rho_scalars(p, pr.f.data(), pr.first, pr.nblk, pr.lsd, ekin, &rg, &rr, csums, csumsabs);
    if (rsum_g) *rsum_g = rg;
    if (rsum_r) *rsum_r = rr;
    if ((pr.flags & CPB_RHO_CHECK_CHARGE) && pr.ngroups == 1 && std::fabs(rr - rg) > 1.0e-6)
      return fail(CPB_ERR_CHARGE, "TOTAL DENSITY SUMS ARE NOT EQUAL");  // rhoofr_utils.mod.F90:625-635
    return CPB_OK;
  } catch (const Error& e) {
    return fail(e.code, e.what());
  }
}

int cpb_lsd_finish_dev(cpb_plan* p, double* rhoe_dev, double* rsum_r, double* csums, double* csumsabs,
                       void* stream) {
  if (!p || !rhoe_dev) return fail(CPB_ERR_INVALID, "null argument");
  try {
    rt::set_device(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    ensure_red(p, 3 * kSumBlocks);
    launch_lsd_sums(p, rhoe_dev, rhoe_dev + p->nnr1(), p->nnr1(), p->d_red, true, st);
    rt::d2h(p->h_red, p->d_red, (size_t)3 * kSumBlocks * sizeof(double), st);
    rt::sync(st);
    resolve_spans(p);
    finish_rho_scalars(p, nullptr, 0, 0, true, nullptr, nullptr, rsum_r, csums, csumsabs);
    return CPB_OK;
  } catch (const Error& e) {
    return fail(e.code, e.what());
  }
}

static int vpsi_dev_impl(cpb_plan* p, const void* c0_dev, void* c2_dev, long ld, int nstate, const double* f,
                         int nsup, const double* vpot_dev, int ngroups, int my_group, unsigned flags, void* stream) {
  if (int e = check_common(p, c0_dev, ld, nstate, f, ngroups, my_group)) return e;
  if (!c2_dev || !vpot_dev) return fail(CPB_ERR_INVALID, "null c2 or vpot");
  if (nsup > nstate) return fail(CPB_ERR_INVALID, "nsup larger than nstate");
  try {
    rt::set_device(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<PairHost> pairs = block_pairs(nstate, my_group, ngroups, nsup);
    std::vector<double> fi, fip1;
    vpsi_coefs(pairs, f, (flags & CPB_VPSI_TKSHAM) != 0, fi, fip1);
    const bool reuse = (flags & CPB_PSI_REUSE) &&
                       psi_key_match(p, c0_dev, ld, nstate, ngroups, my_group, nsup, (int)pairs.size());
    p->psi_valid = false;  // the z pass consumes the cache in place
    run_vpsi(p, (const cplx*)c0_dev, (cplx*)c2_dev, ld, upload_pairs(p, pairs, fi, fip1, st), (int)pairs.size(),
             count_chan0(pairs), vpot_dev, vpot_dev + p->nnr1(),
             !(flags & CPB_VPSI_OVERWRITE), reuse ? p->T2keep : nullptr, st, nullptr);
    if ((flags & CPB_ASYNC) && !p->profiling) return CPB_OK;  // enqueue-only
    rt::sync(st);
    resolve_spans(p);
    return CPB_OK;
  } catch (const Error& e) {
    return fail(e.code, e.what());
  } catch (const std::bad_alloc&) {
    return fail(CPB_ERR_NOMEM, "out of host memory");
  }
}

int cpb_vpsi_dev(cpb_plan* p, const void* c0_dev, void* c2_dev, long ld, int nstate, const double* f,
                 const double* vpot_dev, int ngroups, int my_group, unsigned flags, void* stream) {
  return vpsi_dev_impl(p, c0_dev, c2_dev, ld, nstate, f, -1, vpot_dev, ngroups, my_group, flags, stream);
}

int cpb_vpsi_lsd_dev(cpb_plan* p, const void* c0_dev, void* c2_dev, long ld, int nstate, const double* f, int nsup,
                     const double* vpot_dev, int ngroups, int my_group, unsigned flags, void* stream) {
  if (nsup < 0) return fail(CPB_ERR_INVALID, "negative nsup");
  return vpsi_dev_impl(p, c0_dev, c2_dev, ld, nstate, f, nsup, vpot_dev, ngroups, my_group, flags, stream);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Dense transforms of real fields on the plan's G list (a plan created from the nhg vectors of
// the density cutoff) and the local part of vofrho.
// ---------------------------------------------------------------------------------------------
namespace {

void ensure_gbuf(cpb_plan* p, size_t elems) {
  if (elems <= p->d_gbuf_cap) return;
  rt::dfree(p->d_gbuf);
  p->d_gbuf = nullptr;
  p->d_gbuf_cap = 0;
  p->d_gbuf = (cplx*)rt::dmalloc(elems * sizeof(cplx));
  p->d_gbuf_cap = elems;
}

// fwfftn(v,.FALSE.) of nf real fields + zgthr through nzh: f (nnr1, nf) -> g (ld, nf)
void run_dense_fw(cpb_plan* p, const double* f, int nf, cplx* g, long ld, cudaStream_t st) {
  cpb_plan::WorkSpace& w = p->ws[0];
  const double* fim = nf == 2 ? f + p->nnr1() : nullptr;
  for (int xt0 = 0; xt0 < p->nxt; xt0 += p->chunk_xt) {
    const int nxc = std::min(p->chunk_xt, p->nxt - xt0);
    { Timed t(p, st, CPB_K_DENSE); p->kz->z_fwd_real(st, f, fim, w.T2, p->pd, xt0, nxc, p->half_z, nullptr, 1.0); }
    { Timed t(p, st, CPB_K_Y_FWD); p->ky->y_fwd(st, w.T2, w.T1, p->pd, 1, xt0, nxc, 1, p->half_y); }
  }
  { Timed t(p, st, CPB_K_X_FWD); p->kx->x_fwd(st, w.T1, w.G, p->pd, 1, 1, p->half_x); }
  Timed t(p, st, CPB_K_DENSE);
  auto k = k_gather_g;
  CPB_LAUNCH(k, dim3((p->ngw + 255) / 256), dim3(256), 0, st, (const cplx*)w.G, p->pd, g, nf == 2 ? g + ld : nullptr);
}

// scatter through nzh / indz + invfftn(v,.FALSE.) + REAL(): g (ld, nf) -> f (nnr1, nf)
void run_dense_inv(cpb_plan* p, const cplx* g, long ld, int nf, double* f, bool acc, cudaStream_t st) {
  cpb_plan::WorkSpace& w = p->ws[0];
  std::vector<PairHost> pairs(1);
  pairs[0].s1 = 0;
  pairs[0].s2 = nf == 2 ? 1 : -1;
  const std::vector<double> zero(1, 0.0);
  const PairDev pr = upload_pairs(p, pairs, zero, zero, st);
  double* fim = nf == 2 ? f + p->nnr1() : nullptr;
  if (!acc) rt::dzero(f, (size_t)nf * p->nnr1() * sizeof(double), st);  // pads
  { Timed t(p, st, CPB_K_X_INV); p->kx->x_inv(st, g, ld, w.T1, p->pd, pr, 1, 1, p->half_x); }
  for (int xt0 = 0; xt0 < p->nxt; xt0 += p->chunk_xt) {
    const int nxc = std::min(p->chunk_xt, p->nxt - xt0);
    { Timed t(p, st, CPB_K_Y_INV); p->ky->y_inv(st, w.T1, w.T2, p->pd, 1, xt0, nxc, 1, p->half_y); }
    { Timed t(p, st, CPB_K_DENSE); p->kz->z_inv_real(st, w.T2, f, fim, p->pd, xt0, nxc, acc, p->half_z); }
  }
}

int check_dense(cpb_plan* p, const void* a, const void* b, long ld, int nf) {
  if (!p) return fail(CPB_ERR_INVALID, "null plan");
  if (!a || !b) return fail(CPB_ERR_INVALID, "null array");
  if (nf != 1 && nf != 2) return fail(CPB_ERR_INVALID, "nfields must be 1 or 2");
  if (ld < p->ngw) return fail(CPB_ERR_INVALID, "leading dimension smaller than the plan's G count");
  return 0;
}

// enqueue the local part of vofrho on `st`; the partial sums land in p->d_red[0 .. 8*kSumBlocks)
void run_vofrho_local(cpb_plan* p, const double* rhoe, const double* scg, const cplx* eivps, const cplx* eirop,
                      cplx* rhog, cplx* vtemp, double* v, cudaStream_t st) {
  run_dense_fw(p, rhoe, 1, rhog, p->ngw, st);                      // vofrhoa_utils.mod.F90:88-95
  {
    Timed t(p, st, CPB_K_DENSE);
    auto k = k_ppener;                                             // :102 -> ppener_utils.mod.F90:23-108
    CPB_LAUNCH(k, dim3(kSumBlocks), dim3(256), 8 * 256 * sizeof(double), st, (const cplx*)rhog, scg, eivps, eirop,
               vtemp, p->ngw, p->geq0, p->d_red);
  }
  run_dense_inv(p, vtemp, p->ngw, 1, v, false, st);                // vofrhob_utils.mod.F90:155-173
}

void finish_vofrho_scalars(cpb_plan* p, double eivps0_re, double* ener) {
  if (!ener) return;
  for (int j = 0; j < 8; ++j) {
    double s = 0.0;
    for (int i = 0; i < kSumBlocks; ++i) s += p->h_red[(size_t)j * kSumBlocks + i];
    ener[j] = s;
  }
  ener[8] = p->geq0 ? eivps0_re : 0.0;  // vploc (ppener_utils.mod.F90:59-60,73)
}

}  // namespace

extern "C" {

int cpb_dense_fwfft_dev(cpb_plan* p, const double* f_dev, int nfields, void* g_dev, long ld, void* stream) {
  if (int e = check_dense(p, f_dev, g_dev, ld, nfields)) return e;
  try {
    rt::set_device(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    run_dense_fw(p, f_dev, nfields, (cplx*)g_dev, ld, st);
    rt::check_last("dense forward kernels");
    rt::sync(st);
    resolve_spans(p);
    return CPB_OK;
  } catch (const Error& e) {
    return fail(e.code, e.what());
  }
}

int cpb_dense_invfft_dev(cpb_plan* p, const void* g_dev, long ld, int nfields, double* f_dev, unsigned flags,
                         void* stream) {
  if (int e = check_dense(p, g_dev, f_dev, ld, nfields)) return e;
  try {
    rt::set_device(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    run_dense_inv(p, (const cplx*)g_dev, ld, nfields, f_dev, (flags & CPB_DENSE_ACCUMULATE) != 0, st);
    rt::check_last("dense inverse kernels");
    rt::sync(st);
    resolve_spans(p);
    return CPB_OK;
  } catch (const Error& e) {
    return fail(e.code, e.what());
  }
}

int cpb_vofrho_local_dev(cpb_plan* p, const double* rhoe_dev, const double* scg_dev, const void* eivps_dev,
                         const void* eirop_dev, void* rhog_dev, void* vtemp_dev, double* v_dev, double* ener,
                         void* stream) {
  if (!p) return fail(CPB_ERR_INVALID, "null plan");
  if (!rhoe_dev || !scg_dev || !eivps_dev || !eirop_dev || !v_dev) return fail(CPB_ERR_INVALID, "null array");
  try {
    rt::set_device(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    ensure_red(p, 8 * kSumBlocks + 2);
    ensure_gbuf(p, 2 * (size_t)p->ngw);
    cplx* rhog = rhog_dev ? (cplx*)rhog_dev : p->d_gbuf;
    cplx* vtemp = vtemp_dev ? (cplx*)vtemp_dev : p->d_gbuf + p->ngw;
    run_vofrho_local(p, rhoe_dev, scg_dev, (const cplx*)eivps_dev, (const cplx*)eirop_dev, rhog, vtemp, v_dev, st);
    rt::check_last("vofrho kernels");
    rt::d2h(p->h_red, p->d_red, (size_t)8 * kSumBlocks * sizeof(double), st);
    rt::d2h(p->h_red + 8 * kSumBlocks, eivps_dev, sizeof(cplx), st);
    rt::sync(st);
    resolve_spans(p);
    finish_vofrho_scalars(p, p->h_red[8 * kSumBlocks], ener);
    return CPB_OK;
  } catch (const Error& e) {
    return fail(e.code, e.what());
  }
}

int cpb_vofrho_local(cpb_plan* p, const double* rhoe, const double* scg, const void* eivps, const void* eirop,
                     void* rhog, void* vtemp, double* v, double* ener) {
  if (!p) return fail(CPB_ERR_INVALID, "null plan");
  if (!rhoe || !scg || !eivps || !eirop || !v) return fail(CPB_ERR_INVALID, "null array");
  try {
    rt::set_device(p->device);
    cudaStream_t st = p->s_main;
    const size_t ng = (size_t)p->ngw, nnr1 = p->nnr1();
    ensure_red(p, 8 * kSumBlocks + 2);
    ensure_gbuf(p, 4 * ng);  // rhog, vtemp, eivps, eirop
    if (!p->d_scg) p->d_scg = (double*)rt::dmalloc(ng * sizeof(double));
    if (!p->d_real || p->d_real_cols < 1) {
      rt::dfree(p->d_real);
      p->d_real = nullptr;
      p->d_real = (double*)rt::dmalloc(nnr1 * sizeof(double));
      p->d_real_cols = 1;
    }
    cplx *d_rhog = p->d_gbuf, *d_vtemp = p->d_gbuf + ng, *d_vps = p->d_gbuf + 2 * ng, *d_rop = p->d_gbuf + 3 * ng;
    rt::h2d(p->d_real, rhoe, nnr1 * sizeof(double), st);
    rt::h2d(p->d_scg, scg, ng * sizeof(double), st);
    rt::h2d(d_vps, eivps, ng * sizeof(cplx), st);
    rt::h2d(d_rop, eirop, ng * sizeof(cplx), st);
    // rho and V share the array, like the reference's rhoe (in: density, out: potential)
    run_vofrho_local(p, p->d_real, p->d_scg, d_vps, d_rop, d_rhog, d_vtemp, p->d_real, st);
    rt::check_last("vofrho kernels");
    rt::d2h(p->h_red, p->d_red, (size_t)8 * kSumBlocks * sizeof(double), st);
    rt::d2h(v, p->d_real, nnr1 * sizeof(double), st);
    if (rhog) rt::d2h(rhog, d_rhog, ng * sizeof(cplx), st);
    if (vtemp) rt::d2h(vtemp, d_vtemp, ng * sizeof(cplx), st);
    rt::sync(st);
    resolve_spans(p);
    finish_vofrho_scalars(p, ((const double*)eivps)[0], ener);
    return CPB_OK;
  } catch (const Error& e) {
    return fail(e.code, e.what());
  }
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// k-points (tkpts%tkpnt), one k-point per call: rhoofr_c's inner loops and vpsi's k-branch
// ---------------------------------------------------------------------------------------------
namespace {

struct KptScope {  // the plan is not re-entrant: the mode lives for the duration of one call
  cpb_plan* p;
  KptScope(cpb_plan* p_, const double* hgkp, const double* hgkm) : p(p_) {
    p->kpt_mode = true;
    p->kpt_hgkp = hgkp;
    p->kpt_hgkm = hgkm;
    p->psi_valid = false;
  }
  ~KptScope() {
    p->kpt_mode = false;
    p->kpt_hgkp = p->kpt_hgkm = nullptr;
  }
};

// every state of the group's block is its own transform (njump = 1, vpsi_utils.mod.F90:237-238)
std::vector<PairHost> block_singles(int nstate, int my_group, int ngroups) {
  std::vector<PairHost> out;
  const int nblk = nbr_el_in_blk(nstate, my_group, ngroups);
  for (int i = 1; i <= nblk; ++i) {
    PairHost q;
    q.s1 = get_el_in_blk(i, nstate, my_group, ngroups) - 1;
    q.s2 = -1;
    out.push_back(q);
  }
  return out;
}

int check_kpt(cpb_plan* p, const void* c0, long ld, int nstate, const double* f, const double* hgkp,
              const double* hgkm, int ngroups, int my_group) {
  if (int e = check_common(p, c0, ld, nstate, f, ngroups, my_group)) return e;
  if (ld < 2L * p->ngw) return fail(CPB_ERR_INVALID, "k-points: leading dimension smaller than ngwk = 2*ngw");
  if (!hgkp || !hgkm) return fail(CPB_ERR_INVALID, "null hgkp / hgkm");
  return 0;
}

}  // namespace

extern "C" {

int cpb_rhoofr_kpt_dev(cpb_plan* p, const void* c0_dev, long ld, int nstate, const double* f, double wk,
                       const double* hgkp_dev, const double* hgkm_dev, int ngroups, int my_group, double* rhoe_dev,
                       double* ekin, double* rsum_g, double* rsum_r, unsigned flags, void* stream) {
  if (int e = check_kpt(p, c0_dev, ld, nstate, f, hgkp_dev, hgkm_dev, ngroups, my_group)) return e;
  if (!rhoe_dev) return fail(CPB_ERR_INVALID, "null rhoe");
  try {
    rt::set_device(p->device);
    KptScope scope(p, hgkp_dev, hgkm_dev);
    cudaStream_t st = (cudaStream_t)stream;
    const cplx* c0 = (const cplx*)c0_dev;
    const int nblk = nbr_el_in_blk(nstate, my_group, ngroups);
    const int first = nblk > 0 ? get_el_in_blk(1, nstate, my_group, ngroups) - 1 : 0;
    std::vector<PairHost> pairs;
    std::vector<double> ca, cb;
    for (const PairHost& q : block_singles(nstate, my_group, ngroups)) {
      if (f[q.s1] == 0.0) continue;                      // rhoofr_c_utils.mod.F90:145
      pairs.push_back(q);
      ca.push_back(wk * f[q.s1] / p->omega);             // :165, build_density_sum(coef3, coef3, ...)
      cb.push_back(wk * f[q.s1] / p->omega);
    }
    ensure_red(p, kRedPerState * nblk + 3 * kSumBlocks);
    if (!(flags & CPB_RHO_ACCUMULATE)) rt::dzero(rhoe_dev, p->nnr1() * sizeof(double), st);   // :107
    if (nblk > 0) {
      auto k = k_kin_energy_kpt;                                                               // :117-140
      Timed t(p, st, CPB_K_KIN);
      CPB_LAUNCH(k, dim3(kKinChunks, nblk), dim3(256), 2 * 256 * sizeof(double), st, c0, ld, first, p->ngw, hgkp_dev,
                 hgkm_dev, p->d_red);
    }
    run_rhoofr(p, c0, ld, upload_pairs(p, pairs, ca, cb, st), (int)pairs.size(), (int)pairs.size(), rhoe_dev,
               rhoe_dev, nullptr, st, nullptr);
    launch_sum(p, rhoe_dev, p->nnr1(), p->d_red + kRedPerState * nblk, st);
    rt::d2h(p->h_red, p->d_red, (size_t)(kRedPerState * nblk + kSumBlocks) * sizeof(double), st);
    rt::sync(st);
    resolve_spans(p);
    double xkin = 0.0, rsum = 0.0;
    for (int i = 0; i < nblk; ++i) {
      const double fi = f[first + i];
      if (fi == 0.0) continue;                            // :118
      double sk = 0.0, sd = 0.0;
      for (int c = 0; c < kKinChunks; ++c) {
        sk += p->h_red[(size_t)kRedPerState * i + 2 * c];
        sd += p->h_red[(size_t)kRedPerState * i + 2 * c + 1];
      }
      rsum += wk * fi * sd;                               // :119
      xkin += 0.5 * wk * fi * sk;                         // :138
    }
    double sr = 0.0;
    for (int i = 0; i < kSumBlocks; ++i) sr += p->h_red[(size_t)kRedPerState * nblk + i];
    if (ekin) *ekin = xkin * p->tpiba2;                   // :182
    if (rsum_g) *rsum_g = rsum;
    if (rsum_r) *rsum_r = sr * p->omega / ((double)p->nr[0] * p->nr[1] * p->nr[2]);   // :260-261, of rhoe so far
    return CPB_OK;
  } catch (const Error& e) {
    return fail(e.code, e.what());
  } catch (const std::bad_alloc&) {
    return fail(CPB_ERR_NOMEM, "out of host memory");
  }
}

int cpb_vpsi_kpt_dev(cpb_plan* p, const void* c0_dev, void* c2_dev, long ld, int nstate, const double* f,
                     const double* hgkp_dev, const double* hgkm_dev, const double* vpot_dev, int ngroups,
                     int my_group, unsigned flags, void* stream) {
  if (int e = check_kpt(p, c0_dev, ld, nstate, f, hgkp_dev, hgkm_dev, ngroups, my_group)) return e;
  if (!c2_dev || !vpot_dev) return fail(CPB_ERR_INVALID, "null c2 or vpot");
  try {
    rt::set_device(p->device);
    KptScope scope(p, hgkp_dev, hgkm_dev);
    cudaStream_t st = (cudaStream_t)stream;
    const std::vector<PairHost> pairs = block_singles(nstate, my_group, ngroups);
    std::vector<double> fi(pairs.size()), unused(pairs.size(), 0.0);
    for (size_t i = 0; i < pairs.size(); ++i) {
      fi[i] = f[pairs[i].s1];
      if (fi[i] == 0.0) fi[i] = 2.0;                      // vpsi_utils.mod.F90:563-564
    }
    run_vpsi(p, (const cplx*)c0_dev, (cplx*)c2_dev, ld, upload_pairs(p, pairs, fi, unused, st), (int)pairs.size(),
             (int)pairs.size(), vpot_dev, vpot_dev, !(flags & CPB_VPSI_OVERWRITE), nullptr, st, nullptr);
    rt::sync(st);
    resolve_spans(p);
    return CPB_OK;
  } catch (const Error& e) {
    return fail(e.code, e.what());
  } catch (const std::bad_alloc&) {
    return fail(CPB_ERR_NOMEM, "out of host memory");
  }
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// meta-GGA (cntl%ttau): tauofr and vtaupsi - the rhoofr / vpsi pipelines run once per Cartesian
// direction on the gradient components d_k psi (dpsisc: coefficients scaled by +-gk(k,ig))
// ---------------------------------------------------------------------------------------------
namespace {

void run_x_inv_gk(cpb_plan* p, cpb_plan::WorkSpace& w, const cplx* c0, long ldc, const PairDev& prb, int nb,
                  const double* gk_dir) {
  Timed t(p, w.s, CPB_K_X_INV);
  p->kx->x_inv_gk(w.s, c0, ldc, w.T1, p->pd, prb, nb, pairs_per_group(p, nb, p->nrp / p->kx->sl, p->kx->x_inv_blocks),
                  p->half_x, gk_dir);
}

// tau0 / tau1: channel arrays (no LSD: n0 == np)
void run_tauofr(cpb_plan* p, const cplx* c0, long ldc, const PairDev& pr, int np, int n0, const double* gk, double* tau0,
                double* tau1, cudaStream_t st) {
  const std::vector<BatchSpan> batches = make_batches(p, np, n0);
  const int nbatches = (int)batches.size();
  fork_streams(p, st, nbatches);
  for (int b = 0; b < nbatches; ++b) {
    const int off = batches[b].off, nb = batches[b].n;
    double* tau = batches[b].chan ? tau1 : tau0;
    cpb_plan::WorkSpace& w = p->ws[b % p->nws];
    PairDev prb = offset_pairs(pr, off);
    for (int dir = 0; dir < 3; ++dir) {                                    // tauofr_utils.mod.F90:86-100
      run_x_inv_gk(p, w, c0, ldc, prb, nb, gk + dir);
      for (int xt0 = 0; xt0 < p->nxt; xt0 += p->chunk_xt) {
        const int nxc = std::min(p->chunk_xt, p->nxt - xt0);
        { Timed t(p, w.s, CPB_K_Y_INV); p->ky->y_inv(w.s, w.T1, w.T2, p->pd, nb, xt0, nxc, pairs_per_group(p, nb, nxc * p->nzb, p->ky->yz_blocks_per_sm), p->half_y); }
        if (b > 0 && p->nws > 1 && dir == 0) rt::stream_wait(w.s, p->ws[(b - 1) % p->nws].ev_rho);
        { Timed t(p, w.s, CPB_K_Z_RHO); launch_z_rho(p, w.s, w.T2, tau, prb, nb, xt0, nxc); }
      }
    }
    rt::event_record(w.ev_rho, w.s);
  }
  join_streams(p, st, nbatches);
  rt::check_last("tauofr kernels");
}

void run_vtaupsi(cpb_plan* p, const cplx* c0, cplx* c2, long ldc, const PairDev& pr, int np, int n0, const double* gk,
                 const double* v0, const double* v1, cudaStream_t st) {
  const std::vector<BatchSpan> batches = make_batches(p, np, n0);
  const int nbatches = (int)batches.size();
  fork_streams(p, st, nbatches);
  for (int b = 0; b < nbatches; ++b) {
    const int off = batches[b].off, nb = batches[b].n;
    const double* vtau = batches[b].chan ? v1 : v0;
    cpb_plan::WorkSpace& w = p->ws[b % p->nws];
    PairDev prb = offset_pairs(pr, off);
    for (int dir = 0; dir < 3; ++dir) {                                    // vtaupsi_utils.mod.F90:68-88
      run_x_inv_gk(p, w, c0, ldc, prb, nb, gk + dir);
      for (int xt0 = 0; xt0 < p->nxt; xt0 += p->chunk_xt) {
        const int nxc = std::min(p->chunk_xt, p->nxt - xt0);
        const int ppg_y = pairs_per_group(p, nb, nxc * p->nzb, p->ky->yz_blocks_per_sm);
        { Timed t(p, w.s, CPB_K_Y_INV); p->ky->y_inv(w.s, w.T1, w.T2, p->pd, nb, xt0, nxc, ppg_y, p->half_y); }
        { Timed t(p, w.s, CPB_K_Z_VPSI); launch_z_vpsi(p, w.s, w.T2, vtau, nb, xt0, nxc); }
        { Timed t(p, w.s, CPB_K_Y_FWD); p->ky->y_fwd(w.s, w.T2, w.T1, p->pd, nb, xt0, nxc, ppg_y, p->half_y); }
      }
      for (int o = 0; o < nb; o += p->x_sub) {
        const int ns = std::min(p->x_sub, nb - o);
        PairDev prs = offset_pairs(prb, o);
        {
          Timed t(p, w.s, CPB_K_X_FWD);
          p->kx->x_fwd(w.s, w.T1 + (size_t)o * p->t1_pair, w.G, p->pd, ns,
                       pairs_per_group(p, ns, p->nrp / p->kx->sl, p->kx->x_fwd_blocks), p->half_x);
        }
        Timed t(p, w.s, CPB_K_UNPACK);
        const int ppg = ew_ppg(p, ns, 8);
        auto k = k_unpack_tau;
        CPB_LAUNCH(k, dim3((p->ngw + 255) / 256, (ns + ppg - 1) / ppg), dim3(256), 0, w.s, (const cplx*)w.G, c2, ldc,
                   p->pd, prs, gk + dir, ns, ppg);
      }
    }
  }
  join_streams(p, st, nbatches);
  rt::check_last("vtaupsi kernels");
}

}  // namespace

extern "C" {

int cpb_tauofr_dev(cpb_plan* p, const void* c0_dev, long ld, int nstate, const double* f, int nsup,
                   const double* gk_dev, int ngroups, int my_group, double* tau_dev, unsigned flags, void* stream) {
  (void)flags;
  if (int e = check_common(p, c0_dev, ld, nstate, f, ngroups, my_group)) return e;
  if (!gk_dev || !tau_dev) return fail(CPB_ERR_INVALID, "null gk or tau");
  if (nsup > nstate) return fail(CPB_ERR_INVALID, "nsup larger than nstate");
  try {
    rt::set_device(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const bool lsd = nsup >= 0;
    const size_t nnr1 = p->nnr1();
    p->psi_valid = false;
    std::vector<PairHost> pairs;
    std::vector<double> cre, cim;
    for (const PairHost& q : block_pairs(nstate, my_group, ngroups, nsup)) {
      const double f1 = f[q.s1], f2 = q.s2 >= 0 ? f[q.s2] : 0.0;
      if (f1 == 0.0 && f2 == 0.0) continue;  // adds nothing
      pairs.push_back(q);
      // tauadd (tauofr_utils.mod.F90:147-173): coef1 = tpiba2 f(is1) / (2 omega) weights AIMAG(psi)^2,
      // coef2 (is2) weights REAL(psi)^2 - the gradient of a real state is imaginary
      cim.push_back(0.5 * p->tpiba2 * f1 / p->omega);
      cre.push_back(0.5 * p->tpiba2 * f2 / p->omega);
    }
    rt::dzero(tau_dev, (lsd ? 2 : 1) * nnr1 * sizeof(double), st);  // :80
    run_tauofr(p, (const cplx*)c0_dev, ld, upload_pairs(p, pairs, cre, cim, st), (int)pairs.size(), count_chan0(pairs),
               gk_dev, tau_dev, tau_dev + nnr1, st);
    rt::sync(st);
    resolve_spans(p);
    return CPB_OK;
  } catch (const Error& e) {
    return fail(e.code, e.what());
  } catch (const std::bad_alloc&) {
    return fail(CPB_ERR_NOMEM, "out of host memory");
  }
}

int cpb_vtaupsi_dev(cpb_plan* p, const void* c0_dev, void* c2_dev, long ld, int nstate, const double* f, int nsup,
                    const double* gk_dev, const double* vtau_dev, int ngroups, int my_group, unsigned flags,
                    void* stream) {
  (void)flags;
  if (int e = check_common(p, c0_dev, ld, nstate, f, ngroups, my_group)) return e;
  if (!c2_dev || !gk_dev || !vtau_dev) return fail(CPB_ERR_INVALID, "null c2, gk or vtau");
  if (nsup > nstate) return fail(CPB_ERR_INVALID, "nsup larger than nstate");
  try {
    rt::set_device(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    p->psi_valid = false;
    const std::vector<PairHost> pairs = block_pairs(nstate, my_group, ngroups, nsup);
    std::vector<double> fi1(pairs.size()), fi2(pairs.size());
    for (size_t i = 0; i < pairs.size(); ++i) {
      fi1[i] = 0.25 * f[pairs[i].s1] * p->tpiba2;                                   // vtaupsi_utils.mod.F90:143,154
      fi2[i] = pairs[i].s2 >= 0 ? 0.25 * f[pairs[i].s2] * p->tpiba2 : 0.0;          // :155
    }
    run_vtaupsi(p, (const cplx*)c0_dev, (cplx*)c2_dev, ld, upload_pairs(p, pairs, fi1, fi2, st), (int)pairs.size(),
                count_chan0(pairs), gk_dev, vtau_dev, vtau_dev + p->nnr1(), st);
    rt::sync(st);
    resolve_spans(p);
    return CPB_OK;
  } catch (const Error& e) {
    return fail(e.code, e.what());
  } catch (const std::bad_alloc&) {
    return fail(CPB_ERR_NOMEM, "out of host memory");
  }
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Hartree-Fock exchange: hfx_old with func1%mhfx = 1 at the Gamma point, no LSD, no
// screening (hfx_utils.mod.F90:80-965), assembled from the transforms above
// ---------------------------------------------------------------------------------------------
namespace {

// inverse transform of the packed "pairs" described by pr on plan p: gather + x + y into ws[0].T2
void hfx_xy_inv(cpb_plan* p, const cplx* g, long ldg, const PairDev& pr, int nb, cudaStream_t st) {
  cpb_plan::WorkSpace& w = p->ws[0];
  {
    Timed t(p, st, CPB_K_X_INV);
    if (p->mirror)
      p->kx->x_inv_m(st, g, ldg, w.T1, p->pd, pr, nb, nb, p->half_x, nullptr, p->geq0);
    else
      p->kx->x_inv(st, g, ldg, w.T1, p->pd, pr, nb, nb, p->half_x);
  }
  Timed t(p, st, CPB_K_Y_INV);
  p->ky->y_inv(st, w.T1, w.T2, p->pd, nb, 0, p->nxt, nb, p->half_y);
}

// forward transform of one real-space field scale * mul * (fre + i fim) on plan p into the band-ray storage ws[0].G
void hfx_fwd(cpb_plan* p, const double* fre, const double* fim, const double* mul, double scale, cudaStream_t st) {
  cpb_plan::WorkSpace& w = p->ws[0];
  { Timed t(p, st, CPB_K_DENSE); p->kz->z_fwd_real(st, fre, fim, w.T2, p->pd, 0, p->nxt, p->half_z, mul, scale); }
  { Timed t(p, st, CPB_K_Y_FWD); p->ky->y_fwd(st, w.T2, w.T1, p->pd, 1, 0, p->nxt, 1, p->half_y); }
  { Timed t(p, st, CPB_K_X_FWD); p->kx->x_fwd(st, w.T1, w.G, p->pd, 1, 1, p->half_x); }
}

struct DevBuf {  // RAII device allocation
  void* p = nullptr;
  explicit DevBuf(size_t bytes) : p(rt::dmalloc(bytes)) {}
  ~DevBuf() { rt::dfree(p); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

}  // namespace

extern "C" {

int cpb_hfx_dev(cpb_plan* pw, cpb_plan* pdn, const void* c0_dev, void* c2_dev, long ld, int nstate, const double* f,
                const double* scgx_dev, double pfl, double* ehfx, double* vhfx, unsigned flags, void* stream) {
  (void)flags;
  if (!pw || !pdn) return fail(CPB_ERR_INVALID, "null plan");
  if (int e = check_common(pw, c0_dev, ld, nstate, f, 1, 0)) return e;
  if (!c2_dev || !scgx_dev) return fail(CPB_ERR_INVALID, "null c2 or scgx");
  for (int d = 0; d < 3; ++d)
    if (pw->nr[d] != pdn->nr[d] || pw->kr[d] != pdn->kr[d])
      return fail(CPB_ERR_INVALID, "the wavefunction plan and the pair-density plan must share the mesh");
  if (pw->device != pdn->device) return fail(CPB_ERR_INVALID, "the two plans live on different devices");
  if (pw->chunk_xt != pw->nxt || pdn->chunk_xt != pdn->nxt)
    return fail(CPB_ERR_UNSUPPORTED, "cpb_hfx_dev needs unchunked y/z passes (CPB_CHUNK_XT unset)");
  try {
    rt::set_device(pw->device);
    cudaStream_t st = (cudaStream_t)stream;
    const cplx* c0 = (const cplx*)c0_dev;
    cplx* c2 = (cplx*)c2_dev;
    const size_t nnr1 = pw->nnr1();
    const int jhg = pdn->ngw;
    pw->psi_valid = false;
    std::vector<int> occ;
    for (int i = 0; i < nstate; ++i)
      if (f[i] >= 1.0e-6) occ.push_back(i);  // hfx_utils.mod.F90:474
    const int nocc = (int)occ.size();
    double e_total = 0.0;
    if (nocc > 0) {
      DevBuf rbuf((size_t)nocc * nnr1 * sizeof(double));   // psi_i(r) of every occupied state (rswfx)
      DevBuf vrbuf(2 * nnr1 * sizeof(double));              // vpotr of the two partners of a packet
      DevBuf vgbuf(2 * (size_t)jhg * sizeof(cplx));         // vpotg
      DevBuf ebuf(kSumBlocks * sizeof(double));
      DevBuf dbuf(4 * sizeof(int) + 4 * sizeof(double));
      double* R = (double*)rbuf.p;
      double* vr = (double*)vrbuf.p;
      cplx* vg = (cplx*)vgbuf.p;
      double* eacc = (double*)ebuf.p;
      rt::dzero(eacc, kSumBlocks * sizeof(double), st);
      rt::dzero(R, (size_t)nocc * nnr1 * sizeof(double), st);   // pads
      rt::dzero(vr, 2 * nnr1 * sizeof(double), st);
      // descriptors of the two packings of the pair-density plan: columns (0, 1) and (0, -)
      {
        const int hs[4] = {0, 0, 1, -1};
        const double hz[4] = {0, 0, 0, 0};
        int* di = (int*)dbuf.p;
        double* dd = (double*)(di + 4);
        rt::h2d(di, hs, sizeof hs, st);
        rt::h2d(dd, hz, sizeof hz, st);
        rt::sync(st);  // hs / hz live on this stack frame
      }
      int* di = (int*)dbuf.p;
      double* dd = (double*)(di + 4);
      PairDev two, one;
      two.st1 = di;
      two.st2 = di + 2;
      two.ca = dd;
      two.cb = dd + 2;
      one = offset_pairs(two, 1);
      // ---- real-space states, two per transform (set_psi_2_states_g + invfftn, :296-318)
      {
        std::vector<PairHost> pairs;
        for (int i = 0; i < nocc; i += 2) {
          PairHost q;
          q.s1 = occ[i];
          q.s2 = i + 1 < nocc ? occ[i + 1] : -1;
          pairs.push_back(q);
        }
        const std::vector<double> z(pairs.size(), 0.0);
        const PairDev pr = upload_pairs(pw, pairs, z, z, st);
        cpb_plan::WorkSpace& w = pw->ws[0];
        for (int off = 0; off < (int)pairs.size(); off += pw->max_batch) {
          const int nb = std::min(pw->max_batch, (int)pairs.size() - off);
          hfx_xy_inv(pw, c0, ld, offset_pairs(pr, off), nb, st);
          for (int i = 0; i < nb; ++i) {
            const int k = 2 * (off + i);
            Timed t(pw, st, CPB_K_DENSE);
            pw->kz->z_inv_real(st, w.T2 + (size_t)i * pw->t2_pair, R + (size_t)k * nnr1,
                               k + 1 < nocc ? R + (size_t)(k + 1) * nnr1 : nullptr, pw->pd, 0, pw->nxt, false, pw->half_z);
          }
        }
      }
      // ---- pair terms: state ia with itself (hfxaa = hfxab with half the prefactor) and with every later
      // occupied state, two partners per pair-density transform (hfxab2, :704-745)
      for (int a = 0; a < nocc; ++a) {
        const int ia = occ[a];
        for (int k = a; k < nocc; k += 2) {
          const int b1 = k, b2 = k + 1 < nocc ? k + 1 : -1;
          const double pf1 = pfl * f[ia] * f[occ[b1]] * (b1 == a ? 0.5 : 1.0);
          const double pf2 = b2 >= 0 ? pfl * f[ia] * f[occ[b2]] : 0.0;
          const double* ra = R + (size_t)a * nnr1;
          hfx_fwd(pdn, R + (size_t)b1 * nnr1, b2 >= 0 ? R + (size_t)b2 * nnr1 : nullptr, ra, 1.0 / pw->omega, st);
          {
            Timed t(pdn, st, CPB_K_DENSE);
            auto kc = k_hfx_coulomb;
            CPB_LAUNCH(kc, dim3(kSumBlocks), dim3(256), 256 * sizeof(double), st, (const cplx*)pdn->ws[0].G, pdn->pd, scgx_dev,
                       pf1, pf2, b2 >= 0 ? 1 : 0, pdn->geq0, vg, vg + jhg, eacc);
          }
          hfx_xy_inv(pdn, vg, jhg, b2 >= 0 ? two : one, 1, st);
          {
            Timed t(pdn, st, CPB_K_DENSE);
            pdn->kz->z_inv_real(st, pdn->ws[0].T2, vr, b2 >= 0 ? vr + nnr1 : nullptr, pdn->pd, 0, pdn->nxt, false, pdn->half_z);
          }
          for (int j = 0; j < 2; ++j) {
            const int b = j == 0 ? b1 : b2;
            if (b < 0) continue;
            hfx_fwd(pw, ra, R + (size_t)b * nnr1, vr + (size_t)j * nnr1, 1.0, st);
            Timed t(pw, st, CPB_K_UNPACK);
            auto ka = k_hfx_acc;
            CPB_LAUNCH(ka, dim3((pw->ngw + 255) / 256), dim3(256), 0, st, (const cplx*)pw->ws[0].G, pw->pd,
                       c2 + (size_t)ia * ld, c2 + (size_t)occ[b] * ld, b == a ? 1 : 0);
          }
        }
      }
      rt::check_last("hfx kernels");
      ensure_red(pw, std::max(kSumBlocks, kKinChunks * nstate));
      rt::d2h(pw->h_red, eacc, kSumBlocks * sizeof(double), st);
      rt::sync(st);
      for (int i = 0; i < kSumBlocks; ++i) e_total += pw->h_red[i];
    }
    if (ehfx) *ehfx = e_total * pw->omega;  // :905
    if (vhfx) {
      double v = 0.0;
      if (nstate > 0) {
        ensure_red(pw, std::max(kSumBlocks, kKinChunks * nstate));
        auto kd = k_dotp;
        Timed t(pw, st, CPB_K_KIN);
        CPB_LAUNCH(kd, dim3(kKinChunks, nstate), dim3(256), 256 * sizeof(double), st, c0, (const cplx*)c2, ld, pw->ngw,
                   pw->geq0, pw->d_red);
        rt::d2h(pw->h_red, pw->d_red, (size_t)kKinChunks * nstate * sizeof(double), st);
        rt::sync(st);
        for (int i = 0; i < kKinChunks * nstate; ++i) v += pw->h_red[i];  // :907-909
      }
      *vhfx = v;
    }
    rt::sync(st);
    resolve_spans(pw);
    resolve_spans(pdn);
    return CPB_OK;
  } catch (const Error& e) {
    return fail(e.code, e.what());
  } catch (const std::bad_alloc&) {
    return fail(CPB_ERR_NOMEM, "out of host memory");
  }
}

}  // extern "C"

#include "host_api.inc"
