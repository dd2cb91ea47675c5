// cpb_peer_*: the cross-group collectives of the hot path as hand-written kernels over NVLink peer
// memory (one process per GPU, CUDA IPC mappings of every rank's segment):
//   cpb_peer_allreduce_f64   cp_grp_redist(rhoe) = mp_sum over cp_inter_grp
//                            (rhoofr_utils.mod.F90:457-461, cp_grp_utils.mod.F90:98-120)
//   cpb_peer_bcast_f64       the once-per-step distribution of V(r)
// Two-shot all-reduce: rank r sums slice r of all segments in rank order (deterministic: every rank
// ends with bit-identical data) and writes the result into every segment; the slices are disjoint,
// so the operation is in place.  Broadcast: rank r pulls slice r from the source, then the other
// slices from their owners, so the source's outbound link carries the array once.  Ranks meet at
// flag barriers kept in the segments (release/acquire at system scope, monotonically increasing
// epoch).  A barrier that does not complete within the segment's timeout (cpb_peer_set_timeout_ms,
// default 20 s) raises the error word of EVERY segment instead of hanging the device; once the own
// error word is set the collective kernels leave the data untouched (no half-summed arrays), and
// cpb_peer_check - mandatory after the collectives of a step - reports it on every rank.
//   cpb_peer_allgather_f64   cp_grp_redist(C2_vpsi) (vpsi_utils.mod.F90:708-712): the reference sums
//                            zero-padded blocks, i.e. an all-gather of the owned state blocks
//   cpb_peer_allreduce_scalars  the group-partial ekin / rsum (2-3 doubles) summed in rank order
#include "../../include/cpb200.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "rt.h"

using namespace cpb;

namespace cpb {
constexpr int kPeerMax = 16;
constexpr int kPeerChannels = 6;
constexpr int kPeerScalars = 8;  // doubles per rank in the scalar mailbox
// control block behind the payload of every segment:
//   uint32 flag[kPeerChannels][kPeerMax]   barrier words (written by the peers)
//   uint32 err, pad[15]                    sticky error word
//   double mail[2][kPeerMax][kPeerScalars] scalar mailboxes (two, alternating per call)
constexpr size_t kErrWord = (size_t)kPeerChannels * kPeerMax;
constexpr size_t kMailOff = (kErrWord + 16) * sizeof(uint32_t);  // bytes from the control block
constexpr size_t kCtlBytes = kMailOff + 2 * kPeerMax * kPeerScalars * sizeof(double);
struct PeerPtrs {
  double* buf[kPeerMax];
  uint32_t* flag[kPeerMax];  // control block of every rank
};
CPB_HD double* peer_mail(const PeerPtrs& pp, int q, int which) {
  return reinterpret_cast<double*>(reinterpret_cast<char*>(pp.flag[q]) + kMailOff) +
         (size_t)which * kPeerMax * kPeerScalars;
}
CPB_HD bool peer_failed(const PeerPtrs& pp, int rank) {
  return *reinterpret_cast<volatile uint32_t*>(pp.flag[rank] + kErrWord) != 0u;
}

#if defined(CPB_EMULATE)
inline void st_release_sys(uint32_t* p, uint32_t v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline uint32_t ld_acquire_sys(const uint32_t* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void fence_sys() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
#else
CPB_D void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
CPB_D uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
CPB_D void fence_sys() { __threadfence_system(); }
#endif

// One block, >= world threads.  Thread q tells rank q "rank `rank` reached epoch" and then waits for
// rank q's word in the own segment.  (epoch - seen) as a signed difference tolerates wrap-around.
// timeout: clock64 ticks (product) / spin count (simulator); on expiry the error word of every
// segment is raised, so that the late rank's own kernels skip their work too.
CPB_GLOBAL k_peer_barrier(PeerPtrs pp, int rank, int world, int channel, uint32_t epoch, long long timeout) {
  const int q = threadIdx.x;
  fence_sys();
  if (q < world) st_release_sys(&pp.flag[q][channel * kPeerMax + rank], epoch);
  __syncthreads();
  if (q < world) {
    const uint32_t* mine = &pp.flag[rank][channel * kPeerMax + q];
    bool expired = false;
#if defined(CPB_EMULATE)
    long long spins = 0;
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
      if (++spins > timeout) {
        expired = true;
        break;
      }
    }
#else
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
      if (clock64() - t0 > timeout) {  // a rank is missing
        expired = true;
        break;
      }
      __nanosleep(64);
    }
#endif
    if (expired)
      for (int r = 0; r < world; ++r) st_release_sys(&pp.flag[r][kErrWord], 1u);
  }
}

// slice of rank r in units of double2: [lo, hi)
CPB_HD void peer_slice(size_t n2, int world, int r, size_t& lo, size_t& hi) {
  const size_t per = (n2 + world - 1) / world;
  lo = per * r < n2 ? per * r : n2;
  hi = lo + per < n2 ? lo + per : n2;
}

// reduce-scatter + all-gather of the own slice; `off`/`n` in doubles, both even (16-byte vectors).
template <int W>
CPB_GLOBAL k_peer_allreduce(PeerPtrs pp, int rank, size_t off, size_t n) {
  if (peer_failed(pp, rank)) return;  // a barrier timed out: leave the data alone
  size_t lo, hi;
  peer_slice(n / 2, W, rank, lo, hi);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
    double2 v[W];
#pragma unroll
    for (int q = 0; q < W; ++q) v[q] = reinterpret_cast<const double2*>(pp.buf[q] + off)[i];
    double2 s = v[0];
#pragma unroll
    for (int q = 1; q < W; ++q) {  // fixed rank order: bit-identical on every rank
      s.x += v[q].x;
      s.y += v[q].y;
    }
#pragma unroll
    for (int q = 0; q < W; ++q) reinterpret_cast<double2*>(pp.buf[q] + off)[i] = s;
  }
}

// same for a world size without a template instantiation (9-15 ranks)
CPB_GLOBAL k_peer_allreduce_any(PeerPtrs pp, int rank, int world, size_t off, size_t n) {
  if (peer_failed(pp, rank)) return;
  size_t lo, hi;
  peer_slice(n / 2, world, rank, lo, hi);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
    double2 s = reinterpret_cast<const double2*>(pp.buf[0] + off)[i];
    for (int q = 1; q < world; ++q) {
      const double2 v = reinterpret_cast<const double2*>(pp.buf[q] + off)[i];
      s.x += v.x;
      s.y += v.y;
    }
    for (int q = 0; q < world; ++q) reinterpret_cast<double2*>(pp.buf[q] + off)[i] = s;
  }
}

// all-gather in place: block q (doubles [boff[q], boff[q+1]) past `off`) is valid in rank q's segment;
// every rank pulls the other blocks from their owners, owners staggered over the ranks
struct PeerBlocks {
  size_t boff[kPeerMax + 1];
};
CPB_GLOBAL k_peer_allgather(PeerPtrs pp, int rank, int world, size_t off, PeerBlocks pb) {
  if (peer_failed(pp, rank)) return;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  double2* mine = reinterpret_cast<double2*>(pp.buf[rank] + off);
  for (int dq = 1; dq < world; ++dq) {
    const int q = (rank + dq) % world;
    const double2* s = reinterpret_cast<const double2*>(pp.buf[q] + off);
    for (size_t i = pb.boff[q] / 2 + t0; i < pb.boff[q + 1] / 2; i += stride) mine[i] = s[i];
  }
}

// scalar all-reduce, step 1: rank writes its n values into slot `rank` of mailbox `which` of every
// segment; step 2 (after a barrier): sum of the slots in rank order -> out (bit-identical everywhere)
CPB_GLOBAL k_peer_scalars(PeerPtrs pp, int rank, int world, int which, int step, const double* vals, double* out,
                          int n) {
  if (peer_failed(pp, rank)) return;
  const int t = threadIdx.x;
  if (step == 1) {
    if (t < world * n) {
      const int q = t / n, j = t % n;
      peer_mail(pp, q, which)[rank * kPeerScalars + j] = vals[j];
    }
  } else if (t < n) {
    const double* m = peer_mail(pp, rank, which);
    double s = m[t];
    for (int q = 1; q < world; ++q) s += m[q * kPeerScalars + t];
    out[t] = s;
  }
}

// broadcast, phase 1: own slice from the source; phase 2: every other slice from its owner
CPB_GLOBAL k_peer_bcast(PeerPtrs pp, int rank, int world, int src, int phase, size_t off, size_t n) {
  if (peer_failed(pp, rank)) return;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  double2* mine = reinterpret_cast<double2*>(pp.buf[rank] + off);
  if (phase == 1) {
    if (rank == src) return;
    size_t lo, hi;
    peer_slice(n / 2, world, rank, lo, hi);
    const double2* s = reinterpret_cast<const double2*>(pp.buf[src] + off);
    for (size_t i = lo + t0; i < hi; i += stride) mine[i] = s[i];
  } else {
    if (rank == src) return;
    for (int dq = 1; dq < world; ++dq) {
      const int q = (rank + dq) % world;  // stagger the owners over the ranks
      size_t lo, hi;
      peer_slice(n / 2, world, q, lo, hi);
      const double2* s = reinterpret_cast<const double2*>(pp.buf[q] + off);
      for (size_t i = lo + t0; i < hi; i += stride) mine[i] = s[i];
    }
  }
}
}  // namespace cpb

struct cpb_peer {
  int device = 0, rank = 0, world = 1;
  size_t bytes = 0;       // payload bytes of every segment
  char* local = nullptr;  // own segment: payload, then the control block
  void* mapped[kPeerMax] = {nullptr};
  bool connected = false;
  uint32_t epoch = 0;
  uint32_t mail_calls = 0;
  PeerPtrs pp;
  int n_sm = 148;
  long long timeout = 0;      // barrier timeout in clock64 ticks (simulator: spins)
  uint32_t* h_err = nullptr;  // pinned
  double* h_scal = nullptr;   // pinned, 2 * kPeerScalars
  double* d_scal = nullptr;   // 2 * kPeerScalars
};

namespace {
thread_local std::string g_peer_error;
int pfail(int code, const std::string& m) {
  g_peer_error = m;
  return code;
}
size_t payload_pad(size_t bytes) { return (bytes + 255) / 256 * 256; }

long long timeout_ticks(double ms) {
#if defined(CPB_EMULATE)
  return (long long)(ms * 2.0e5);  // spins of the simulator's wait loop
#else
  return (long long)(ms * 1.9e6);  // clock64 ticks at ~1.9 GHz
#endif
}

void barrier(cpb_peer* p, int channel, cudaStream_t st) {
  p->epoch += 1;
  auto k = k_peer_barrier;
  CPB_LAUNCH(k, dim3(1), dim3(32), 0, st, p->pp, p->rank, p->world, channel, p->epoch, p->timeout);
}

int check_range(cpb_peer* p, size_t off, size_t n) {
  if (!p) return pfail(CPB_ERR_INVALID, "null peer segment");
  if (!p->connected) return pfail(CPB_ERR_INVALID, "peer segment not connected (cpb_peer_connect)");
  if ((off & 1) || (n & 1)) return pfail(CPB_ERR_INVALID, "offset and count must be even (16-byte vectors)");
  if ((off + n) * sizeof(double) > p->bytes) return pfail(CPB_ERR_INVALID, "range outside the segment");
  return 0;
}

// the collectives only enqueue work; cpb_peer_check synchronises and reads the error word
int finish(cpb_peer*, cudaStream_t, const char* what) {
  rt::check_last(what);
  return CPB_OK;
}
int check_now(cpb_peer* p, cudaStream_t st, const char* what) {
  rt::d2h(p->h_err, p->local + payload_pad(p->bytes) + kErrWord * sizeof(uint32_t), sizeof(uint32_t), st);
  rt::sync(st);
  if (*p->h_err)
    return pfail(CPB_ERR_CUDA, std::string(what) +
                                   ": a rank did not reach a barrier within the timeout; the collectives after it "
                                   "were skipped and the segment is unusable");
  return CPB_OK;
}
}  // namespace

extern "C" {

const char* cpb_peer_last_error(void) { return g_peer_error.c_str(); }

int cpb_peer_create(cpb_peer** out, int device, int rank, int world, size_t bytes, void* handle_out) {
  if (!out || !handle_out) return pfail(CPB_ERR_INVALID, "null argument");
  *out = nullptr;
  if (world < 1 || world > kPeerMax || rank < 0 || rank >= world) return pfail(CPB_ERR_INVALID, "bad (rank, world)");
  if (bytes == 0 || (bytes & 15)) return pfail(CPB_ERR_INVALID, "segment size must be a positive multiple of 16");
  cpb_peer* p = nullptr;
  try {
    p = new cpb_peer();
    p->device = device;
    p->rank = rank;
    p->world = world;
    p->bytes = bytes;
    double ms = 20000.0;
    if (const char* e = std::getenv("CPB_PEER_TIMEOUT_MS")) ms = std::max(1.0, std::atof(e));
    p->timeout = timeout_ticks(ms);
    rt::set_device(device);
    p->n_sm = rt::sm_count(device);
    const size_t total = payload_pad(bytes) + kCtlBytes;
    p->local = (char*)rt::dmalloc(total);
    rt::dzero(p->local, total, 0);
    rt::sync(0);
    p->h_err = (uint32_t*)rt::hmalloc_pinned(sizeof(uint32_t));
    *p->h_err = 0;
    p->h_scal = (double*)rt::hmalloc_pinned(2 * kPeerScalars * sizeof(double));
    p->d_scal = (double*)rt::dmalloc(2 * kPeerScalars * sizeof(double));
    std::memset(handle_out, 0, CPB_PEER_HANDLE_BYTES);
#if defined(CPB_EMULATE)
    std::memcpy(handle_out, &p->local, sizeof(char*));  // simulator: ranks are threads of one process
#else
    cudaIpcMemHandle_t h;
    static_assert(sizeof(h) <= CPB_PEER_HANDLE_BYTES, "handle size");
    rt::ck(cudaIpcGetMemHandle(&h, p->local), "cudaIpcGetMemHandle");
    std::memcpy(handle_out, &h, sizeof(h));
#endif
    *out = p;
    return CPB_OK;
  } catch (const Error& e) {
    if (p) {
      rt::dfree(p->local);
      rt::dfree(p->d_scal);
      rt::hfree_pinned(p->h_err);
      rt::hfree_pinned(p->h_scal);
      delete p;
    }
    return pfail(e.code, e.what());
  }
}

int cpb_peer_set_timeout_ms(cpb_peer* p, double ms) {
  if (!p || !(ms > 0.0)) return pfail(CPB_ERR_INVALID, "null segment or non-positive timeout");
  p->timeout = timeout_ticks(ms);
  return CPB_OK;
}

int cpb_peer_connect(cpb_peer* p, const void* all_handles) {
  if (!p || !all_handles) return pfail(CPB_ERR_INVALID, "null argument");
  try {
    rt::set_device(p->device);
    const char* hs = (const char*)all_handles;
    for (int q = 0; q < p->world; ++q) {
      char* base = nullptr;
      if (q == p->rank) {
        base = p->local;
      } else {
#if defined(CPB_EMULATE)
        std::memcpy(&base, hs + (size_t)q * CPB_PEER_HANDLE_BYTES, sizeof(char*));
#else
        cudaIpcMemHandle_t h;
        std::memcpy(&h, hs + (size_t)q * CPB_PEER_HANDLE_BYTES, sizeof(h));
        void* m = nullptr;
        rt::ck(cudaIpcOpenMemHandle(&m, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
        p->mapped[q] = m;
        base = (char*)m;
#endif
      }
      p->pp.buf[q] = (double*)base;
      p->pp.flag[q] = (uint32_t*)(base + payload_pad(p->bytes));
    }
    for (int q = p->world; q < kPeerMax; ++q) {
      p->pp.buf[q] = nullptr;
      p->pp.flag[q] = nullptr;
    }
    p->connected = true;
    return CPB_OK;
  } catch (const Error& e) {
    return pfail(e.code, e.what());
  }
}

void* cpb_peer_local_ptr(cpb_peer* p) { return p ? (void*)p->local : nullptr; }

int cpb_peer_barrier(cpb_peer* p, void* stream) {
  if (!p || !p->connected) return pfail(CPB_ERR_INVALID, "peer segment not connected");
  try {
    rt::set_device(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    barrier(p, 0, st);
    rt::check_last("cpb_peer_barrier");
    return check_now(p, st, "cpb_peer_barrier");
  } catch (const Error& e) {
    return pfail(e.code, e.what());
  }
}

int cpb_peer_allreduce_f64(cpb_peer* p, size_t offset, size_t n, void* stream) {
  if (int e = check_range(p, offset, n)) return e;
  try {
    rt::set_device(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid(4 * p->n_sm), block(256);
    barrier(p, 1, st);  // every rank's partial array is complete
    switch (p->world) {
#define CPB_PEER_CASE(W)                                                            \
  case W: {                                                                         \
    auto k = k_peer_allreduce<W>;                                                   \
    CPB_LAUNCH(k, grid, block, 0, st, p->pp, p->rank, offset, n);                   \
  } break;
      CPB_PEER_CASE(1) CPB_PEER_CASE(2) CPB_PEER_CASE(3) CPB_PEER_CASE(4) CPB_PEER_CASE(5) CPB_PEER_CASE(6)
      CPB_PEER_CASE(7) CPB_PEER_CASE(8) CPB_PEER_CASE(16)
#undef CPB_PEER_CASE
      default: {
        auto k = k_peer_allreduce_any;
        CPB_LAUNCH(k, grid, block, 0, st, p->pp, p->rank, p->world, offset, n);
      }
    }
    barrier(p, 2, st);  // every slice has been written everywhere
    return finish(p, st, "cpb_peer_allreduce_f64");
  } catch (const Error& e) {
    return pfail(e.code, e.what());
  }
}

int cpb_peer_bcast_f64(cpb_peer* p, size_t offset, size_t n, int src, void* stream) {
  if (int e = check_range(p, offset, n)) return e;
  if (src < 0 || src >= p->world) return pfail(CPB_ERR_INVALID, "bad source rank");
  try {
    rt::set_device(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid(4 * p->n_sm), block(256);
    auto k = k_peer_bcast;
    barrier(p, 1, st);  // the source's array is complete, nobody still reads the old one
    CPB_LAUNCH(k, grid, block, 0, st, p->pp, p->rank, p->world, src, 1, offset, n);
    barrier(p, 2, st);  // every owner holds its slice
    CPB_LAUNCH(k, grid, block, 0, st, p->pp, p->rank, p->world, src, 2, offset, n);
    barrier(p, 3, st);  // nobody still reads a peer's slice
    return finish(p, st, "cpb_peer_bcast_f64");
  } catch (const Error& e) {
    return pfail(e.code, e.what());
  }
}

int cpb_peer_allgather_f64(cpb_peer* p, size_t offset, const size_t* counts, void* stream) {
  if (!counts) return pfail(CPB_ERR_INVALID, "null counts");
  if (!p) return pfail(CPB_ERR_INVALID, "null peer segment");
  PeerBlocks pb;
  pb.boff[0] = 0;
  for (int q = 0; q < p->world; ++q) {
    if (counts[q] & 1) return pfail(CPB_ERR_INVALID, "block sizes must be even (16-byte vectors)");
    pb.boff[q + 1] = pb.boff[q] + counts[q];
  }
  if (int e = check_range(p, offset, pb.boff[p->world])) return e;
  try {
    rt::set_device(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    auto k = k_peer_allgather;
    barrier(p, 4, st);  // every owner's block is complete
    CPB_LAUNCH(k, dim3(4 * p->n_sm), dim3(256), 0, st, p->pp, p->rank, p->world, offset, pb);
    barrier(p, 5, st);  // nobody still reads a peer's block
    return finish(p, st, "cpb_peer_allgather_f64");
  } catch (const Error& e) {
    return pfail(e.code, e.what());
  }
}

// part_1d.mod.F90:22-57 (same formulas as cpb_part_1d_*)
static size_t blk_count(int n, int proc, int nproc) { return (size_t)(n / nproc + (proc < n % nproc ? 1 : 0)); }

int cpb_peer_redist_c2(cpb_peer* p, size_t offset, long ld, int nstate, void* stream) {
  if (!p) return pfail(CPB_ERR_INVALID, "null peer segment");
  if (ld <= 0 || nstate < 0) return pfail(CPB_ERR_INVALID, "bad (ld, nstate)");
  size_t counts[kPeerMax];
  for (int q = 0; q < p->world; ++q) counts[q] = 2 * (size_t)ld * blk_count(nstate, q, p->world);
  return cpb_peer_allgather_f64(p, offset, counts, stream);
}

int cpb_peer_allreduce_scalars(cpb_peer* p, double* vals, int n, void* stream) {
  if (!p || !p->connected) return pfail(CPB_ERR_INVALID, "peer segment not connected");
  if (!vals || n < 1 || n > kPeerScalars) return pfail(CPB_ERR_INVALID, "1..8 scalars");
  try {
    rt::set_device(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int which = (int)(p->mail_calls++ & 1u);
    for (int j = 0; j < n; ++j) p->h_scal[j] = vals[j];
    rt::h2d(p->d_scal, p->h_scal, n * sizeof(double), st);
    auto k = k_peer_scalars;
    CPB_LAUNCH(k, dim3(1), dim3(kPeerMax * kPeerScalars), 0, st, p->pp, p->rank, p->world, which, 1,
               (const double*)p->d_scal, p->d_scal + kPeerScalars, n);
    barrier(p, 0, st);  // every rank's values have landed (and everybody finished the call before last)
    CPB_LAUNCH(k, dim3(1), dim3(kPeerMax * kPeerScalars), 0, st, p->pp, p->rank, p->world, which, 2,
               (const double*)p->d_scal, p->d_scal + kPeerScalars, n);
    rt::check_last("cpb_peer_allreduce_scalars");
    rt::d2h(p->h_scal + kPeerScalars, p->d_scal + kPeerScalars, n * sizeof(double), st);
    if (int e = check_now(p, st, "cpb_peer_allreduce_scalars")) return e;
    for (int j = 0; j < n; ++j) vals[j] = p->h_scal[kPeerScalars + j];
    return CPB_OK;
  } catch (const Error& e) {
    return pfail(e.code, e.what());
  }
}

int cpb_peer_check(cpb_peer* p, void* stream) {
  if (!p || !p->connected) return pfail(CPB_ERR_INVALID, "peer segment not connected");
  try {
    rt::set_device(p->device);
    return check_now(p, (cudaStream_t)stream, "cpb_peer_check");
  } catch (const Error& e) {
    return pfail(e.code, e.what());
  }
}

int cpb_peer_destroy(cpb_peer* p) {
  if (!p) return CPB_OK;
  try {
    rt::set_device(p->device);
  } catch (...) {
  }
#if !defined(CPB_EMULATE)
  for (int q = 0; q < kPeerMax; ++q)
    if (p->mapped[q]) cudaIpcCloseMemHandle(p->mapped[q]);
#endif
  rt::dfree(p->local);
  rt::dfree(p->d_scal);
  rt::hfree_pinned(p->h_err);
  rt::hfree_pinned(p->h_scal);
  delete p;
  return CPB_OK;
}

}  // extern "C"
