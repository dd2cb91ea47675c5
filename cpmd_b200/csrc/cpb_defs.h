// Common definitions for the cpb200 library (B200 / sm_100a vpsi + rhoofr path).
//
// The same kernel sources compile in two ways:
//   * nvcc, sm_100a  -> the product (libcpb200.so).
//   * g++ -DCPB_EMULATE -> tests/emu/libcpb200_emu.so, a *functional simulator* of the CUDA
//     kernels (one fiber per CUDA thread, __syncthreads = fiber barrier).  It exists because the
//     authoring container has no GPU: it lets the index arithmetic of every kernel be checked
//     against the oracle on the CPU.  It is test infrastructure only; the Python package never
//     loads it (cpmd_b200/lib.py loads libcpb200.so or raises).
#pragma once
#include <cstddef>
#include <cstdint>
#include <type_traits>

#if defined(CPB_EMULATE)
#include "emu_cuda.h"
#define CPB_HD inline
#define CPB_D inline
#define CPB_GLOBAL static void
#define CPB_LAUNCH_BOUNDS(...)
#define CPB_RESTRICT
#define CPB_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(::emu::dyn_smem())
#define CPB_SHARED static thread_local
#define CPB_LAUNCH(kern, grid, block, smem, stream, ...) \
  ::emu::launch((grid), (block), (smem), [=]() { kern(__VA_ARGS__); })
#define CPB_LAUNCH_PDL(cls, kern, grid, block, smem, stream, ...) CPB_LAUNCH(kern, grid, block, smem, stream, __VA_ARGS__)
#else
#include <cuda_runtime.h>
#define CPB_HD __host__ __device__ __forceinline__
#define CPB_D __device__ __forceinline__
#define CPB_GLOBAL __global__ void
#define CPB_LAUNCH_BOUNDS(...) __launch_bounds__(__VA_ARGS__)
#define CPB_RESTRICT __restrict__
#define CPB_DYN_SMEM(type, name)                                   \
  extern __shared__ __align__(16) unsigned char cpb_dyn_smem_[];   \
  type* name = reinterpret_cast<type*>(cpb_dyn_smem_)
#define CPB_SHARED __shared__
#define CPB_LAUNCH(kern, grid, block, smem, stream, ...) \
  kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
// launch with programmatic stream serialisation (see pdl_wait below); only for kernels that call pdl_wait()
// cls: kernel class 0..6 (x_inv, y_inv, z_rho, z_vpsi, y_fwd, x_fwd, dense z), bit of the CPB_PDL mask
#define CPB_LAUNCH_PDL(cls, kern, grid, block, smem, stream, ...) \
  ::cpb::launch_pdl((cls), (kern), (grid), (block), (smem), (stream), __VA_ARGS__)
#endif

namespace cpb {

typedef double2 cplx;

template <int I>
using IC = std::integral_constant<int, I>;

// Compile-time loop: f(IC<B>{}), f(IC<B+1>{}), ... f(IC<E-1>{}).  Every index is a constant
// expression inside f, so arrays indexed with it live in registers.
template <int B, int E, class F>
CPB_HD void static_for(F&& f) {
  if constexpr (B < E) {
    f(IC<B>{});
    static_for<B + 1, E>(f);
  }
}

CPB_HD cplx mk(double x, double y) {
  cplx r;
  r.x = x;
  r.y = y;
  return r;
}
CPB_HD cplx cadd(cplx a, cplx b) { return mk(a.x + b.x, a.y + b.y); }
CPB_HD cplx csub(cplx a, cplx b) { return mk(a.x - b.x, a.y - b.y); }
CPB_HD cplx cmul(cplx a, cplx b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// a * conj(b)
CPB_HD cplx cmulc(cplx a, cplx b) { return mk(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
CPB_HD cplx cscale(cplx a, double s) { return mk(a.x * s, a.y * s); }

// Cache-policy helpers.  Streaming accesses (the big intermediates that are written once and read
// once) are marked evict-first so that they do not push the gather-heavy plane-wave columns, which
// are staged in L2 by l2_prefetch(), out of the 126 MB L2.
#if defined(CPB_EMULATE)
inline void st_stream(cplx* p, cplx v) { *p = v; }
inline cplx ld_stream(const cplx* p) { return *p; }
inline void l2_prefetch(const void*, unsigned) {}
#else
CPB_D void st_stream(cplx* p, cplx v) { __stcs(p, v); }
CPB_D cplx ld_stream(const cplx* p) { return __ldcs(p); }
// TMA bulk prefetch of `bytes` (multiple of 16, 16-byte aligned address) into L2
CPB_D void l2_prefetch(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
#endif

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch.  The kernels of a batch form a chain on one stream, each consuming what the
// previous one wrote; between two of them the SMs drain (last wave of the producer), the front end launches
// the consumer and its blocks run their prologue (twiddle table, index tables, barrier set-up) - a few
// microseconds, several hundred times per CP step.  Kernels launched with CPB_LAUNCH_PDL may become resident
// as soon as every block of the preceding kernel has called pdl_trigger() (first statement of every hot-path
// kernel), i.e. while its last wave is still running; they do the part of their prologue that only reads
// plan constants and then pdl_wait(): that returns when the preceding kernel has completed and its writes are
// visible.  Everything produced by earlier work on the stream is touched after pdl_wait() only.  Both are
// no-ops when the kernel was launched without the attribute (CPB_PDL=0) and in the simulator build.
// ---------------------------------------------------------------------------------------------
#if defined(CPB_EMULATE)
inline void pdl_wait() {}
inline void pdl_trigger() {}
#else
CPB_D void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
CPB_D void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
unsigned pdl_mask();  // cpb200.cu: kernel classes that may start early (CPB_PDL environment switch)
template <class... KArgs, class... Args>
inline void launch_pdl(int cls, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = ((pdl_mask() >> cls) & 1u) ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

// ---------------------------------------------------------------------------------------------
// Bulk asynchronous copies global -> shared (TMA engine, 1-D: cp.async.bulk, SASS UBLKCP) with
// mbarrier transaction counting.  One elected thread arms the barrier with the byte count and
// issues the copy; every consumer thread waits on the barrier's phase parity before it reads the
// tile with ordinary shared-memory loads.  `bytes` and both addresses are multiples of 16.
// In the simulator build the copy happens at issue time and the waits are no-ops, which is a legal
// schedule: the source was written by an earlier kernel, and the destination stage is only
// re-armed after a block barrier that follows its last read.
// ---------------------------------------------------------------------------------------------
#if defined(CPB_EMULATE)
inline void mbar_init(uint64_t* bar, unsigned) { *bar = 0; }
inline void mbar_fence_init() {}
inline void mbar_expect_tx(uint64_t*, unsigned) {}
inline void mbar_wait(uint64_t*, unsigned) {}
inline void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t*) {
  const char* s = static_cast<const char*>(src);
  char* d = static_cast<char*>(dst);
  for (unsigned i = 0; i < bytes; ++i) d[i] = s[i];
}
#else
CPB_D unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
CPB_D void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
CPB_D void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
CPB_D void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
CPB_D void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "CPB_MBAR_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra CPB_MBAR_DONE_%=;\n"
      "bra CPB_MBAR_WAIT_%=;\n"
      "CPB_MBAR_DONE_%=:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
CPB_D void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
#endif

// ---------------------------------------------------------------------------------------------
// 16-byte asynchronous copies global -> shared (cp.async, SASS LDGSTS): the gather of plane-wave
// coefficients lands in shared memory without holding registers while in flight.  Groups are
// per thread; in the simulator build the copy happens at issue time.
// ---------------------------------------------------------------------------------------------
#if defined(CPB_EMULATE)
inline void cp_async16(void* dst, const void* src) {
  const char* s = static_cast<const char*>(src);
  char* d = static_cast<char*>(dst);
  for (int i = 0; i < 16; ++i) d[i] = s[i];
}
inline uint64_t l2_keep_policy() { return 0; }
inline void cp_async16_keep(void* dst, const void* src, uint64_t) { cp_async16(dst, src); }
inline void cp_async_commit() {}
inline void cp_async_wait_all() {}
#else
CPB_D void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// the same copy with an L2 evict-last policy: the plane-wave columns are gathered as 16-byte halves of 32-byte
// sectors whose other half belongs to a different ray (another block); the sector should survive in L2 until
// that block asks for it while the streaming intermediates pass through
CPB_D uint64_t l2_keep_policy() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
CPB_D void cp_async16_keep(void* dst, const void* src, uint64_t pol) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "l"(pol)
               : "memory");
}
CPB_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
CPB_D void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
#endif

}  // namespace cpb
