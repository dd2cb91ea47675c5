// Thin runtime layer: CUDA runtime calls with error checking (product build) or plain host
// memory (kernel functional simulator build, tests only — see cpb_defs.h).
#pragma once
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include "cpb_defs.h"

namespace cpb {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

namespace rt {

#if defined(CPB_EMULATE)

inline void set_device(int) {}
inline int device_count() { return 1; }
inline int sm_count(int) { return 4; }
inline void* dmalloc(size_t n) {
  void* p = std::malloc(n ? n : 1);
  if (!p) throw Error(-3, "out of host memory (emulated device)");
  std::memset(p, 0xCD, n);  // poison
  return p;
}
inline void dfree(void* p) { std::free(p); }
inline void* hmalloc_pinned(size_t n) { return std::malloc(n ? n : 1); }
inline void hfree_pinned(void* p) { std::free(p); }
inline void h2d(void* d, const void* h, size_t n, cudaStream_t) { std::memcpy(d, h, n); }
inline void d2h(void* h, const void* d, size_t n, cudaStream_t) { std::memcpy(h, d, n); }
inline void d2d(void* dst, const void* src, size_t n, cudaStream_t) { std::memcpy(dst, src, n); }
inline void dzero(void* d, size_t n, cudaStream_t) { std::memset(d, 0, n); }
inline void sync(cudaStream_t) {}
inline void check_last(const char*) {}
inline void check_last_clear() {}
inline cudaStream_t stream_create() { return nullptr; }
inline void stream_destroy(cudaStream_t) {}
typedef void* event_t;
inline event_t event_create() { return nullptr; }
inline void event_destroy(event_t) {}
inline void event_record(event_t, cudaStream_t) {}
inline void stream_wait(cudaStream_t, event_t) {}
inline void event_sync(event_t) {}
inline event_t timing_event_create() { return nullptr; }
inline double event_elapsed_ms(event_t, event_t) { return 0.0; }

#else

inline void ck(cudaError_t e, const char* what) {
  if (e != cudaSuccess) {
    throw Error(e == cudaErrorMemoryAllocation ? -3 : -2,
                std::string(what) + ": " + cudaGetErrorString(e));
  }
}
inline void set_device(int d) { ck(cudaSetDevice(d), "cudaSetDevice"); }
inline int device_count() {
  int n = 0;
  ck(cudaGetDeviceCount(&n), "cudaGetDeviceCount");
  return n;
}
inline int sm_count(int dev) {
  int n = 0;
  ck(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute");
  return n > 0 ? n : 148;
}
inline void* dmalloc(size_t n) {
  void* p = nullptr;
  ck(cudaMalloc(&p, n ? n : 1), "cudaMalloc");
  return p;
}
inline void dfree(void* p) {
  if (p) cudaFree(p);
}
inline void* hmalloc_pinned(size_t n) {
  void* p = nullptr;
  ck(cudaMallocHost(&p, n ? n : 1), "cudaMallocHost");
  return p;
}
inline void hfree_pinned(void* p) {
  if (p) cudaFreeHost(p);
}
inline void h2d(void* d, const void* h, size_t n, cudaStream_t s) {
  ck(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s), "cudaMemcpyAsync H2D");
}
inline void d2h(void* h, const void* d, size_t n, cudaStream_t s) {
  ck(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync D2H");
}
inline void d2d(void* dst, const void* src, size_t n, cudaStream_t s) {
  ck(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, s), "cudaMemcpyAsync D2D");
}
inline void dzero(void* d, size_t n, cudaStream_t s) { ck(cudaMemsetAsync(d, 0, n, s), "cudaMemsetAsync"); }
inline void sync(cudaStream_t s) { ck(cudaStreamSynchronize(s), "cudaStreamSynchronize"); }
inline void check_last(const char* what) { ck(cudaGetLastError(), what); }
inline void check_last_clear() { (void)cudaGetLastError(); }
inline cudaStream_t stream_create() {
  cudaStream_t s;
  ck(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate");
  return s;
}
inline void stream_destroy(cudaStream_t s) {
  if (s) cudaStreamDestroy(s);
}
typedef cudaEvent_t event_t;
inline event_t event_create() {
  cudaEvent_t e;
  ck(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
  return e;
}
inline void event_destroy(event_t e) {
  if (e) cudaEventDestroy(e);
}
inline void event_record(event_t e, cudaStream_t s) { ck(cudaEventRecord(e, s), "cudaEventRecord"); }
inline void stream_wait(cudaStream_t s, event_t e) { ck(cudaStreamWaitEvent(s, e, 0), "cudaStreamWaitEvent"); }
inline void event_sync(event_t e) { ck(cudaEventSynchronize(e), "cudaEventSynchronize"); }
inline event_t timing_event_create() {
  cudaEvent_t e;
  ck(cudaEventCreate(&e), "cudaEventCreate");
  return e;
}
inline double event_elapsed_ms(event_t a, event_t b) {
  float ms = 0.f;
  ck(cudaEventElapsedTime(&ms, a, b), "cudaEventElapsedTime");
  return (double)ms;
}

#endif

}  // namespace rt
}  // namespace cpb
