// Functional simulator of the small CUDA subset the cpb200 kernels use (TEST INFRASTRUCTURE).
//
// One ucontext fiber per CUDA thread; a block runs to completion before the next one starts on
// the same OS thread; __syncthreads() yields to the block scheduler, which resumes every fiber of
// the block once per barrier phase.  Blocks are distributed over OS threads with OpenMP; all
// simulator state is thread_local, and kernels use dynamic shared memory only (CPB_DYN_SMEM),
// so blocks are independent exactly as on the device.
//
// This is NOT a CPU fallback of the product: it is compiled only into tests/emu/libcpb200_emu.so
// which only tests/ load, to validate kernel index arithmetic where no GPU is available.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

struct double2 {
  double x, y;
};
static inline double2 make_double2(double x, double y) {
  double2 r;
  r.x = x;
  r.y = y;
  return r;
}

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3 {
  unsigned x, y, z;
};

typedef void* cudaStream_t;

extern thread_local uint3 threadIdx;
extern thread_local uint3 blockIdx;
extern thread_local dim3 blockDim;
extern thread_local dim3 gridDim;

void __syncthreads();
void __syncwarp();  // barrier over the (up to) 32 fibers of a warp that have not exited

template <class T>
static inline T __ldg(const T* p) {
  return *p;
}

namespace emu {
void* dyn_smem();
void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body);
}  // namespace emu
