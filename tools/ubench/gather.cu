// SM-side cost of 16-byte gathers on sm_100a: how many elements per clock and SM can a kernel pull out of an
// L2-resident array when every lane of a load instruction hits a different 128-byte line?
// Variants: cp.async (LDGSTS.128) into private shared-memory slots, ld.global.nc (LDG.128) into registers,
// each with scattered and with coalesced addresses (lanes read consecutive 16-byte elements), U loads in
// flight per thread.  The array (4 MB) stays in L2, so DRAM does not enter.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather gather.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* d, const void* s) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d)), "l"(s) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

constexpr int kElems = 1 << 18;  // 16-byte elements: 4 MB

// MODE 0: LDGSTS, 1: LDG.  SC: scattered (index hashed per lane) or coalesced.
template <int MODE, bool SC, int U>
__global__ void k(const double2* __restrict__ a, double* out, int iters) {
  extern __shared__ double2 S[];
  const int tid = threadIdx.x;
  unsigned h = (blockIdx.x * blockDim.x + tid) * 2654435761u;
  double2 acc = make_double2(0.0, 0.0);
  for (int it = 0; it < iters; ++it) {
    unsigned idx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      h = h * 1664525u + 1013904223u;
      // scattered: every lane its own line; coalesced: the warp reads 32 consecutive elements
      idx[u] = SC ? (h >> 8) % kElems : (((h >> 8) * 0u + (blockIdx.x * 977u + it * U + u) * 64u + (tid & ~31u) * 7u) % (kElems - 32)) + (tid & 31);
    }
    if (MODE == 0) {
#pragma unroll
      for (int u = 0; u < U; ++u) cp_async16(&S[u * blockDim.x + tid], a + idx[u]);
      cp_commit();
      cp_wait();
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const double2 v = S[u * blockDim.x + tid];
        acc.x += v.x;
        acc.y += v.y;
      }
    } else {
      double2 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = __ldg(a + idx[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        acc.x += v[u].x;
        acc.y += v[u].y;
      }
    }
  }
  if (acc.x == 123.456) out[0] = acc.y;
}

template <int MODE, bool SC, int U>
void run(const double2* a, double* out, int warps_per_sm, int clock_mhz) {
  const int threads = 128, blocks_per_sm = warps_per_sm / 4, iters = 400;
  const size_t smem = (size_t)U * threads * sizeof(double2);
  cudaFuncSetAttribute(k<MODE, SC, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE, SC, U><<<148 * blocks_per_sm, threads, smem>>>(a, out, 20);
  cudaEventRecord(e0);
  k<MODE, SC, U><<<148 * blocks_per_sm, threads, smem>>>(a, out, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double elems = (double)148 * blocks_per_sm * threads * iters * U;
  const double per_clk_sm = elems / (ms * 1e-3) / (clock_mhz * 1e6) / 148.0;
  printf("%-6s %-9s U=%d warps/SM=%2d: %7.3f ms  %6.2f elements/clk/SM  (%5.1f GB/s useful)\n", MODE ? "LDG" : "LDGSTS",
         SC ? "scattered" : "coalesced", U, warps_per_sm, ms, per_clk_sm, elems * 16 / (ms * 1e-3) / 1e9);
}

int main() {
  double2* a;
  double* out;
  cudaMalloc(&a, (size_t)kElems * sizeof(double2));
  cudaMalloc(&out, 8);
  cudaMemset(a, 0, (size_t)kElems * sizeof(double2));
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  clk /= 1000;
  for (int w : {8, 16, 32}) {
    run<0, true, 8>(a, out, w, clk);
    run<1, true, 8>(a, out, w, clk);
    run<0, false, 8>(a, out, w, clk);
    run<1, false, 8>(a, out, w, clk);
  }
  run<0, true, 4>(a, out, 16, clk);
  run<1, true, 4>(a, out, 16, clk);
  run<0, true, 16>(a, out, 16, clk);
  run<1, true, 16>(a, out, 16, clk);
  return 0;
}
