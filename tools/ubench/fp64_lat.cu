// FP64 dependent-issue latency / ILP curve on sm_100a: C independent DFMA (or DADD) chains per thread,
// W warps per SM sub-partition.  Prints cycles per instruction per warp.
#include <cstdio>
#include <cuda_runtime.h>
template <int C, int OP>
__global__ void k(double* out, int iters, double a, double b, long long* cyc) {
  double x[C];
#pragma unroll
  for (int i = 0; i < C; ++i) x[i] = threadIdx.x * 1e-3 + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
#pragma unroll
      for (int i = 0; i < C; ++i) x[i] = OP ? x[i] + a : fma(x[i], a, b);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < C; ++i) s += x[i];
  if (s == 123.456) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int C, int OP>
void run(int warps_per_smsp) {
  double* d; long long* c; cudaMalloc(&d, 8); cudaMalloc(&c, 8);
  int iters = 2000;
  k<C, OP><<<148, 128 * warps_per_smsp>>>(d, iters, 1.0000001, 1e-9, c);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  double per = (double)h / (iters * 16.0 * C);
  printf("%s chains=%d warps/SMSP=%d: %.2f cycles per instr per warp  (SMSP issue interval %.2f cycles)\n", OP ? "DADD" : "DFMA", C,
         warps_per_smsp, per, per / warps_per_smsp);
}
int main() {
  run<1, 0>(1); run<2, 0>(1); run<4, 0>(1); run<8, 0>(1);
  run<1, 1>(1); run<2, 1>(1); run<4, 1>(1); run<8, 1>(1);
  run<1, 1>(2); run<2, 1>(2); run<1, 1>(4); run<2, 1>(4); run<4, 1>(3);
  return 0;
}
