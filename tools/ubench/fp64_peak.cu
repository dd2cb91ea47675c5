// FP64 pipe micro-benchmark for the roofline's secondary bound (B200, sm_100a).
// Measures DFMA / DADD / DMUL issue rates per SM with 8 independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double* out, int iters, double a, double b) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) x[i] = fma(x[i], a, b);
      if (OP == 1) x[i] = x[i] + a;
      if (OP == 2) x[i] = x[i] * a;
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 123.456) out[0] = s;
}
template <int OP>
void run(const char* name, int threads, int blocks_per_sm) {
  int nsm = 148;
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); nsm = p.multiProcessorCount;
  double* d; cudaMalloc(&d, 8);
  int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<OP><<<nsm * blocks_per_sm, threads>>>(d, 100, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<OP><<<nsm * blocks_per_sm, threads>>>(d, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)nsm * blocks_per_sm * threads * iters * 8.0;
  printf("%s threads=%d blocks/SM=%d: %.2f Tinstr-lane/s  (%.1f lanes/clk/SM at 1.965 GHz, %d SMs) %.3f ms\n", name, threads,
         blocks_per_sm, ops / ms * 1e-9, ops / (ms * 1e-3) / nsm / 1.965e9, nsm, ms);
}
int main() {
  for (int t : {128, 256, 1024}) {
    run<0>("DFMA", t, 1); run<1>("DADD", t, 1); run<2>("DMUL", t, 1);
  }
  run<0>("DFMA", 256, 4); run<1>("DADD", 256, 4);
  return 0;
}
