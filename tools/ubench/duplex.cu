// PCIe duplex micro-benchmark: does a D2H stream overlap with an H2D stream when all H2D chunks are
// enqueued first (the host-pointer cpb_vpsi pattern)?  1-D vs 2-D async copies, submission orders.
// build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o duplex duplex.cu
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void spin(double* p, size_t n, int it) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (int k = 0; k < it; ++k)
    for (size_t j = i; j < n; j += (size_t)gridDim.x * blockDim.x) p[j] = p[j] * 1.0000001 + 1e-9;
}

int main() {
  const size_t chunk = 237ull << 20, nch = 8, width = 3702016;  // ~ one batch of c2 columns
  const size_t rows = chunk / width;
  char *hin, *hout, *din, *dout;
  double* work;
  CK(cudaMallocHost(&hin, chunk * nch));
  CK(cudaMallocHost(&hout, chunk * nch));
  CK(cudaMalloc(&din, chunk * nch));
  CK(cudaMalloc(&dout, chunk * nch));
  CK(cudaMalloc(&work, 1ull << 30));
  cudaStream_t sm, si, so;
  CK(cudaStreamCreateWithFlags(&sm, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&si, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&so, cudaStreamNonBlocking));
  cudaEvent_t ein[8], edone[8];
  for (int i = 0; i < 8; ++i) { CK(cudaEventCreateWithFlags(&ein[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&edone[i], cudaEventDisableTiming)); }
  auto h2d = [&](int c, bool two_d) {
    if (two_d) CK(cudaMemcpy2DAsync(din + c * chunk, width, hin + c * chunk, width, width, rows, cudaMemcpyHostToDevice, si));
    else CK(cudaMemcpyAsync(din + c * chunk, hin + c * chunk, rows * width, cudaMemcpyHostToDevice, si));
    CK(cudaEventRecord(ein[c], si));
  };
  auto d2h = [&](int c, bool two_d) {
    CK(cudaEventRecord(edone[c], sm));
    CK(cudaStreamWaitEvent(so, edone[c], 0));
    if (two_d) CK(cudaMemcpy2DAsync(hout + c * chunk, width, dout + c * chunk, width, width, rows, cudaMemcpyDeviceToHost, so));
    else CK(cudaMemcpyAsync(hout + c * chunk, dout + c * chunk, rows * width, cudaMemcpyDeviceToHost, so));
  };
  for (int kern = 0; kern < 2; ++kern)
    for (int two_d = 0; two_d < 2; ++two_d)
      for (int mode = 0; mode < 4; ++mode) {
        // mode 0: H2D only; 1: D2H only; 2: all H2D first, D2H after each "batch"; 3: H2D chunk c+1 submitted after D2H c-1
        double best = 1e30;
        for (int rep = 0; rep < 3; ++rep) {
          CK(cudaDeviceSynchronize());
          auto t0 = std::chrono::steady_clock::now();
          if (mode == 0 || mode == 2) for (int c = 0; c < (int)nch; ++c) h2d(c, two_d);
          if (mode == 3) h2d(0, two_d);
          for (int c = 0; c < (int)nch; ++c) {
            if (mode == 3 && c + 1 < (int)nch) h2d(c + 1, two_d);
            if (mode != 1) CK(cudaStreamWaitEvent(sm, ein[c], 0));
            if (kern) spin<<<1184, 256, 0, sm>>>(work, (1ull << 30) / 8, 8);
            if (mode != 0) d2h(c, two_d);
          }
          CK(cudaDeviceSynchronize());
          double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
          if (ms < best) best = ms;
        }
        const double gb = chunk * nch / 1e9 * (mode >= 2 ? 2 : 1);
        printf("kernels %d  %s  mode %d: %.1f ms  (%.1f GB/s aggregate)\n", kern, two_d ? "2D" : "1D", mode, best, gb / best * 1e3);
      }
  return 0;
}
