/* CPU stand-in for <cuda.h> (TEST INFRASTRUCTURE, authored here - not a copy of any CUDA header).
 *
 * oracle/Makefile compiles the REFERENCE's own kernel file /root/reference/src/cuuser_utils_kernels.cu
 * for the host with g++ (output: oracle/_ref/libcuuser_ref.so).  Those kernels are plain
 * one-thread-per-element functions (no shared memory, no __syncthreads), so the few CUDA names they
 * use are provided here as ordinary C++: __global__ expands to nothing and threadIdx / blockIdx /
 * blockDim / gridDim are globals that the driver (oracle/ref_cuuser_driver.cpp) steps through the
 * launch grid. */
#ifndef CPB_REF_SHIM_CUDA_H
#define CPB_REF_SHIM_CUDA_H
#define __global__
#define __host__
#define __device__
struct ref_uint3 {
  unsigned x, y, z;
};
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
extern ref_uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;
typedef void* cudaStream_t;
#endif
