// The REFERENCE's own GPU path for rhoofr / vpsi, assembled from the reference's CUDA sources for
// bench.py's `reference_gpu` leg and for a device-side parity check made of reference code.
//
// TEST / BASELINE INFRASTRUCTURE ONLY - never linked into the product (cpmd_b200/ loads libcpb200.so and
// nothing else).  Built by oracle/Makefile into oracle/_ref/libref_gpu.so from
//   /root/reference/src/cuuser_utils.cu          the reference's kernel launch wrappers (CuUser_C_*)
//   /root/reference/src/cuuser_utils_kernels.cu  the reference's kernels
// compiled with nvcc WHERE THEY LIE (nothing is copied into this repo), plus this driver, which plays
// the part of the Fortran callers that cannot be compiled here (no Fortran compiler):
//   cuFFT plans laid out exactly like fft_create_cufft_plan        cp_cufft_utils.mod.F90:336-372
//   mltfft_cuda = cufftExecZ2Z + cublasZdscal + SetBlock2Zero        mltfft_utils.mod.F90:612-656
//   the stage order of fftcu_inv_sprs_1/2 and fftcu_frw_sprs_1/2     fftcu_methods.mod.F90:47-129, 224-300
//   incl. its host round trip per 3-D transform: the all2all lives on the host, and with the shipped
//   settings use_cpu_unpack_x2y = use_cpu_pack_y2x = .TRUE. (fftcu_methods.mod.F90:40-41) so do the two
//   scatters (unpack_x2y / pack_y2x, fftutil_utils.mod.F90:394-477, 206-290, restated below)
//   the state loops of rhoofr / vpsi on their GPU branches            rhoofr_utils.mod.F90:305-410,
//                                                                     vpsi_utils.mod.F90:376-675
// One task (nproc = 1: mp_all2all is a copy, all2all.inc:14-16), one device, one stream.
// `device_scatter` != 0 selects the reference's GPU kernels for the two scatters instead (the code path
// behind use_cpu_* = .FALSE., which the reference ships disabled); the host copies stay.
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cufft.h>

#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#define __HAS_CUDA 1
#include "cuuser_utils.h"  // the reference's header (-I /root/reference/src)

typedef std::complex<double> cpx;

namespace {
thread_local std::string g_err;
int fail(const std::string& m) {
  g_err = m;
  return -1;
}
#define CK(call)                                                                      \
  do {                                                                                \
    cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess) throw std::string(#call ": ") + cudaGetErrorString(e_);    \
  } while (0)
#define CKF(call)                                                        \
  do {                                                                   \
    if ((call) != CUFFT_SUCCESS) throw std::string(#call " failed");     \
  } while (0)

struct Plans {
  // key: transa, transb, ldax, ldbx, n, m   (cp_cufft_get_plan)
  std::map<std::tuple<char, char, int, int, int, int>, cufftHandle> p;
  cufftHandle get(char ta, char tb, int ldax, int ldbx, int n, int m, cudaStream_t st) {
    auto key = std::make_tuple(ta, tb, ldax, ldbx, n, m);
    auto it = p.find(key);
    if (it != p.end()) return it->second;
    cufftHandle h;
    int nn = n;
    // fft_create_cufft_plan, cp_cufft_utils.mod.F90:353-367: (istride, idist, ostride, odist)
    int is, id, os, od;
    if (ta == 'N') { is = 1; id = ldax; } else { is = ldax; id = 1; }
    if (tb == 'N') { os = 1; od = ldbx; } else { os = ldbx; od = 1; }
    CKF(cufftPlanMany(&h, 1, &nn, &nn, is, id, &nn, os, od, CUFFT_Z2Z, m));
    CKF(cufftSetStream(h, st));
    p[key] = h;
    return h;
  }
};
}  // namespace

struct refgpu {
  int n1, n2, n3, kr1, kr2, kr3, ngw, nrays, kr3min, kr3max, geq0;
  double tpiba2, omega;
  size_t maxfft;  // kr1*kr2*kr3 (fft_maxfft: the size of t1/t2 and of the host buffers xf/yf)
  std::vector<int> nzhs, indzs, msp;
  std::vector<double> hg;
  cudaStream_t st = nullptr;
  cublasHandle_t blas = nullptr;
  Plans plans;
  cufftDoubleComplex *t1 = nullptr, *t2 = nullptr;
  int *nzfs_d = nullptr, *inzs_d = nullptr, *msqs_d = nullptr, *sp9_d = nullptr, *sp5_d = nullptr, *lrxpl_d = nullptr;
  cpx *xf = nullptr, *yf = nullptr, *fh = nullptr;  // pinned host buffers
  double* real_d = nullptr;                          // rho or V
  cufftDoubleComplex* c0_d = nullptr;
  size_t c0_cap = 0;
  long launches = 0;
};

namespace {

// mltfft_cuda (mltfft_utils.mod.F90:612-656)
void mltfft_cuda(refgpu* h, char ta, char tb, cufftDoubleComplex* a, int ldax, int lday, cufftDoubleComplex* b,
                 int ldbx, int ldby, int n, int m, int isign, double scale) {
  cufftHandle plan = h->plans.get(ta, tb, ldax, ldbx, n, m, h->st);
  CKF(cufftExecZ2Z(plan, a, b, isign == 1 ? CUFFT_FORWARD : CUFFT_INVERSE));
  if (std::fabs(scale - 1.0) > 1e-12) {
    cublasZdscal(h->blas, ldax * lday, &scale, (cuDoubleComplex*)b, 1);  // :644 (with the reference's own count)
    h->launches += 1;
  }
  char t[2] = {tb, 0};
  CuUser_C_SetBlock2Zero(b, t, n, m, ldbx, ldby, h->st);  // :645
  h->launches += 2;
}

// unpack_x2y for one task (fftutil_utils.mod.F90:394-477, default copy branch): zeroing(yf) + scatter
void host_unpack_x2y(refgpu* h, const cpx* xf, cpx* yf, int mm, int lr1) {
  const size_t mf = h->maxfft;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)mf; ++i) yf[i] = 0.0;  // zeroing is threaded in the reference (zeroing_utils)
  const int mxrp = h->nrays;
  for (int i = 0; i < lr1; ++i) {                   // the reference parallelises over tasks only: one here
    const cpx* s = xf + (size_t)i * mxrp;
    cpx* d = yf + (size_t)i * mm;
    for (int k = 0; k < mxrp; ++k) d[h->msp[k] - 1] = s[k];  // zsctr_no_omp
  }
}
// pack_y2x for one task (fftutil_utils.mod.F90:206-290): gather
void host_pack_y2x(refgpu* h, cpx* xf, const cpx* yf, int mm, int lr1) {
  const int mxrp = h->nrays;
  for (int i = 0; i < lr1; ++i) {
    cpx* d = xf + (size_t)i * mxrp;
    const cpx* s = yf + (size_t)i * mm;
    for (int k = 0; k < mxrp; ++k) d[k] = s[h->msp[k] - 1];  // zgthr_no_omp
  }
}

// invfftn(psi,.TRUE.) with copy_data_to_device = copy_data_to_host = .FALSE.: input in t1 (ray storage),
// output in t1 (kr1,kr2s,kr3s)                                         fftmain_utils.mod.F90:215-228
void inv_sparse(refgpu* h, int device_scatter) {
  const int nzb = h->kr3max - h->kr3min + 1;
  int m = h->nrays;
  const int lda = h->nrays * h->kr1;  // lsrm*lr1m
  const int mm = h->kr2 * nzb;
  // fftcu_inv_sprs_1 (fftcu_methods.mod.F90:47-84)
  mltfft_cuda(h, 'N', 'T', h->t1, h->kr1, m, h->t2, m, h->kr1, h->n1, m, -1, 1.0);
  CuUser_C_Pack_x2y(h->t2, h->t1, h->nrays, lda, h->lrxpl_d, h->sp5_d, (int)h->maxfft, 1, false, h->st);
  h->launches += 1;
  CK(cudaMemcpyAsync(h->yf, h->t1, h->maxfft * sizeof(cpx), cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  std::memcpy(h->xf, h->yf, (size_t)lda * sizeof(cpx));  // fft_comm -> mp_all2all with one task = copy
  // fftcu_inv_sprs_2 (:86-129)
  if (!device_scatter) {
    host_unpack_x2y(h, h->xf, h->yf, mm, h->n1);
    CK(cudaMemcpyAsync(h->t2, h->yf, h->maxfft * sizeof(cpx), cudaMemcpyHostToDevice, h->st));
  } else {
    CK(cudaMemcpyAsync(h->t1, h->xf, h->maxfft * sizeof(cpx), cudaMemcpyHostToDevice, h->st));
    CuUser_C_Unpack_x2y(h->t1, h->t2, mm, h->n1, lda, h->msqs_d, h->nrays, h->sp9_d, (int)h->maxfft, 1, false, h->st);
    h->launches += 2;
  }
  m = nzb * h->kr1;
  mltfft_cuda(h, 'N', 'T', h->t2, h->kr2, m, h->t1, m, h->kr2, h->n2, m, -1, 1.0);
  m = h->kr1 * h->kr2;
  CuUser_C_PutZ(h->t1, h->t2, h->kr3min, h->kr3max, h->kr3, m, h->st);
  h->launches += 2;
  mltfft_cuda(h, 'N', 'T', h->t2, h->kr3, m, h->t1, m, h->kr3, h->n3, m, -1, 1.0);
}

// fwfftn(psi,.TRUE.) with copy_data_to_device = .FALSE., copy_to_host = .TRUE.: input t1 (kr1,kr2s,kr3s),
// output on the host in fh (kr1s * nrays ray storage)               fftmain_utils.mod.F90:244-256
void fwd_sparse(refgpu* h, int device_scatter) {
  const int nzb = h->kr3max - h->kr3min + 1;
  const int lda = h->nrays * h->kr1;
  const int mm = h->kr2 * nzb;
  int m = h->kr1 * h->kr2;
  // fftcu_frw_sprs_1 (fftcu_methods.mod.F90:224-268)
  mltfft_cuda(h, 'T', 'N', h->t1, m, h->kr3, h->t2, h->kr3, m, h->n3, m, 1, 1.0);
  CuUser_C_GetZ(h->t2, h->t1, h->kr3min, h->kr3max, h->kr3, m, h->st);
  h->launches += 1;
  m = nzb * h->kr1;
  mltfft_cuda(h, 'T', 'N', h->t1, m, h->kr2, h->t2, h->kr2, m, h->n2, m, 1, 1.0);
  if (!device_scatter) {
    CK(cudaMemcpyAsync(h->yf, h->t2, h->maxfft * sizeof(cpx), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    host_pack_y2x(h, h->xf, h->yf, mm, h->n1);
  } else {
    CuUser_C_Pack_y2x(h->t1, h->t2, mm, h->n1, lda, h->msqs_d, h->nrays, h->sp9_d, (int)h->maxfft, 1, false, h->st);
    h->launches += 1;
    CK(cudaMemcpyAsync(h->xf, h->t1, h->maxfft * sizeof(cpx), cudaMemcpyDeviceToHost, h->st));
  }
  CK(cudaStreamSynchronize(h->st));
  std::memcpy(h->yf, h->xf, (size_t)lda * sizeof(cpx));  // fft_comm
  // fftcu_frw_sprs_2 (:270-300)
  const double scale = 1.0 / ((double)h->n1 * h->n2 * h->n3);
  CK(cudaMemcpyAsync(h->t1, h->yf, h->maxfft * sizeof(cpx), cudaMemcpyHostToDevice, h->st));
  CuUser_C_Unpack_y2x(h->t2, h->t1, mm, h->nrays, lda, h->lrxpl_d, h->sp5_d, (int)h->maxfft, 1, false, h->st);
  h->launches += 1;
  m = h->nrays;
  mltfft_cuda(h, 'T', 'N', h->t2, m, h->kr1, h->t1, h->kr1, m, h->n1, m, 1, scale);
  CK(cudaMemcpyAsync(h->fh, h->t1, (size_t)h->kr1 * m * sizeof(cpx), cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
}

void upload_c0(refgpu* h, const cpx* c0, long ld, int nstate) {
  const size_t n = (size_t)ld * nstate;
  if (n > h->c0_cap) {
    if (h->c0_d) cudaFree(h->c0_d);
    h->c0_d = nullptr;
    CK(cudaMalloc(&h->c0_d, n * sizeof(cpx)));
    h->c0_cap = n;
  }
  CK(cudaMemcpyAsync(h->c0_d, c0, n * sizeof(cpx), cudaMemcpyHostToDevice, h->st));  // cp_cuwfn: c0 lives on the device
}

void set_psi(refgpu* h, long ld, int is1, int is2) {
  CK(cudaMemsetAsync(h->t1, 0, h->maxfft * sizeof(cpx), h->st));  // cuda_mem_zero_bytes(psi_d)
  if (is2 < 0)
    CuUser_C_Set_Psi_1_Stage_G(1.0, 0.0, h->c0_d + (size_t)is1 * ld, h->t1, h->ngw, h->nzfs_d, h->inzs_d, h->geq0 != 0,
                               h->st);
  else
    CuUser_C_Set_Psi_2_Stages_G(h->c0_d + (size_t)is1 * ld, h->c0_d + (size_t)is2 * ld, h->t1, h->ngw, h->nzfs_d,
                                h->inzs_d, h->geq0 != 0, h->st);
  h->launches += 1;
}

}  // namespace

extern "C" {

const char* refgpu_last_error(void) { return g_err.c_str(); }
const char* refgpu_source(void) { return REF_SRC_DIR; }

int refgpu_create(refgpu** out, const int* nr, const int* kr, int ngw, int nrays, const int* nzhs, const int* indzs,
                  const int* msp2, int kr3min, int kr3max, const double* hg, int geq0, double tpiba2, double omega) {
  refgpu* h = new refgpu();
  try {
    h->n1 = nr[0]; h->n2 = nr[1]; h->n3 = nr[2];
    h->kr1 = kr[0]; h->kr2 = kr[1]; h->kr3 = kr[2];
    h->ngw = ngw; h->nrays = nrays; h->kr3min = kr3min; h->kr3max = kr3max; h->geq0 = geq0;
    h->tpiba2 = tpiba2; h->omega = omega;
    h->maxfft = (size_t)kr[0] * kr[1] * kr[2];
    h->nzhs.assign(nzhs, nzhs + ngw);
    h->indzs.assign(indzs, indzs + ngw);
    h->msp.assign(msp2, msp2 + nrays);
    h->hg.assign(hg, hg + ngw);
    CK(cudaStreamCreate(&h->st));
    if (cublasCreate(&h->blas) != CUBLAS_STATUS_SUCCESS) throw std::string("cublasCreate failed");
    cublasSetStream(h->blas, h->st);
    CK(cudaMalloc(&h->t1, h->maxfft * sizeof(cpx)));
    CK(cudaMalloc(&h->t2, h->maxfft * sizeof(cpx)));
    CK(cudaMalloc(&h->real_d, h->maxfft * sizeof(double)));
    CK(cudaMallocHost(&h->xf, h->maxfft * sizeof(cpx)));
    CK(cudaMallocHost(&h->yf, h->maxfft * sizeof(cpx)));
    CK(cudaMallocHost(&h->fh, h->maxfft * sizeof(cpx)));
    auto up = [&](const std::vector<int>& v) {
      int* d = nullptr;
      CK(cudaMalloc(&d, v.size() * sizeof(int)));
      CK(cudaMemcpy(d, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice));
      return d;
    };
    h->nzfs_d = up(h->nzhs);
    h->inzs_d = up(h->indzs);
    h->msqs_d = up(h->msp);
    h->sp9_d = up(std::vector<int>(1, nrays));   // rays of task 0
    h->sp5_d = up(std::vector<int>(1, nr[0]));   // x planes of task 0
    h->lrxpl_d = up(std::vector<int>(1, 1));     // first x plane of task 0
    *out = h;
    return 0;
  } catch (const std::string& e) {
    delete h;
    return fail(e);
  }
}

int refgpu_destroy(refgpu* h) {
  if (!h) return 0;
  for (auto& kv : h->plans.p) cufftDestroy(kv.second);
  cudaFree(h->t1); cudaFree(h->t2); cudaFree(h->real_d); cudaFree(h->c0_d);
  cudaFree(h->nzfs_d); cudaFree(h->inzs_d); cudaFree(h->msqs_d); cudaFree(h->sp9_d); cudaFree(h->sp5_d);
  cudaFree(h->lrxpl_d);
  cudaFreeHost(h->xf); cudaFreeHost(h->yf); cudaFreeHost(h->fh);
  if (h->blas) cublasDestroy(h->blas);
  if (h->st) cudaStreamDestroy(h->st);
  delete h;
  return 0;
}

long refgpu_launches(const refgpu* h) { return h ? h->launches : 0; }

// rhoofr on its GPU branch (rhoofr_utils.mod.F90:305-410, 414-440): c0 (ld, nstate) host -> rhoe (maxfft) host
int refgpu_rhoofr(refgpu* h, const void* c0, long ld, int nstate, const double* f, double* rhoe, int device_scatter) {
  try {
    upload_c0(h, (const cpx*)c0, ld, nstate);
    CK(cudaMemsetAsync(h->real_d, 0, h->maxfft * sizeof(double), h->st));
    for (int i = 0; i < nstate; i += 2) {
      const int is1 = i, is2 = (i + 1 < nstate) ? i + 1 : -1;
      if (f[is1] == 0.0 && (is2 < 0 || f[is2] == 0.0)) continue;  // :312-316
      set_psi(h, ld, is1, is2);
      inv_sparse(h, device_scatter);
      const double coef3 = f[is1] / h->omega, coef4 = is2 < 0 ? 0.0 : f[is2] / h->omega;  // :369-374
      CuUser_C_Build_Density_Sum(coef3, coef4, h->t1, h->real_d, (int)h->maxfft, h->st);  // cp_cubuild_density_sum
      h->launches += 1;
    }
    CK(cudaMemcpyAsync(rhoe, h->real_d, h->maxfft * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    return 0;
  } catch (const std::string& e) {
    return fail(e);
  }
}

// vpsi on its GPU branch (vpsi_utils.mod.F90:376-675): c2 (ld, nstate) host += result
int refgpu_vpsi(refgpu* h, const void* c0v, void* c2v, long ld, int nstate, const double* f, const double* vpot,
                int device_scatter) {
  try {
    const cpx* c0 = (const cpx*)c0v;
    cpx* c2 = (cpx*)c2v;
    upload_c0(h, c0, ld, nstate);
    CK(cudaMemcpyAsync(h->real_d, vpot, h->maxfft * sizeof(double), cudaMemcpyHostToDevice, h->st));
    for (int i = 0; i < nstate; i += 2) {
      const int is1 = i, is2 = (i + 1 < nstate) ? i + 1 : -1;
      set_psi(h, ld, is1, is2);
      inv_sparse(h, device_scatter);
      CuUser_C_Build_Pointwise_CxR(h->t1, h->real_d, (int)h->maxfft, h->st);  // cp_cuapply_potential
      h->launches += 1;
      fwd_sparse(h, device_scatter);
      // the decode loop runs on the host in the reference too (vpsi_utils.mod.F90:626-673), restated here
      double fi = f[is1] * 0.5, fip1 = is2 < 0 ? 0.0 : f[is2] * 0.5;
      if (fi == 0.0) fi = 1.0;
      if (fip1 == 0.0) fip1 = 1.0;
#pragma omp parallel for schedule(static)
      for (int ig = 0; ig < h->ngw; ++ig) {
        const cpx psin = h->fh[h->nzhs[ig] - 1], psii = h->fh[h->indzs[ig] - 1];
        const cpx fp = psin + psii, fm = psin - psii;
        const double g2 = h->tpiba2 * h->hg[ig];
        const cpx a = c0[(size_t)is1 * ld + ig];
        c2[(size_t)is1 * ld + ig] += cpx(-fi * (g2 * a.real() + fp.real()), -fi * (g2 * a.imag() + fm.imag()));
        if (is2 >= 0) {
          const cpx b = c0[(size_t)is2 * ld + ig];
          c2[(size_t)is2 * ld + ig] += cpx(-fip1 * (g2 * b.real() + fp.imag()), -fip1 * (g2 * b.imag() - fm.real()));
        }
      }
    }
    return 0;
  } catch (const std::string& e) {
    return fail(e);
  }
}

}  // extern "C"
