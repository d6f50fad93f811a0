// deb_background.cu -- sm_100a kernel and C-ABI of the table producer (deb_background.cuh): one CTA per cosmology.
//
// The only sequential part is the recombination chain (256 intervals x 2-9 GRKT4 steps, one thread); the neutrino tables,
// the 256 Romberg sums of d tau/d a and the Saha branch run on the other threads of the CTA beside it, so a batch of up
// to #SM x CTAs-per-SM cosmologies costs the latency of one chain.  The working set (BgWork, ~200 KB per cosmology)
// lives in the caller's workspace in HBM/L2; the chain itself runs in registers.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "../../include/discoeb_b200.h"
#include "deb_background.cuh"

using namespace deb::bg;

#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "[discoeb_b200] CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return DEB_E_CUDA; } } while (0)

__constant__ double g_nu_q[NNUQ], g_nu_w[NNUQ];

__global__ void __launch_bounds__(128) k_background(int ncosmo, int nth, const double* __restrict__ bg_in, double* __restrict__ scalars,
                                                     double* __restrict__ tables, BgWork* __restrict__ work, double* __restrict__ extras) {
  const int c = blockIdx.x;
  if (c >= ncosmo) return;
  const size_t tl = 3 * (size_t)(5 * nth + 2 * NNU);
  background_one(bg_in + (size_t)c * NBGIN, g_nu_q, g_nu_w, nth, scalars + (size_t)c * DEB_NSCAL, tables + (size_t)c * tl, work[c], threadIdx.x, blockDim.x,
                 extras ? extras + (size_t)c * extras_len(nth) : nullptr);
}

extern "C" {

size_t deb_background_workspace_bytes(int32_t ncosmo, int32_t nth) { (void)nth; return ncosmo > 0 ? (size_t)ncosmo * sizeof(BgWork) : 0; }
size_t deb_background_nin(void) { return NBGIN; }

size_t deb_background_extras_len(int32_t nth) { return nth > 0 ? extras_len(nth) : 0; }

int deb_background_f64(int32_t ncosmo, int32_t nth, const double* bg_in, double* scalars, double* tables, void* workspace,
                       size_t workspace_bytes, void* stream) {
  return deb_background_ex_f64(ncosmo, nth, bg_in, scalars, tables, nullptr, workspace, workspace_bytes, stream);
}

int deb_background_ex_f64(int32_t ncosmo, int32_t nth, const double* bg_in, double* scalars, double* tables, double* extras, void* workspace,
                          size_t workspace_bytes, void* stream) {
  if (ncosmo < 1 || nth < 16 || nth > NTH_MAX || !bg_in || !scalars || !tables || !workspace) return DEB_E_ARG;
  if (workspace_bytes < deb_background_workspace_bytes(ncosmo, nth)) return DEB_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  double q[NNUQ], w[NNUQ];
  nu_quadrature(q, w);
  CUDA_TRY(cudaMemcpyToSymbolAsync(g_nu_q, q, sizeof(q), 0, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyToSymbolAsync(g_nu_w, w, sizeof(w), 0, cudaMemcpyHostToDevice, st));
  k_background<<<ncosmo, 128, 0, st>>>(ncosmo, nth, bg_in, scalars, tables, (BgWork*)workspace, extras);
  CUDA_TRY(cudaGetLastError());
  return DEB_OK;
}

// host-buffer convenience: H2D of the parameter block, kernel, D2H of scalars and tables
int deb_background_host_f64(int32_t device, int32_t ncosmo, int32_t nth, const double* bg_in, double* scalars, double* tables, float* kernel_ms) {
  return deb_background_host_ex_f64(device, ncosmo, nth, bg_in, scalars, tables, nullptr, kernel_ms);
}

int deb_background_host_ex_f64(int32_t device, int32_t ncosmo, int32_t nth, const double* bg_in, double* scalars, double* tables, double* extras,
                               float* kernel_ms) {
  if (ncosmo < 1 || nth < 16 || nth > NTH_MAX || !bg_in || !scalars || !tables) return DEB_E_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return DEB_E_NODEVICE;
  CUDA_TRY(cudaSetDevice(device));
  const size_t tl = 3 * (size_t)(5 * nth + 2 * NNU);
  double *d_in = nullptr, *d_sc = nullptr, *d_tb = nullptr, *d_ex = nullptr; void* d_ws = nullptr;
  const size_t exb = extras ? (size_t)ncosmo * extras_len(nth) * 8 : 0;
  const size_t wsb = deb_background_workspace_bytes(ncosmo, nth);
  int rc = DEB_OK;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (cudaMalloc(&d_in, (size_t)ncosmo * NBGIN * 8) != cudaSuccess || cudaMalloc(&d_sc, (size_t)ncosmo * DEB_NSCAL * 8) != cudaSuccess ||
      cudaMalloc(&d_tb, (size_t)ncosmo * tl * 8) != cudaSuccess || cudaMalloc(&d_ws, wsb) != cudaSuccess ||
      (exb && cudaMalloc(&d_ex, exb) != cudaSuccess) ||
      cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) rc = DEB_E_CUDA;
  if (rc == DEB_OK && cudaMemcpy(d_in, bg_in, (size_t)ncosmo * NBGIN * 8, cudaMemcpyHostToDevice) != cudaSuccess) rc = DEB_E_CUDA;
  if (rc == DEB_OK) {
    cudaMemset(d_sc, 0, (size_t)ncosmo * DEB_NSCAL * 8);
    cudaEventRecord(e0, 0);
    rc = deb_background_ex_f64(ncosmo, nth, d_in, d_sc, d_tb, d_ex, d_ws, wsb, nullptr);
    cudaEventRecord(e1, 0);
  }
  if (rc == DEB_OK && (cudaMemcpy(scalars, d_sc, (size_t)ncosmo * DEB_NSCAL * 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
                       cudaMemcpy(tables, d_tb, (size_t)ncosmo * tl * 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
                       (exb && cudaMemcpy(extras, d_ex, exb, cudaMemcpyDeviceToHost) != cudaSuccess))) rc = DEB_E_CUDA;
  if (rc == DEB_OK && kernel_ms) cudaEventElapsedTime(kernel_ms, e0, e1);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  cudaFree(d_in); cudaFree(d_sc); cudaFree(d_tb); cudaFree(d_ws); cudaFree(d_ex);
  return rc;
}

}  // extern "C"
