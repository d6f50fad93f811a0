// deb_spectra.cu -- CUDA kernels and C-ABI of the spectra epilogues (deb_spectra.cuh).  Device-pointer entry for use
// right behind deb_evolve_f64 on the same stream (y never leaves the GPU), host-buffer entry for the Python API.
#include <cuda_runtime.h>
#include <stdio.h>
#include "../../include/discoeb_b200.h"
#include "deb_spectra.cuh"

using namespace deb::sp;

#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "[discoeb_b200] CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return DEB_E_CUDA; } } while (0)

__global__ void k_spec_point(int n, const double* y, const double* k, double As, double ns, double kp, double bias, double* dm, double* tm, double* P0,
                             double* P2, double* P4, double* lPd, double* lPt, double* Pd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) point(i, y, k, As, ns, kp, bias, dm, tm, P0, P2, P4, lPd, lPt, Pd);
}
__global__ void k_spec_smooth(int n, const double* lPd, const double* lPt, const double* coef, int w, double* Psd, double* Pst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { Psd[i] = smooth_at(i, n, lPd, coef, w); Pst[i] = smooth_at(i, n, lPt, coef, w); }
}
__global__ void k_spec_kaiser(int n, int nmu, const double* dm, const double* tm, const double* Psd, const double* Pst, int smoothed, double bias,
                              const double* mu, double* Pkmu) {
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (idx >= (long)n * nmu) return;
  const int i = (int)(idx / nmu), j = (int)(idx % nmu);
  const double d = smoothed ? sqrt(Psd[i]) : dm[i], t = smoothed ? -sqrt(Pst[i]) : tm[i];
  const double v = bias * d - mu[j] * mu[j] * t;
  Pkmu[idx] = v * v;
}
__global__ void k_spec_xi_fwd(int N, const double* k, const double* Pk, int ell, Cx* F) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m <= N / 2) F[m] = fftlog_forward(m, N, k, Pk, ell);
}
__global__ void k_spec_xi_bwd(int N, const double* k, const Cx* F, int ell, double* xi, double* r) {
  const int nn = blockIdx.x * blockDim.x + threadIdx.x;
  if (nn < N) { xi[N - 1 - nn] = fftlog_backward(nn, N, k, F, ell); r[N - 1 - nn] = 2.0 * M_PI / k[nn]; }     // ascending r
}

extern "C" {

size_t deb_spectra_workspace_bytes(int32_t nk) { return (size_t)(6 * nk) * 8 + (size_t)(nk / 2 + 1) * 16 + 256; }

int deb_spectra_f64(int32_t nk, int32_t nmu, const double* y, const double* kmodes, double As, double ns, double kp, double bias,
                    const double* sg_coef, int32_t sg_window, const double* mu, int32_t ell,
                    double* P0, double* P2, double* P4, double* Pkmu, double* Ps_delta, double* Ps_theta, double* xi, double* r,
                    void* workspace, size_t workspace_bytes, void* stream) {
  if (nk < 2 || !y || !kmodes || !P0 || !P2 || !P4 || !workspace) return DEB_E_ARG;
  if (workspace_bytes < deb_spectra_workspace_bytes(nk)) return DEB_E_WORKSPACE;
  if ((sg_window > 0) != (sg_coef != nullptr) || (sg_window > 0 && (!Ps_delta || !Ps_theta))) return DEB_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  double* w = (double*)workspace;
  double *dm = w, *tm = w + nk, *lPd = w + 2 * nk, *lPt = w + 3 * nk, *Pd = w + 4 * nk, *unused = w + 5 * nk;
  (void)unused;
  Cx* F = (Cx*)(((uintptr_t)(w + 6 * nk) + 15) & ~(uintptr_t)15);
  const int nb = (nk + 127) / 128;
  k_spec_point<<<nb, 128, 0, st>>>(nk, y, kmodes, As, ns, kp, bias, dm, tm, P0, P2, P4, lPd, lPt, Pd);
  if (sg_window > 0) k_spec_smooth<<<nb, 128, 0, st>>>(nk, lPd, lPt, sg_coef, sg_window, Ps_delta, Ps_theta);
  if (Pkmu && nmu > 0 && mu) {
    const long tot = (long)nk * nmu;
    k_spec_kaiser<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(nk, nmu, dm, tm, Ps_delta, Ps_theta, sg_window > 0, bias, mu, Pkmu);
  }
  if (xi && r) {
    // P_delta(k) for the transform: the smoothed spectrum when one was asked for, the raw one otherwise
    const double* Pk = sg_window > 0 ? Ps_delta : Pd;
    k_spec_xi_fwd<<<(nk / 2 + 1 + 63) / 64, 64, 0, st>>>(nk, kmodes, Pk, ell, F);
    k_spec_xi_bwd<<<(nk + 63) / 64, 64, 0, st>>>(nk, kmodes, F, ell, xi, r);
  }
  CUDA_TRY(cudaGetLastError());
  return DEB_OK;
}

int deb_spectra_host_f64(int32_t device, int32_t nk, int32_t nmu, const double* y, const double* kmodes, double As, double ns, double kp, double bias,
                         const double* sg_coef, int32_t sg_window, const double* mu, int32_t ell,
                         double* P0, double* P2, double* P4, double* Pkmu, double* Ps_delta, double* Ps_theta, double* xi, double* r) {
  if (nk < 2 || !y || !kmodes || !P0 || !P2 || !P4) return DEB_E_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return DEB_E_NODEVICE;
  CUDA_TRY(cudaSetDevice(device));
  const size_t n = nk, wsb = deb_spectra_workspace_bytes(nk);
  const size_t b_y = n * 20 * 8, b_k = n * 8, b_c = (size_t)(sg_window > 0 ? sg_window : 1) * 8, b_mu = (size_t)(nmu > 0 ? nmu : 1) * 8,
               b_kmu = (size_t)(nmu > 0 ? n * nmu : 1) * 8;
  char* d = nullptr;
  const size_t total = b_y + b_k + b_c + b_mu + 7 * b_k + b_kmu + wsb + 4096;
  CUDA_TRY(cudaMalloc((void**)&d, total));
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  char* p = d;
  double* d_y = (double*)p; p += al(b_y); double* d_k = (double*)p; p += al(b_k); double* d_c = (double*)p; p += al(b_c); double* d_mu = (double*)p; p += al(b_mu);
  double* o[7];
  for (int i = 0; i < 7; ++i) { o[i] = (double*)p; p += al(b_k); }
  double* d_kmu = (double*)p; p += al(b_kmu);
  void* d_ws = p;
  int rc = DEB_OK;
  cudaMemcpy(d_y, y, b_y, cudaMemcpyHostToDevice); cudaMemcpy(d_k, kmodes, b_k, cudaMemcpyHostToDevice);
  if (sg_window > 0) cudaMemcpy(d_c, sg_coef, b_c, cudaMemcpyHostToDevice);
  if (nmu > 0 && mu) cudaMemcpy(d_mu, mu, b_mu, cudaMemcpyHostToDevice);
  rc = deb_spectra_f64(nk, nmu, d_y, d_k, As, ns, kp, bias, sg_window > 0 ? d_c : nullptr, sg_window, (nmu > 0 && mu) ? d_mu : nullptr, ell, o[0], o[1], o[2],
                       (nmu > 0 && Pkmu) ? d_kmu : nullptr, o[3], o[4], (xi && r) ? o[5] : nullptr, (xi && r) ? o[6] : nullptr, d_ws, wsb, nullptr);
  if (rc == DEB_OK && cudaDeviceSynchronize() != cudaSuccess) rc = DEB_E_CUDA;
  if (rc == DEB_OK) {
    cudaMemcpy(P0, o[0], b_k, cudaMemcpyDeviceToHost); cudaMemcpy(P2, o[1], b_k, cudaMemcpyDeviceToHost); cudaMemcpy(P4, o[2], b_k, cudaMemcpyDeviceToHost);
    if (sg_window > 0) { cudaMemcpy(Ps_delta, o[3], b_k, cudaMemcpyDeviceToHost); cudaMemcpy(Ps_theta, o[4], b_k, cudaMemcpyDeviceToHost); }
    if (nmu > 0 && Pkmu) cudaMemcpy(Pkmu, d_kmu, b_kmu, cudaMemcpyDeviceToHost);
    if (xi && r) { cudaMemcpy(xi, o[5], b_k, cudaMemcpyDeviceToHost); cudaMemcpy(r, o[6], b_k, cudaMemcpyDeviceToHost); }
  }
  cudaFree(d);
  return rc;
}

}  // extern "C"
