// deb_spectra.cuh -- spectra post-processing of the solver output: what the reference's host API
// computes from `y` right after evolve_perturbations (/root/reference/src/discoeb/perturbations.py):
//   power_multipoles   :1202-1224   P0, P2, P4 of the Kaiser spectrum
//   power_Kaiser       :1162-1199   P(k, mu) = (b delta_m - mu^2 theta_m)^2, optionally from Savitzky-Golay smoothed spectra
//   get_power_smoothed :1126-1160   exp(SG filter of log P) with the raw signal kept in the half-windows at both ends
//                                   (util.savgol_filter :407-444: zero-padded 'same' convolution with the centre weights)
//   get_xi_from_P      :1065-1098   FFTlog (Talman 1978, Hamilton 2000): rfft, phase from Im log Gamma (util.py:12-44, Lanczos), irfft
// Element functions shared by the CUDA kernels (deb_spectra.cu) and the CPU build (tests/emu: test infrastructure).
// The DFTs are evaluated directly (N <= 4096 modes: <= 8 M complex multiply-adds, one thread per output bin) with
// sincospi on the exactly reduced phase (m n mod N) / N.
#pragma once
#include <math.h>

namespace deb {
namespace sp {

#ifdef DEB_CPU_EMU
#define SP_DEV inline
inline void sp_sincospi(double x, double* s, double* c) { *s = sin(M_PI * x); *c = cos(M_PI * x); }
#else
#define SP_DEV __host__ __device__ __forceinline__
SP_DEV void sp_sincospi(double x, double* s, double* c) { sincospi(x, s, c); }
#endif

struct Cx { double re, im; };
SP_DEV Cx cx(double a, double b) { Cx r; r.re = a; r.im = b; return r; }
SP_DEV Cx cadd(Cx a, Cx b) { return cx(a.re + b.re, a.im + b.im); }
SP_DEV Cx csub(Cx a, Cx b) { return cx(a.re - b.re, a.im - b.im); }
SP_DEV Cx cmul(Cx a, Cx b) { return cx(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
SP_DEV Cx clog(Cx a) { return cx(log(hypot(a.re, a.im)), atan2(a.im, a.re)); }
SP_DEV Cx csinpi(Cx z) {        // sin(pi z)
  double s, c;
  sp_sincospi(z.re, &s, &c);
  const double y = M_PI * z.im;
  return cx(s * cosh(y), c * sinh(y));
}

// amplitude sqrt(2 pi^2 A_s (k/k_p)^(n_s-1) k^-3)
SP_DEV double amp_of_k(double k, double As, double ns, double kp) { return sqrt(2.0 * M_PI * M_PI * As * pow(k / kp, ns - 1.0) * pow(k, -3.0)); }

// log Gamma(z), complex z: Lanczos (g = 7) with reflection for Re z <= 1/2 (util.py:12-44, after GSL)
SP_DEV Cx lngamma_lanczos(Cx z) {
  const double c[9] = {0.99999999999980993227684700473478, 676.520368121885098567009190444019, -1259.13921672240287047156078755283,
                       771.3234287776530788486528258894, -176.61502916214059906584551354, 12.507343278686904814458936853,
                       -0.13857109526572011689554707, 9.984369578019570859563e-6, 1.50563273514931155834e-7};
  z.re -= 1.0;
  Cx Ag = cx(c[0], 0.0);
  for (int i = 1; i <= 8; ++i) {
    const double tr = z.re + i, ti = z.im, n2 = tr * tr + ti * ti;
    Ag.re += c[i] * tr / n2; Ag.im -= c[i] * ti / n2;         // c_i conj(t) / |t|^2
  }
  const Cx a = cmul(cx(z.re + 0.5, z.im), clog(cx(z.re + 7.5, z.im)));
  return cadd(csub(a, cx(z.re + 7.5, z.im)), cadd(cx(0.9189385332046727418, 0.0), clog(Ag)));
}
SP_DEV Cx lngamma_complex(Cx z) {
  if (z.re <= 0.5) {
    const Cx l = lngamma_lanczos(cx(1.0 - z.re, -z.im));
    return csub(csub(cx(1.14472988584940017414342735135, 0.0), clog(csinpi(z))), l);
  }
  return lngamma_lanczos(z);
}

// ---- per-element pieces ----
// (1) point-wise: amplitudes, multipoles, log spectra for the smoother
SP_DEV void point(int i, const double* y, const double* k, double As, double ns, double kp, double bias, double* dm, double* tm, double* P0, double* P2,
                  double* P4, double* lPd, double* lPt, double* Pdraw) {
  const double a = amp_of_k(k[i], As, ns, kp);
  const double d = a * y[(size_t)i * 20 + 4], t = a * y[(size_t)i * 20 + 5];
  dm[i] = d; tm[i] = t;
  P0[i] = bias * bias * d * d - 2.0 * bias / 3.0 * d * t + 1.0 / 5.0 * t * t;
  P2[i] = -4.0 * bias / 3.0 * d * t + 4.0 / 7.0 * t * t;
  P4[i] = 8.0 / 35.0 * t * t;
  const double Pd = 2.0 * M_PI * M_PI * As * pow(k[i] / kp, ns - 1.0) * pow(k[i], -3.0) * y[(size_t)i * 20 + 4] * y[(size_t)i * 20 + 4];
  const double Pt = 2.0 * M_PI * M_PI * As * pow(k[i] / kp, ns - 1.0) * pow(k[i], -3.0) * y[(size_t)i * 20 + 5] * y[(size_t)i * 20 + 5];
  lPd[i] = log(Pd); lPt[i] = log(Pt); Pdraw[i] = Pd;
}
// (2) Savitzky-Golay: zero-padded 'same' convolution of log P with w centre weights, exp, raw signal in the end zones
SP_DEV double smooth_at(int i, int n, const double* lP, const double* coef, int w) {
  const int h = w / 2;
  const int hi_raw = (w + 1) / 2;                    // the reference's Pms.at[-w//2:] is floor(-w/2): (w+1)/2 samples for odd w
  if (i < h || i >= n - hi_raw) return exp(lP[i]);
  // numpy.convolve(y, c, 'same')[i] = sum_j c[j] y[i + (w-1)/2 - j]
  double acc = 0.0;
  const int off = (w - 1) / 2;
  for (int j = 0; j < w; ++j) { const int q = i + off - j; if (q >= 0 && q < n) acc += coef[j] * lP[q]; }
  return exp(acc);
}
// (3) FFTlog forward: F[m] = rfft(P k^1.5)[m] * exp(2i (Im lnGamma(z_m) - log(pi) ki_m)), m = 0 .. N/2
SP_DEV Cx fftlog_forward(int m, int N, const double* k, const double* Pk, int ell) {
  double fr = 0.0, fi = 0.0;
  for (int nn = 0; nn < N; ++nn) {
    const double x = Pk[nn] * k[nn] * sqrt(k[nn]);
    double s, c;
    sp_sincospi(2.0 * (double)(((long long)m * nn) % N) / (double)N, &s, &c);
    fr += x * c; fi -= x * s;
  }
  const double L = log(k[N - 1] / k[0]);
  const double ki = M_PI * m / L;
  const double theta = lngamma_complex(cx((1.5 + ell) / 2.0, ki)).im;
  const double ph = 2.0 * (theta - log(M_PI) * ki);
  return cmul(cx(fr, fi), cx(cos(ph), sin(ph)));
}
// (4) FFTlog backward: xi at r_n = 2 pi / k_n (irfft of F with n = N), still in the order of k
SP_DEV double fftlog_backward(int nn, int N, const double* k, const Cx* F, int ell) {
  const int M = N / 2;
  double acc = F[0].re;
  const int last = (N % 2 == 0) ? M - 1 : M;
  for (int m = 1; m <= last; ++m) {
    double s, c;
    sp_sincospi(2.0 * (double)(((long long)m * nn) % N) / (double)N, &s, &c);
    acc += 2.0 * (F[m].re * c - F[m].im * s);
  }
  if (N % 2 == 0) acc += (nn % 2 == 0 ? 1.0 : -1.0) * F[M].re;
  acc /= (double)N;
  const double r = 2.0 * M_PI / k[nn];
  const double sgn = (ell % 4 == 0) ? 1.0 : ((ell % 4 == 2) ? -1.0 : 0.0);       // Re(i^ell)
  return sgn * acc / pow(2.0 * M_PI * r, 1.5);
}

}  // namespace sp
}  // namespace deb
