// deb_background.cuh -- table production for the hot path: what the reference's
// evolve_background(param, thermo_module='RECFAST') leaves in `param` for evolve_perturbations, computed on the GPU for a
// batch of cosmologies, one CTA per cosmology.  Citations are into /root/reference/src/discoeb/.
//
//   setup_background_evolution   background.py:140-188   densities, neutrino rho/p tables (512 knots, 8-point
//                                                        generalised Gauss-Laguerre rule util.py:82-123), OmegaDE, taumin
//   compute_thermal_history      thermodynamics_recfast.py:362-452   adaptive a-grid (:324-360), Saha switches, and per
//                                                        interval the stiff solve of `ionization` (:124-281) by
//   GRKT4.step + diffrax PID     ode_integrators_stiff.py:77-215, thermodynamics_recfast.py:283-299 (rtol 1e-3, atol 1e-6,
//                                                        dt0 = 1e-3 a, I-controller of order 3, <= 128 steps)
//   evaluate_thermo              thermodynamics_recfast.py:455-500   HeII Saha (:302-321) and its derivative, x_e, c_s^2,
//                                                        tau(a) by Romberg sums per interval (jax_cosmo romb, 65 points)
//   spline constructor           spline_interpolation.py:8-113   natural cubic splines (Thomas algorithm)
//
// jax.jacfwd / jax.grad are forward duals here (DN<N>: a value and N tangents in one pass).  Optical depth and visibility
// (background.py:300-342) are not inputs of evolve_perturbations and are not produced.  The physical constants and the
// fitting formulae are the reference's (RECFAST); they admit no other form.
//
// Compiles for the device and, with -DDEB_CPU_EMU, as plain C++ (tests/emu: test infrastructure).
#pragma once
#include <math.h>
#include <stdint.h>

namespace deb {
namespace bg {

#ifdef DEB_CPU_EMU
#define BG_DEV inline
#define BG_HD inline
#else
#define BG_DEV __device__ __forceinline__
#define BG_HD __host__ __device__ inline
#endif

constexpr int NBGIN = 16;     // doubles per cosmology in the input block (order below)
enum { BI_OMEGAM = 0, BI_OMEGAB, BI_OMEGAK, BI_W0, BI_WA, BI_CS2DE, BI_H0, BI_TCMB, BI_YHE, BI_NEFF, BI_NMNU, BI_MNU, BI_AS, BI_NS, BI_KP };
constexpr int NNUQ = 8;       // momentum nodes of nu_background (background.py:48)
constexpr double AMIN = 1e-9, AMAX = 1.01;          // background.py:222-223

// ---- forward duals: value + N tangents ----
template <int N> struct DN { double v; double d[N]; };
template <int N> BG_DEV DN<N> dn_const(double v) { DN<N> r; r.v = v; for (int i = 0; i < N; ++i) r.d[i] = 0.0; return r; }
template <int N> BG_DEV DN<N> operator+(DN<N> a, DN<N> b) { DN<N> r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> BG_DEV DN<N> operator-(DN<N> a, DN<N> b) { DN<N> r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> BG_DEV DN<N> operator*(DN<N> a, DN<N> b) { DN<N> r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int N> BG_DEV DN<N> operator/(DN<N> a, DN<N> b) { DN<N> r; const double ib = 1.0 / b.v; r.v = a.v * ib; for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib; return r; }
template <int N> BG_DEV DN<N> operator+(DN<N> a, double b) { a.v += b; return a; }
template <int N> BG_DEV DN<N> operator+(double b, DN<N> a) { a.v += b; return a; }
template <int N> BG_DEV DN<N> operator-(DN<N> a, double b) { a.v -= b; return a; }
template <int N> BG_DEV DN<N> operator-(double b, DN<N> a) { DN<N> r; r.v = b - a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
template <int N> BG_DEV DN<N> operator-(DN<N> a) { DN<N> r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
template <int N> BG_DEV DN<N> operator*(DN<N> a, double b) { a.v *= b; for (int i = 0; i < N; ++i) a.d[i] *= b; return a; }
template <int N> BG_DEV DN<N> operator*(double b, DN<N> a) { return a * b; }
template <int N> BG_DEV DN<N> operator/(DN<N> a, double b) { return a * (1.0 / b); }
template <int N> BG_DEV DN<N> operator/(double b, DN<N> a) { DN<N> r; const double ia = 1.0 / a.v; r.v = b * ia; for (int i = 0; i < N; ++i) r.d[i] = -r.v * a.d[i] * ia; return r; }
template <int N> BG_DEV DN<N> xsqrt(DN<N> a) { DN<N> r; r.v = sqrt(a.v); const double h = 0.5 / r.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * h; return r; }
template <int N> BG_DEV DN<N> xexp(DN<N> a) { DN<N> r; r.v = exp(a.v); for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * r.v; return r; }
template <int N> BG_DEV DN<N> xlog(DN<N> a) { DN<N> r; r.v = log(a.v); const double ia = 1.0 / a.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * ia; return r; }
template <int N> BG_DEV DN<N> xpow(DN<N> a, double p) { DN<N> r; r.v = pow(a.v, p); const double f = p * r.v / a.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * f; return r; }
template <int N> BG_DEV DN<N> xabs(DN<N> a) { return a.v < 0.0 ? -a : a; }
template <int N> BG_DEV double val(const DN<N>& a) { return a.v; }
BG_DEV double xsqrt(double a) { return sqrt(a); }
BG_DEV double xexp(double a) { return exp(a); }
BG_DEV double xlog(double a) { return log(a); }
BG_DEV double xpow(double a, double p) { return pow(a, p); }
BG_DEV double xabs(double a) { return fabs(a); }
BG_DEV double val(double a) { return a; }
template <class T> BG_DEV T xsel(bool c, T a, T b) { return c ? a : b; }

// ---- physical constants (thermodynamics_recfast.py:36-122) ----
#define BGC_G 6.67430e-11
#define BGC_MH 1.67353284e-27
#define BGC_ME 9.1093837015e-31
#define BGC_MHE 3.97146570884
#define BGC_C 2.99792458e+08
#define BGC_HP 6.62607015e-34
#define BGC_KB 1.380649e-23
#define BGC_SIGT 6.6524587321e-29
#define BGC_ARAD (4 * 5.670374419e-8 / BGC_C)
#define BGC_BIGH 3.2407792902755102e-18
#define BGC_DENSFAC 11.223810928601939
#define BGC_C2OK 1.62581581e4
#define BGC_EHE2 6.314878282674e5
#define BGC_LAM_H 8.2245809
#define BGC_LAM_HE 51.3
#define BGC_L_HION 1.096787737e7
#define BGC_L_HA 8.225916453e6
#define BGC_L_HE1 1.98310772e7
#define BGC_L_HE2S 1.66277434e7
#define BGC_L_HE2P 1.71134891e7
#define BGC_A2PS 1.798287e9
#define BGC_A2PT 177.58
#define BGC_L_HE2PT 1.690871466e7
#define BGC_L_HE2ST 1.5985597526e7
#define BGC_L_HE2ST_ION 3.8454693845e6
#define BGC_SIG2PS 1.436289e-22
#define BGC_SIG2PT 1.484872e-22
#define BGC_PI 3.141592653589793

struct Consts {        // derived once per thread (cheap) -- thermodynamics_recfast.py:88-122
  double CDB, CDB_He, CB1, CB1_He1, CR, CK, CK_He, CL, CL_He, CT, Bfact, CL_PSt, hcL2St, hcL2Stion, aVF, T0, T1, atrip;
};
BG_DEV Consts make_consts() {
  Consts c;
  const double lam_a = 1.0 / BGC_L_HA, lam_aHe = 1.0 / BGC_L_HE2P;
  c.CDB = BGC_HP * BGC_C * (BGC_L_HION - BGC_L_HA) / BGC_KB;
  c.CDB_He = BGC_HP * BGC_C * (BGC_L_HE1 - BGC_L_HE2S) / BGC_KB;
  c.CB1 = BGC_HP * BGC_C * BGC_L_HION / BGC_KB;
  c.CB1_He1 = BGC_HP * BGC_C * BGC_L_HE1 / BGC_KB;
  c.CR = 2.0 * BGC_PI * (BGC_ME / BGC_HP) * (BGC_KB / BGC_HP);
  c.CK = lam_a * lam_a * lam_a / (8.0 * BGC_PI);
  c.CK_He = lam_aHe * lam_aHe * lam_aHe / (8.0 * BGC_PI);
  c.CL = BGC_C * BGC_HP / (BGC_KB * lam_a);
  c.CL_He = BGC_C * BGC_HP / (BGC_KB / BGC_L_HE2S);
  c.CT = (8.0 / 3.0) * (BGC_SIGT / (BGC_ME * BGC_C)) * BGC_ARAD;
  c.Bfact = BGC_HP * BGC_C * (BGC_L_HE2P - BGC_L_HE2S) / BGC_KB;
  c.CL_PSt = BGC_HP * BGC_C * (BGC_L_HE2PT - BGC_L_HE2ST) / BGC_KB;
  c.hcL2St = BGC_HP * BGC_C * BGC_L_HE2ST / BGC_KB;
  c.hcL2Stion = BGC_HP * BGC_C * BGC_L_HE2ST_ION / BGC_KB;
  c.aVF = pow(10.0, -16.744); c.T0 = pow(10.0, 0.477121); c.T1 = pow(10.0, 5.114); c.atrip = pow(10.0, -16.306);
  return c;
}

// per-cosmology scalars and the neutrino density spline every later phase reads
struct BgP {
  double Omegam, Omegab, Omegak, w0, wa, H0, Tcmb, YHe, Neff, Nmnu, mnu;
  double grhom, grhog, grhor, adotrad, amnu, OmegaDE, taumin, fHe;
  const double* lx; const double* ly; const double* lS; int nnu;       // log rho_nu (log a)
};

// natural cubic spline value (spline_interpolation.py:130-153)
BG_DEV double spline_eval(const double* x, const double* y, const double* S, int n, double xn) {
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (x[mid] < xn) lo = mid + 1; else hi = mid; }
  int i = lo - 1;
  if (i < 0) i = 0;
  if (i > n - 2) i = n - 2;
  const double h = x[i + 1] - x[i], t = (xn - x[i]) / h, A = 1.0 - t, B = t;
  return A * y[i] + B * y[i + 1] + ((A * A * A - A) * S[i] + (B * B * B - B) * S[i + 1]) * (h * h) / 6.0;
}
// first and second derivative at knot i of the spline itself, with the interval the reference's searchsorted picks for
// x_new = x[i] (spline_interpolation.py:203-260: idx = clip(searchsorted(x, x_i) - 1, 0, n-2) = i-1, the interval to the LEFT)
BG_DEV void spline_knot_derivs(const double* x, const double* y, const double* S, int n, int i, double* d1, double* d2) {
  int k = i - 1;
  if (k < 0) k = 0;
  if (k > n - 2) k = n - 2;
  const double h = x[k + 1] - x[k], dv = x[i] - x[k];
  const double b = (y[k + 1] - y[k]) / h - h * (S[k + 1] + 2.0 * S[k]) / 6.0, c = S[k] / 2.0, d = (S[k + 1] - S[k]) / (6.0 * h);
  *d1 = b + 2.0 * c * dv + 3.0 * d * dv * dv;
  *d2 = 2.0 * c + 6.0 * d * dv;
}
// integral of the spline from x[0] to xn given the cumulative knot integrals (spline_interpolation.py:95-104, 155-189)
BG_DEV double spline_interval_integral(const double* x, const double* y, const double* S, int k, double dv) {
  const double h = x[k + 1] - x[k];
  const double b = (y[k + 1] - y[k]) / h - h * (S[k + 1] + 2.0 * S[k]) / 6.0, c = S[k] / 2.0, d = (S[k + 1] - S[k]) / (6.0 * h);
  return y[k] * dv + b * (dv * dv) / 2.0 + c * (dv * dv * dv) / 3.0 + d * (dv * dv * dv * dv) / 4.0;
}
BG_DEV double spline_integral_from_start(const double* x, const double* y, const double* S, const double* icum, int n, double xn) {
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (x[mid] < xn) lo = mid + 1; else hi = mid; }
  int k = lo - 1;
  if (k < 0) k = 0;
  if (k > n - 2) k = n - 2;
  return icum[k] + spline_interval_integral(x, y, S, k, xn - x[k]);
}
// natural cubic spline second derivatives by the Thomas algorithm (spline_interpolation.py:45-86); cp, dp: scratch [n]
BG_DEV void spline_build(const double* x, const double* y, double* S, int n, double* cp, double* dp) {
  S[0] = 0.0; S[n - 1] = 0.0;
  const int m = n - 2;
  if (m <= 0) return;
  // interior unknown j (0-based) is S[j+1]; a_j = h_j, b_j = 2 (h_j + h_{j+1}), c_j = h_{j+1}
  {
    const double h0 = x[1] - x[0], h1 = x[2] - x[1];
    const double b = 2.0 * (h0 + h1), d = 6.0 * ((y[2] - y[1]) / h1 - (y[1] - y[0]) / h0);
    cp[0] = h1 / b; dp[0] = d / b;
  }
  for (int j = 1; j < m; ++j) {
    const double h0 = x[j + 1] - x[j], h1 = x[j + 2] - x[j + 1];
    const double b = 2.0 * (h0 + h1), d = 6.0 * ((y[j + 2] - y[j + 1]) / h1 - (y[j + 1] - y[j]) / h0);
    const double denom = b - h0 * cp[j - 1];
    cp[j] = (j < m - 1) ? h1 / denom : 0.0;
    dp[j] = (d - h0 * dp[j - 1]) / denom;
  }
  double s = dp[m - 1];
  S[m] = s;
  for (int j = m - 2; j >= 0; --j) { s = dp[j] - cp[j] * s; S[j + 1] = s; }
}

// conformal Hubble rate a'/a (background.py:100-122) and d tau / d a (background.py:75-85)
BG_DEV double grho2_of_a(const BgP& p, double a) {           // 8 pi G rho a^4 ("grho2" of background.py:80)
  const double rhonu = exp(spline_eval(p.lx, p.ly, p.lS, p.nnu, log(a)));
  const double rho_de = pow(a, -3.0 * (1.0 + p.w0 + p.wa)) * exp(3.0 * (a - 1.0) * p.wa);
  return p.grhom * p.Omegam * a + (p.grhog + p.grhor * (p.Neff + p.Nmnu * rhonu)) + p.grhom * p.OmegaDE * rho_de * (a * a * a * a)
       + p.grhom * p.Omegak * (a * a);
}
BG_DEV double dtauda(const BgP& p, double a) { return sqrt(3.0 / grho2_of_a(p, a)); }
BG_DEV double aprimeoa(const BgP& p, double a) {
  const double rhonu = exp(spline_eval(p.lx, p.ly, p.lS, p.nnu, log(a)));
  const double rho_de = pow(a, -3.0 * (1.0 + p.w0 + p.wa)) * exp(3.0 * (a - 1.0) * p.wa);
  const double grho = p.grhom * p.Omegam / a + (p.grhog + p.grhor * (p.Neff + p.Nmnu * rhonu)) / (a * a) + p.grhom * p.OmegaDE * rho_de * (a * a)
                    + p.grhom * p.Omegak;
  return sqrt(grho / 3.0);
}
// Romberg integration of d tau/d a on [lo, hi], 2^6 + 1 points (jax_cosmo.scipy.integrate.romb as called at
// background.py:182 and thermodynamics_recfast.py:484: trapezoid sums by interval halving + Richardson extrapolation)
BG_DEV double romb_dtauda(const BgP& p, double lo, double hi) {
  constexpr int DIVMAX = 6;
  double st[DIVMAX + 1], nw[DIVMAX + 1];
  const double range = hi - lo;
  double ordsum = 0.5 * (dtauda(p, lo) + dtauda(p, hi));
  for (int i = 0; i <= DIVMAX; ++i) st[i] = range * ordsum;
  for (int i = 1; i <= DIVMAX; ++i) {
    const int n = 1 << i, half = n >> 1;
    const double h = range / half, x0 = lo + 0.5 * h;
    double s = 0.0;
    for (int q = 0; q < half; ++q) s += dtauda(p, x0 + h * q);
    ordsum += s;
    double x = range * ordsum / n;
    nw[0] = x;
    double f4 = 1.0;
    for (int kk = 0; kk < i; ++kk) { f4 *= 4.0; x = (f4 * x - st[kk]) / (f4 - 1.0); nw[kk + 1] = x; }
    for (int kk = 0; kk <= i; ++kk) st[kk] = nw[kk];
  }
  return st[DIVMAX];
}

// neutrino density and pressure of one massive flavour in units of a massless one (background.py:48-72)
BG_DEV void nu_background(double a, double amnu, const double* q, const double* w, double* rho, double* pres, double* ppseudo = nullptr) {
  double r = 0.0, pn = 0.0, pp = 0.0;
  for (int i = 0; i < NNUQ; ++i) {
    const double aq = a * amnu / q[i];
    const double v = 1.0 / sqrt(1.0 + aq * aq);
    r += w[i] * (1.0 / v);
    pn += w[i] * (v / 3.0);
    pp += w[i] * (v * v * v / 3.0);
  }
  *rho = r; *pres = pn;
  if (ppseudo) *ppseudo = pp;
}

// adaptive sampling of the scale factor (thermodynamics_recfast.py:324-360): point i of N
BG_DEV double geom_point(double lo, double hi, int i, int num, bool endpoint) {
  const int div = endpoint ? num - 1 : num;
  if (endpoint && i == num - 1) return hi;
  if (i == 0) return lo;
  const double l0 = log10(lo), l1 = log10(hi), step = (l1 - l0) / (double)div;
  return pow(10.0, (double)i * step + l0);
}
BG_DEV double adaptive_a(double a0, double a1, int N, int i) {
  int n1 = (int)(N * 0.05); if (n1 < 8) n1 = 8;
  int n2 = (int)(N * 0.10); if (n2 < 8) n2 = 8;
  const int n3 = (int)(N * 0.50);
  const int n4 = N - n1 - n2 - n3;
  double b1 = 1.0 / 3001.0; if (b1 < a0) b1 = a0;
  const double b2 = 1.0 / 1401.0, b3 = 1.0 / 601.0;
  if (i < n1) return geom_point(a0, b1, i, n1, false);
  if (i < n1 + n2) return geom_point(b1, b2, i - n1, n2, false);
  if (i < n1 + n2 + n3) return geom_point(b2, b3, i - n1 - n2, n3, false);
  return geom_point(b3, a1, i - n1 - n2 - n3, n4, true);
}

// ---- RECFAST right-hand side d(x_H, x_He, T_m)/da (thermodynamics_recfast.py:124-281) ----
template <class T>
BG_DEV void ionization(const BgP& p, const Consts& c, double a, const T* y, T* f) {
  const double z1 = 1.0 / a;                      // 1 + z
  const double hh = p.H0 / 100.0, HO = hh * BGC_BIGH, fu = 1.105, fHe = p.fHe;
  const double Nnow = BGC_DENSFAC * hh * hh * p.Omegab * (1.0 - p.YHe);
  const T xH = y[0], xHe = y[1];
  const T x = xH + fHe * xHe;
  const T Tm = xabs(y[2]);
  const double n = Nnow * z1 * z1 * z1, nHe = fHe * n, Tr = p.Tcmb * z1;
  const double Hz = (1e-5 * aprimeoa(p, a)) / a * BGC_C * BGC_BIGH;
  const T T4 = Tm / 1e4;
  const T crt = xpow(c.CR * Tm, 1.5);
  const T Rdn = 1e-19 * 4.309 * xpow(T4, -0.6166) / (1.0 + 0.6703 * xpow(T4, 0.5300));
  const T Rup = Rdn * crt * xexp(-(c.CDB / Tm));
  const T s0 = xsqrt(Tm / c.T0), s1 = xsqrt(Tm / c.T1);
  const T s0t = 1.0 + s0, s1t = 1.0 + s1;
  const T RdnHe = c.aVF / (s0 * xpow(s0t, 1.0 - 0.711) * xpow(s1t, 1.0 + 0.711));
  const T RupHe = 4.0 * RdnHe * crt * xexp(-(c.CDB_He / Tm));
  T heb_arg = c.Bfact / Tm;
  if (val(heb_arg) > 680.0) heb_arg = 0.0 * heb_arg + 680.0;
  const T HeB = xexp(heb_arg);
  const T Rdn_t = c.atrip / (s0 * xpow(s0t, 1.0 - 0.761) * xpow(s1t, 1.0 + 0.761));
  const T Rup_t = Rdn_t * xexp(-(c.hcL2Stion / Tm)) * crt * (4.0 / 3.0);
  const double lz = log(z1);
  const double g1 = (lz - 7.28) / 0.18, g2 = (lz - 6.75) / 0.33;
  const double K = c.CK / Hz * (1.0 + (-0.14) * exp(-(g1 * g1)) + 0.05 * exp(-(g2 * g2)));
  const T omHe = 1.0 - xHe, omH = 1.0 - xH;
  const T nHe1 = nHe * omHe;
  const T tauHe_s = BGC_A2PS * c.CK_He * 3.0 * nHe1 / Hz;
  const T pHe_s = (1.0 - xexp(-tauHe_s)) / tauHe_s;
  const T dop = xsqrt(2.0 * BGC_KB * Tm / (BGC_MH * BGC_MHE * BGC_C * BGC_C));
  const double cl2p = BGC_C * BGC_L_HE2P, cl2pt = BGC_C * BGC_L_HE2PT;
  const T g2Ps = (3.0 * BGC_A2PS * fHe * omHe * (BGC_C * BGC_C)) / (sqrt(BGC_PI) * BGC_SIG2PS * 8.0 * BGC_PI * (cl2p * dop) * omH * (cl2p * cl2p));
  const T AHcon = BGC_A2PS / (1.0 + 0.36 * xpow(g2Ps, 0.86));
  const bool he_edge = val(xHe) < 5e-9 || val(xHe) > 0.98;
  T KHe;
  if (he_edge) KHe = 0.0 * xH + c.CK_He / Hz;
  else if (val(xH) < 0.9999999) KHe = 1.0 / ((BGC_A2PS * pHe_s + AHcon) * 3.0 * nHe1);
  else KHe = 1.0 / (BGC_A2PS * pHe_s * 3.0 * nHe1);
  const T tauHe_t = BGC_A2PT * nHe1 * 3.0 / (8.0 * BGC_PI * Hz * (BGC_L_HE2PT * BGC_L_HE2PT * BGC_L_HE2PT));
  const T pHe_t = (1.0 - xexp(-tauHe_t)) / tauHe_t;
  const T g2Pt = (3.0 * BGC_A2PT * fHe * omHe * (BGC_C * BGC_C)) / (sqrt(BGC_PI) * BGC_SIG2PT * 8.0 * BGC_PI * (cl2pt * dop) * omH * (cl2pt * cl2pt));
  const T AHcon_t = BGC_A2PT / (1.0 + 0.66 * xpow(g2Pt, 0.9)) / 3.0;
  const T epst = xexp(-(c.CL_PSt / Tm));
  T Cf = val(xH) > 0.99999 ? BGC_A2PT * pHe_t * epst : (BGC_A2PT * pHe_t + AHcon_t) * epst;
  Cf = Cf / (Rup_t + Cf);
  const T timeTh = (1.0 / (c.CT * (Tr * Tr * Tr * Tr))) * (1.0 + x + fHe) / x;
  const double timeH = 2.0 / (3.0 * HO * pow(z1, 1.5));
  const double Hzz = Hz * z1;
  const T rd = x * xH * n * Rdn - Rup * omH * xexp(-(c.CL / Tm));
  T f0;
  if (val(xH) > 0.99) f0 = 0.0 * xH;
  else if (val(xH) > 0.985) f0 = rd / Hzz;
  else {
    const T KL = K * BGC_LAM_H * n * omH;
    f0 = (rd * (1.0 + KL)) / (Hzz * (1.0 / fu + KL / fu + K * Rup * n * omH));
  }
  T f1;
  if (val(xHe) < 1e-8) f1 = 0.0 * xHe;
  else {
    const T rdHe = x * xHe * n * RdnHe - RupHe * omHe * xexp(-(c.CL_He / Tm));
    const T KLHe = KHe * BGC_LAM_HE * nHe1 * HeB;
    f1 = (rdHe * (1.0 + KLHe)) / (Hzz * (1.0 + KHe * (BGC_LAM_HE + RupHe) * nHe1 * HeB));
    if (!he_edge) {
      const T tr = x * xHe * n * Rdn_t - omHe * 3.0 * Rup_t * xexp(-(c.hcL2St / Tm));
      f1 = f1 + tr * Cf / Hzz;
    }
  }
  T f2;
  if (val(timeTh) < 1e-3 * timeH) f2 = Tm / z1;
  else f2 = c.CT * (Tr * Tr * Tr * Tr) * x / (1.0 + x + fHe) * (Tm - Tr) / Hzz + 2.0 * Tm / z1;
  const double dzda = -1.0 / (a * a);
  f[0] = f0 * dzda; f[1] = f1 * dzda; f[2] = f2 * dzda;
}

// 3x3 pivoted LU (what jax.scipy.linalg.lu_factor / lu_solve do at ode_integrators_stiff.py:185-210)
struct LU3 { double m[3][3]; int piv[3]; };
BG_DEV void lu3_factor(LU3& L) {
  for (int k = 0; k < 3; ++k) {
    int pk = k; double best = fabs(L.m[k][k]);
    for (int r = k + 1; r < 3; ++r) if (fabs(L.m[r][k]) > best) { best = fabs(L.m[r][k]); pk = r; }
    L.piv[k] = pk;
    if (pk != k) for (int cc = 0; cc < 3; ++cc) { const double t = L.m[k][cc]; L.m[k][cc] = L.m[pk][cc]; L.m[pk][cc] = t; }
    for (int r = k + 1; r < 3; ++r) {
      L.m[r][k] /= L.m[k][k];
      for (int cc = k + 1; cc < 3; ++cc) L.m[r][cc] -= L.m[r][k] * L.m[k][cc];
    }
  }
}
BG_DEV void lu3_solve(const LU3& L, const double* b, double* x) {
  double v[3] = {b[0], b[1], b[2]};
  for (int k = 0; k < 3; ++k) { const int pk = L.piv[k]; if (pk != k) { const double t = v[k]; v[k] = v[pk]; v[pk] = t; } }
  v[1] -= L.m[1][0] * v[0];
  v[2] -= L.m[2][0] * v[0] + L.m[2][1] * v[1];
  v[2] /= L.m[2][2];
  v[1] = (v[1] - L.m[1][2] * v[2]) / L.m[1][1];
  v[0] = (v[0] - L.m[0][1] * v[1] - L.m[0][2] * v[2]) / L.m[0][0];
  x[0] = v[0]; x[1] = v[1]; x[2] = v[2];
}

// one GRKT4 step (Kaps-Rentrop GRK4T; ode_integrators_stiff.py:122-215): y0 at t0 -> y1, error estimate.  All stage
// evaluations are at t0, as in the reference (the "t1, y1 = t0, ..." assignments there).
BG_DEV void grkt4_step(const BgP& p, const Consts& c, double t0, double dt, const double* y0, double* y1, double* err) {
  const double gamma = 0.395, g21 = -0.767672395484, g31 = -0.851675323742, g32 = 0.522967289188, g41 = 0.288463109545,
               g42 = 0.880214273381e-1, g43 = -0.337389840627, a21 = 0.438, a31 = 0.796920457938, a32 = 0.730795420615e-1,
               ch1 = 0.346325833758, ch2 = 0.285693175712, ch3 = 0.367980990530, c1 = 0.199293275701, c2 = 0.482645235674,
               c3 = 0.680614886256e-1, c4 = 0.25;
  // Jacobian of f dt at (t0, y0) by one pass of 3-tangent duals, and f itself
  DN<3> yd[3], fd[3];
  for (int i = 0; i < 3; ++i) { yd[i] = dn_const<3>(y0[i]); yd[i].d[i] = 1.0; }
  ionization<DN<3>>(p, c, t0, yd, fd);
  double J[3][3], b[3], k1[3], k2[3], k3[3], k4[3], u[3], f[3], Jk1[3], Jk2[3], Jk3[3];
  LU3 L;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { J[i][j] = fd[i].d[j] * dt; L.m[i][j] = (i == j ? 1.0 : 0.0) - gamma * J[i][j]; }
  lu3_factor(L);
  for (int i = 0; i < 3; ++i) b[i] = fd[i].v * dt;
  lu3_solve(L, b, k1);
  for (int i = 0; i < 3; ++i) u[i] = y0[i] + a21 * k1[i];
  ionization<double>(p, c, t0, u, f);
  for (int i = 0; i < 3; ++i) Jk1[i] = J[i][0] * k1[0] + J[i][1] * k1[1] + J[i][2] * k1[2];
  for (int i = 0; i < 3; ++i) b[i] = f[i] * dt + g21 * Jk1[i];
  lu3_solve(L, b, k2);
  for (int i = 0; i < 3; ++i) u[i] = y0[i] + a31 * k1[i] + a32 * k2[i];
  ionization<double>(p, c, t0, u, f);
  for (int i = 0; i < 3; ++i) Jk2[i] = J[i][0] * k2[0] + J[i][1] * k2[1] + J[i][2] * k2[2];
  for (int i = 0; i < 3; ++i) b[i] = f[i] * dt + g31 * Jk1[i] + g32 * Jk2[i];
  lu3_solve(L, b, k3);
  for (int i = 0; i < 3; ++i) Jk3[i] = J[i][0] * k3[0] + J[i][1] * k3[1] + J[i][2] * k3[2];
  for (int i = 0; i < 3; ++i) b[i] = f[i] * dt + g41 * Jk1[i] + g42 * Jk2[i] + g43 * Jk3[i];
  lu3_solve(L, b, k4);
  for (int i = 0; i < 3; ++i) {
    y1[i] = y0[i] + c1 * k1[i] + c2 * k2[i] + c3 * k3[i] + c4 * k4[i];
    err[i] = (c1 - ch1) * k1[i] + (c2 - ch2) * k2[i] + (c3 - ch3) * k3[i] + c4 * k4[i];
  }
}

// solve_ionization (thermodynamics_recfast.py:283-299): diffrax loop with PIDController(rtol, atol) defaults
// (I-controller: factor = clip(0.9 E^(-1/3), keep ? 1 : 0.2, 10), error order 3 = GRKT4.error_order), dt0 = 1e-3 a,
// at most max_steps attempted steps, throw=False (the state reached is returned); out = (y, dy/da) at the end
BG_DEV void solve_ionization(const BgP& p, const Consts& c, double a0, double a1, const double* ystart, double rtol, double atol,
                             int max_steps, double* out6) {
  double y[3] = {ystart[0], ystart[1], ystart[2]};
  double tprev = a0, tnext = a0 + fabs(a0 * 1e-3);
  if (tnext > a1) tnext = a1;
  int ns = 0;
  while (tprev < a1 && ns < max_steps) {
    double y1[3], err[3];
    grkt4_step(p, c, tprev, tnext - tprev, y, y1, err);
    bool nan1 = false;
    for (int i = 0; i < 3; ++i) { if (err[i] != err[i]) err[i] = INFINITY; nan1 = nan1 || (y1[i] != y1[i]); }
    double s2 = 0.0;
    for (int i = 0; i < 3; ++i) {
      const double yc = nan1 ? y[i] : y1[i];
      const double sc = err[i] / (atol + fmax(fabs(y[i]), fabs(yc)) * rtol);
      s2 += sc * sc;
    }
    const double E = sqrt(s2 / 3.0);
    const bool keep = E < 1.0;
    const double inv = 1.0 / E;
    double fac = 0.9 * pow(inv, 1.0 / 3.0);
    const double fmin_ = keep ? 1.0 : 0.2;
    fac = fac < fmin_ ? fmin_ : (fac > 10.0 ? 10.0 : fac);
    if (!(fac == fac)) fac = NAN;
    const double dt = (tnext - tprev) * fac;
    ++ns;
    if (keep) { y[0] = y1[0]; y[1] = y1[1]; y[2] = y1[2]; tprev = tnext < a1 ? tnext : a1; }
    double tn = tprev + dt;
    if (tn > a1 - 1e-10) tn = keep ? a1 : tprev + 0.5 * (a1 - tprev);
    tnext = tn;
    if (!(tnext == tnext) || isinf(tnext)) break;
  }
  double f[3];
  ionization<double>(p, c, a1, y, f);
  for (int i = 0; i < 3; ++i) { out6[i] = y[i]; out6[3 + i] = f[i]; }
}

// HeII Saha equilibrium (thermodynamics_recfast.py:302-321); T = double or DN<1> (jax.grad at :471)
template <class T>
BG_DEV T saha_HeII(const BgP& p, T a) {
  const double fHe = p.fHe;
  const double Hfac = 1.0 / (1.0e+06 * 3.0856775807e+13);
  const double rho_c = 3.0 * (p.H0 * Hfac) * (p.H0 * Hfac) / (8.0 * BGC_PI * BGC_G);
  const T nH = rho_c * p.Omegab / (BGC_MH * (1.0 / (1.0 - p.YHe))) / (a * a * a);
  const T Tt = p.Tcmb / a;
  const T betaE = BGC_EHE2 / Tt;
  const double A = 1.0 + fHe, B = 1.0 + 2.0 * fHe;
  const T R = xpow(2.0 * BGC_PI * BGC_ME * BGC_KB / (BGC_HP * BGC_HP) * Tt, 1.5) / nH * xexp(-betaE);
  if (val(R) > 1e5) return fHe * (1.0 - B / R + (1.0 + 5.0 * fHe + 6.0 * fHe * fHe) / (R * R));
  const T rma = R - A;
  return -rma / 2.0 + xsqrt(rma * rma / 4.0 + R * B) - A;
}

// the closed forms of the Saha intervals (thermodynamics_recfast.py:400-434): values and d/dz as the reference writes them
BG_DEV void saha_HeI_interval(const BgP& p, const Consts& c, double Nnow, double z1, double* xHe_state, double* dxdz) {
  const double Tc = p.Tcmb, fHe = p.fHe;
  const double rhs = exp(1.5 * log(c.CR * Tc / z1) - c.CB1_He1 / (Tc * z1)) / Nnow * 4.0;
  const double x0 = 0.5 * (sqrt((rhs - 1.0) * (rhs - 1.0) + 4.0 * (1.0 + fHe) * rhs) - (rhs - 1.0));
  const double D = -c.CB1_He1 + c.CR * Tc * Tc;                 // (C_R T^2 - C_B1)
  const double w = pow(D / (Tc * z1), 1.5);                      // (-((CB1 - CR T^2)/(T (1+z))))^1.5
  const double root = sqrt(Nnow * Nnow + 16.0 * (D * D * D) / (Tc * Tc * Tc * z1 * z1 * z1) + 8.0 * (1.0 + 2.0 * fHe) * Nnow * w);
  const double num = -3.0 * pow(-(c.CB1_He1 / Tc) + c.CR * Tc, 1.5) * (Nnow + 2.0 * fHe * Nnow + 4.0 * w - root);
  const double den = Nnow * pow(z1, 2.5) * root;
  *xHe_state = (x0 - 1.0) / fHe;
  *dxdz = num / den;
}
BG_DEV void saha_H_interval(const BgP& p, const Consts& c, double Nnow, double z1, double* xH, double* dxdz) {
  const double Tc = p.Tcmb;
  const double rhs = exp(1.5 * log(c.CR * Tc / z1) - c.CB1 / (Tc * z1)) / Nnow;
  *xH = 0.5 * (sqrt(rhs * rhs + 4.0 * rhs) - rhs);
  const double E = c.CB1 - c.CR * Tc * Tc;                        // (CB1 - CR T^2) < 0
  const double s = sqrt(-(E / (Tc * z1)));
  const double w = pow(-(E / (Tc * z1)), 1.5);
  const double poly = c.CB1 * c.CB1 - 2.0 * c.CB1 * c.CR * Tc * Tc + Tc * Tc * (c.CR * c.CR * Tc * Tc + Nnow * z1 * z1 * s);
  const double t1 = (2.0 * w) / z1;
  const double t2 = (E * (2.0 * c.CB1 * c.CB1 - 4.0 * c.CB1 * c.CR * Tc * Tc + Tc * Tc * (2.0 * c.CR * c.CR * Tc * Tc + Nnow * z1 * z1 * s)))
                  / (pow(Tc, 1.5) * pow(z1, 2.5) * sqrt((-E) * poly));
  *dxdz = (3.0 * (t1 + t2)) / (2.0 * Nnow);
}

// compute_thermal_history (thermodynamics_recfast.py:362-452): the sequential chain over the N intervals of the adaptive
// grid; y6[i] = (x_H, x_He, T_m, d/da of the three) at a_grid[i+1].  One thread.
BG_DEV void thermal_history(const BgP& p, const Consts& c, int N, const double* agrid /*[N+1]*/, double* y6 /*[N][6]*/) {
  const double hh = p.H0 / 100.0, HO = hh * BGC_BIGH;
  const double Nnow = 3.0 * HO * HO * p.Omegab / (8.0 * BGC_PI * BGC_G * (1.0 / (1.0 - p.YHe)) * BGC_MH);
  const double Tc = p.Tcmb;
  double prev[6] = {1.0, 1.0, 0.0, 0.0, 0.0, 0.0};
  for (int i = 0; i < N; ++i) {
    const double a0 = agrid[i], a1 = agrid[i + 1];
    const double z1s = 1.0 / a0, z1e = 1.0 / a1, dzda = -1.0 / (a1 * a1);
    if (i == 0) { prev[0] = 1.0; prev[1] = 1.0; prev[2] = Tc * z1s; prev[3] = 0.0; prev[4] = 0.0; prev[5] = -Tc * z1s; }
    double* o = y6 + (size_t)i * 6;
    if (z1e - 1.0 > 3500.0) {
      o[0] = 1.0; o[1] = 1.0; o[2] = Tc * z1e; o[3] = 0.0; o[4] = 0.0; o[5] = -Tc * z1e;
    } else if (i > 0 && prev[1] > 0.99) {
      double xs, dx;
      saha_HeI_interval(p, c, Nnow, z1e, &xs, &dx);
      o[0] = 1.0; o[1] = xs; o[2] = Tc * z1e; o[3] = 0.0; o[4] = dx * dzda; o[5] = -Tc * z1e;
    } else if (i > 0 && prev[0] > 0.99) {
      double xh, dx;
      saha_H_interval(p, c, Nnow, z1e, &xh, &dx);
      solve_ionization(p, c, a0, a1, prev, 1e-3, 1e-6, 128, o);
      o[0] = xh; o[3] = dx * dzda;
    } else {
      solve_ionization(p, c, a0, a1, prev, 1e-3, 1e-6, 128, o);
    }
    for (int q = 0; q < 6; ++q) prev[q] = o[q];
  }
}

// fill the per-cosmology scalars that do not need the neutrino tables (background.py:148-156)
BG_DEV void base_scalars(const double* in, BgP& p) {
  p.Omegam = in[BI_OMEGAM]; p.Omegab = in[BI_OMEGAB]; p.Omegak = in[BI_OMEGAK]; p.w0 = in[BI_W0]; p.wa = in[BI_WA]; p.H0 = in[BI_H0];
  p.Tcmb = in[BI_TCMB]; p.YHe = in[BI_YHE]; p.Neff = in[BI_NEFF]; p.Nmnu = in[BI_NMNU]; p.mnu = in[BI_MNU];
  const double T2 = p.Tcmb * p.Tcmb, T4 = T2 * T2;
  p.grhom = 3.33795017e-11 * p.H0 * p.H0;
  p.grhog = 1.49594245e-13 * T4;
  p.grhor = 3.39739477e-14 * T4;
  p.adotrad = sqrt((p.grhog + p.grhor * (p.Neff + p.Nmnu)) / 3.0);
  p.amnu = p.mnu * BGC_C2OK / p.Tcmb;
  p.taumin = AMIN / p.adotrad;
  p.fHe = p.YHe / (BGC_MHE * (1.0 - p.YHe));
}

// nodes and weights of nu_background's momentum integral: n-point generalised Gauss-Laguerre rule for x^alpha e^-x
// (util.py:82-123: eigen-decomposition of the Jacobi matrix) re-weighted for the Fermi-Dirac kernel (background.py:40-44)
inline void nu_quadrature(double* q, double* w) {
  const int n = NNUQ;
  const double alpha = 1.0;
  double A[NNUQ][NNUQ] = {{0}}, V[NNUQ][NNUQ] = {{0}};
  for (int i = 0; i < n; ++i) {
    A[i][i] = 2.0 * (i + 1) - 1.0 + alpha; V[i][i] = 1.0;
    if (i + 1 < n) { const double b = sqrt((double)(i + 1) * ((i + 1) + alpha)); A[i][i + 1] = b; A[i + 1][i] = b; }
  }
  for (int sweep = 0; sweep < 100; ++sweep) {            // cyclic Jacobi rotations
    double off = 0.0;
    for (int i = 0; i < n; ++i) for (int j = i + 1; j < n; ++j) off += A[i][j] * A[i][j];
    if (off < 1e-300) break;
    for (int p = 0; p < n; ++p) for (int r = p + 1; r < n; ++r) {
      if (fabs(A[p][r]) < 1e-300) continue;
      const double theta = (A[r][r] - A[p][p]) / (2.0 * A[p][r]);
      const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
      for (int k = 0; k < n; ++k) { const double akp = A[k][p], akr = A[k][r]; A[k][p] = cs * akp - sn * akr; A[k][r] = sn * akp + cs * akr; }
      for (int k = 0; k < n; ++k) { const double apk = A[p][k], ark = A[r][k]; A[p][k] = cs * apk - sn * ark; A[r][k] = sn * apk + cs * ark; }
      for (int k = 0; k < n; ++k) { const double vkp = V[k][p], vkr = V[k][r]; V[k][p] = cs * vkp - sn * vkr; V[k][r] = sn * vkp + cs * vkr; }
    }
  }
  int idx[NNUQ];
  for (int i = 0; i < n; ++i) idx[i] = i;
  for (int i = 0; i < n; ++i) for (int j = i + 1; j < n; ++j) if (A[idx[j]][idx[j]] < A[idx[i]][idx[i]]) { const int t = idx[i]; idx[i] = idx[j]; idx[j] = t; }
  for (int i = 0; i < n; ++i) {
    const double node = A[idx[i]][idx[i]], v0 = V[0][idx[i]];
    double wt = v0 * v0 * 1.0;                              // Gamma(alpha + 1) = 1! = 1
    wt *= node * node * node / (1.0 + exp(-node)) * pow(node, -alpha);
    q[i] = node; w[i] = wt / 5.682196976983475;
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// one cosmology, start to finish.  Device: executed by all threads of one CTA (tid, nthr), phases separated by
// barriers; host build: tid = 0, nthr = 1.  `ws` = BgWork in shared memory (device) or on the heap (host).
// ---------------------------------------------------------------------------------------------------------------------
#ifdef DEB_CPU_EMU
#define BG_SYNC()
#else
#define BG_SYNC() __syncthreads()
#endif
constexpr int NTH_MAX = 1024, NNU = 512;
struct BgWork {
  double nx[NNU], ry[NNU], py[NNU], rS[NNU], pS[NNU];      // neutrino tables: log a, log rho, log p and their second derivatives
  double cp[5][NTH_MAX > NNU ? NTH_MAX : NNU], dp[5][NTH_MAX > NNU ? NTH_MAX : NNU];      // Thomas scratch, 5 splines at a time
  double ag[NTH_MAX + 1], tau[NTH_MAX], dtau[NTH_MAX], y6[NTH_MAX * 6];
  double la[NTH_MAX], xe[NTH_MAX], cs2a[NTH_MAX];
  double ppy[NNU];                                              // log pseudo-pressure of the neutrino table (extras only)
  double tba[NTH_MAX], opS[NTH_MAX], icum[NTH_MAX];             // a T_m, opacity spline, its cumulative integral (extras only)
  BgP p;
};

// Everything else the reference's evolve_background leaves in `param` (background.py:238-246, 250-251, 167, 306-342;
// thermodynamics_recfast.py:460-478), per cosmology, in this order; rows of nth doubles unless noted:
enum BgExtra { BX_XEHI, BX_XEHEI, BX_XEHEII, BX_CS2, BX_TM, BX_XEPRIME_RECF, BX_XEPRIME, BX_OPAC, BX_OPTICAL_DEPTH, BX_GVIS, BX_GVISPRIME,
               BX_GVISPPRIME, BX_CS2A_OF_TAU_S, BX_TEMPBA_OF_TAU_S, BX_NROWS };       // then log ppseudo_nu [NNU] and its S [NNU]
BG_HD size_t extras_len(int nth) { return (size_t)BX_NROWS * nth + 2 * NNU; }

// tables of one cosmology in the layout of deb_spline (include/discoeb_b200.h): 7 splines x (x, y, S)
BG_DEV void background_one(const double* in, const double* q8, const double* w8, int nth, double* scal, double* tab, BgWork& W, int tid, int nthr,
                           double* ext = nullptr) {
  const Consts c = make_consts();
  BgP& p = W.p;
  double* t_cs2a = tab;                     // x, y, S each nth
  double* t_xe = t_cs2a + 3 * nth;
  double* t_lrn = t_xe + 3 * nth;           // each NNU
  double* t_lpn = t_lrn + 3 * NNU;
  double* t_aot = t_lpn + 3 * NNU;
  double* t_xot = t_aot + 3 * nth;
  double* t_toa = t_xot + 3 * nth;
  if (tid == 0) { base_scalars(in, p); p.lx = W.nx; p.ly = W.ry; p.lS = W.rS; p.nnu = NNU; }
  BG_SYNC();
  // ---- neutrino density / pressure tables on 512 log-spaced knots (background.py:157-167) ----
  for (int i = tid; i < NNU; i += nthr) {
    const double a = geom_point(AMIN * 0.9, AMAX * 1.1, i, NNU, true);
    double r, pr, pp;
    nu_background(a, p.amnu, q8, w8, &r, &pr, &pp);
    W.nx[i] = log(a); W.ry[i] = log(r); W.py[i] = log(pr);
    if (ext) W.ppy[i] = log(pp);
  }
  BG_SYNC();
  if (tid == 0) spline_build(W.nx, W.ry, W.rS, NNU, W.cp[0], W.dp[0]);
  if (tid == (nthr > 1 ? 1 : 0)) spline_build(W.nx, W.py, W.pS, NNU, W.cp[1], W.dp[1]);
  BG_SYNC();
  if (tid == 0) {
    // closure: densities today (background.py:169-178)
    const double rhonu0 = exp(spline_eval(W.nx, W.ry, W.rS, NNU, 0.0));
    const double Omegar = (p.Neff + p.Nmnu * rhonu0) * p.grhor / p.grhom, Omegag = p.grhog / p.grhom;
    p.OmegaDE = 1.0 - p.Omegak - Omegar - Omegag - p.Omegam;
    scal[20] = p.grhor * rhonu0 / p.grhom;      // Omegamnu
  }
  for (int i = tid; i <= nth; i += nthr) W.ag[i] = adaptive_a(AMIN, AMAX, nth + 1, i);
  BG_SYNC();
  // ---- conformal time: Romberg sum of d tau/d a over every interval of the thermo grid (thermodynamics_recfast.py:480-497);
  //      the knots are ag[1..nth], tau of the first knot is taumin (the reference's convention) ----
  for (int i = tid; i < nth - 1; i += nthr) W.dtau[i] = romb_dtauda(p, W.ag[i + 1], W.ag[i + 2]);
  if (tid == (nthr > 2 ? 2 : 0)) scal[19] = p.taumin + romb_dtauda(p, AMIN, AMAX);      // taumax (background.py:181-186)
  // ---- thermal history: sequential over the intervals, one thread (the other threads run the Romberg sums above) ----
  if (tid == (nthr > 32 ? 32 : 0)) thermal_history(p, c, nth, W.ag, W.y6);
  BG_SYNC();
  if (tid == 0) {
    double t = p.taumin;
    W.tau[0] = t;
    for (int i = 0; i < nth - 1; ++i) { t += W.dtau[i]; W.tau[i + 1] = t; }
  }
  // ---- x_e, c_s^2 on the knots (thermodynamics_recfast.py:460-478) ----
  for (int i = tid; i < nth; i += nthr) {
    const double a = W.ag[i + 1];
    const double* y = W.y6 + (size_t)i * 6;
    DN<1> ad; ad.v = a; ad.d[0] = 1.0;
    const DN<1> he2 = saha_HeII<DN<1>>(p, ad);
    const double xe = y[0] + p.fHe * y[1] + he2.v;
    const double mu = 1.0 / (1.0 + (1.0 / BGC_MHE - 1.0) * p.YHe + (1.0 - p.YHe) * xe);
    const double Tm = y[2], daTmda = Tm + a * y[5];
    const double cs2 = BGC_KB / BGC_MH / (BGC_C * BGC_C) / mu * Tm * (4.0 - daTmda / Tm) / 3.0;
    W.la[i] = log(a); W.xe[i] = xe; W.cs2a[i] = a * cs2;
    if (ext) {
      ext[BX_XEHI * nth + i] = y[0]; ext[BX_XEHEI * nth + i] = y[1]; ext[BX_XEHEII * nth + i] = he2.v;
      ext[BX_CS2 * nth + i] = cs2; ext[BX_TM * nth + i] = Tm;
      // d x_e / d tau as RECFAST knows it (thermodynamics_recfast.py:477): d/da of the three species times a' (background.py:83-92)
      ext[BX_XEPRIME_RECF * nth + i] = (y[3] + p.fHe * y[4] + he2.d[0]) / dtauda(p, a);
      W.tba[i] = a * Tm;
    }
  }
  BG_SYNC();
  // ---- the five thermo splines (background.py:240-255), one thread each ----
  {
    const int who = nthr >= 5 ? tid : -1;
    for (int sp = 0; sp < 5; ++sp) {
      if (who >= 0 ? who != sp : tid != 0) continue;
      const double* x = sp == 0 || sp == 1 ? W.la : (sp == 4 ? W.ag + 1 : W.tau);
      const double* y = sp == 0 ? W.cs2a : (sp == 1 || sp == 3 ? W.xe : (sp == 2 ? W.ag + 1 : W.tau));
      double* dst = sp == 0 ? t_cs2a : (sp == 1 ? t_xe : (sp == 2 ? t_aot : (sp == 3 ? t_xot : t_toa)));
      spline_build(x, y, dst + 2 * nth, nth, W.cp[sp], W.dp[sp]);
    }
  }
  for (int i = tid; i < nth; i += nthr) {
    t_cs2a[i] = W.la[i]; t_cs2a[nth + i] = W.cs2a[i];
    t_xe[i] = W.la[i]; t_xe[nth + i] = W.xe[i];
    t_aot[i] = W.tau[i]; t_aot[nth + i] = W.ag[i + 1];
    t_xot[i] = W.tau[i]; t_xot[nth + i] = W.xe[i];
    t_toa[i] = W.ag[i + 1]; t_toa[nth + i] = W.tau[i];
  }
  for (int i = tid; i < NNU; i += nthr) {
    t_lrn[i] = W.nx[i]; t_lrn[NNU + i] = W.ry[i]; t_lrn[2 * NNU + i] = W.rS[i];
    t_lpn[i] = W.nx[i]; t_lpn[NNU + i] = W.py[i]; t_lpn[2 * NNU + i] = W.pS[i];
  }
  if (ext) {
    // ---- the rest of evolve_background's outputs (background.py:250-251, 167, 306-342) ----
    BG_SYNC();
    double* logpp = ext + (size_t)BX_NROWS * nth;
    if (tid == 0) spline_build(W.tau, W.cs2a, ext + BX_CS2A_OF_TAU_S * nth, nth, W.cp[0], W.dp[0]);
    if (tid == (nthr > 1 ? 1 : 0)) spline_build(W.tau, W.tba, ext + BX_TEMPBA_OF_TAU_S * nth, nth, W.cp[1], W.dp[1]);
    if (tid == (nthr > 2 ? 2 : 0)) spline_build(W.nx, W.ppy, logpp + NNU, NNU, W.cp[2], W.dp[2]);
    for (int i = tid; i < NNU; i += nthr) logpp[i] = W.ppy[i];
    // optical depth and visibility on the knots: opacity x_e sigma_T n_e / a^2 with x_e frozen at full ionisation before
    // a = 1e-4, its natural spline over tau, the spline's derivatives and its integral from tau to the last knot, shifted to
    // vanish today
    const double akthom = 2.3038921003709498e-9 * (1.0 - p.YHe) * p.Omegab * p.H0 * p.H0;
    const double tau_pre = spline_eval(t_toa, t_toa + nth, t_toa + 2 * nth, nth, 1e-4);
    const double xe_full = 1.0 + p.YHe / (1.0 - p.YHe);
    double* opac = ext + BX_OPAC * nth;
    for (int i = tid; i < nth; i += nthr) {
      const bool pre = W.tau[i] <= tau_pre;
      double d1, d2;
      spline_knot_derivs(W.tau, W.xe, t_xot + 2 * nth, nth, i, &d1, &d2);
      ext[BX_XEPRIME * nth + i] = pre ? 0.0 : d1;
      const double a = W.ag[i + 1];
      opac[i] = (pre ? xe_full : W.xe[i]) * akthom / (a * a);
    }
    BG_SYNC();
    if (tid == 0) {
      spline_build(W.tau, opac, W.opS, nth, W.cp[3], W.dp[3]);
      double acc = 0.0;
      W.icum[0] = 0.0;
      for (int k = 0; k < nth - 1; ++k) { acc += spline_interval_integral(W.tau, opac, W.opS, k, W.tau[k + 1] - W.tau[k]); W.icum[k + 1] = acc; }
    }
    BG_SYNC();
    {
      const double itot = W.icum[nth - 1];
      const double tau0 = spline_eval(t_toa, t_toa + nth, t_toa + 2 * nth, nth, 1.0);
      const double today = itot - spline_integral_from_start(W.tau, opac, W.opS, W.icum, nth, tau0);
      for (int i = tid; i < nth; i += nthr) {
        double o1, o2;
        spline_knot_derivs(W.tau, opac, W.opS, nth, i, &o1, &o2);
        // (the reference evaluates the integral at the knot through the interval to its left, like the derivatives)
        int k = i - 1;
        if (k < 0) k = 0;
        const double od = (itot - (W.icum[k] + spline_interval_integral(W.tau, opac, W.opS, k, W.tau[i] - W.tau[k]))) - today;
        const double em = exp(-od), op = opac[i];
        ext[BX_OPTICAL_DEPTH * nth + i] = od;
        ext[BX_GVIS * nth + i] = op * em;
        ext[BX_GVISPRIME * nth + i] = (o1 + op * op) * em;
        ext[BX_GVISPPRIME * nth + i] = (o2 + 3.0 * op * o1 + op * op * op) * em;
      }
    }
  }
  if (tid == 0) {
    // scalars in the order of deb_scalar (include/discoeb_b200.h)
    scal[0] = p.Omegam; scal[1] = p.Omegab; scal[2] = p.OmegaDE; scal[3] = p.Omegak; scal[4] = p.grhom; scal[5] = p.grhog; scal[6] = p.grhor;
    scal[7] = p.Neff; scal[8] = p.Nmnu; scal[9] = p.amnu; scal[10] = p.w0; scal[11] = p.wa; scal[12] = in[BI_CS2DE]; scal[13] = p.YHe;
    scal[14] = p.H0; scal[15] = p.taumin; scal[16] = in[BI_AS]; scal[17] = in[BI_NS]; scal[18] = in[BI_KP];
    scal[21] = p.Tcmb; scal[22] = p.mnu; scal[23] = 0.0;
  }
  BG_SYNC();
}

}  // namespace bg
}  // namespace deb
