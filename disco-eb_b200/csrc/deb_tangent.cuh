// deb_tangent.cuh -- forward-mode tangents of the per-mode integrator.
//
// What jax.jvp / jax.jacfwd of the reference's evolve_perturbations computes is the exact derivative of the
// DISCRETE solve: tangents flow through the start time, the initial conditions, every Rosenbrock stage
// including W = I/(gamma dt) - J (ode_integrators_stiff.py:772-779 carry no stop_gradient), the step end
// points, the SaveAt interpolation weights and the output conversion, while accept/reject, the PID factor,
// spline intervals and bisection branches are decided on primal values.  Differentiating stage i,
//     W k_i = r_i ,   r_i = f(t_i, u_i) + dt d_i dT + sum_j C_ij/dt k_j ,
// along one direction (dot = d/d eps) gives a linear system with the SAME matrix:
//     W kdot_i = rdot_i + ddt/(gamma dt^2) k_i + Jdot k_i .
// The right-hand side f is linear in every state variable except the scale factor a = y[0],
//     f = A(a, t; theta) y~ + b(a; theta) e_0 ,          J k = A k~ + (df/da) k_0 ,
// so   Jdot k_i = Adot_0 k~_i + (d/d eps df/da) k_{i,0}   needs
//   * Adot_0 k~_i : the same row code evaluated on first-order duals with the state held constant (per stage);
//   * d/d eps df/da : the row code evaluated once per step on second-order duals (eps x delta_a).
// Hence a tangent costs 8 extra structured solves with the primal factorisation + 17 dual row evaluations per
// step, and nothing is ever refactored.  The row code below is a plain, generic (templated on the scalar type)
// statement of perturbations.py:84-371; the primal path keeps its own tuned table-driven version.
#pragma once

namespace deb {

// ---- nested forward duals -------------------------------------------------------------------------
template <class S> struct DualOf { S v, d; };
typedef DualOf<Dual> HD;              // v = (value, d/d eps), d = (d/da, d2/(da d eps))

template <class S> DEB_DEV DualOf<S> mkd(S v, S d) { DualOf<S> r; r.v = v; r.d = d; return r; }
template <class S> DEB_DEV DualOf<S> operator+(DualOf<S> a, DualOf<S> b) { return mkd<S>(a.v + b.v, a.d + b.d); }
template <class S> DEB_DEV DualOf<S> operator-(DualOf<S> a, DualOf<S> b) { return mkd<S>(a.v - b.v, a.d - b.d); }
template <class S> DEB_DEV DualOf<S> operator*(DualOf<S> a, DualOf<S> b) { return mkd<S>(a.v * b.v, a.d * b.v + a.v * b.d); }
template <class S> DEB_DEV DualOf<S> operator/(DualOf<S> a, DualOf<S> b) { S q = a.v / b.v; return mkd<S>(q, (a.d - q * b.d) / b.v); }
template <class S> DEB_DEV DualOf<S> operator+(DualOf<S> a, double b) { return mkd<S>(a.v + b, a.d); }
template <class S> DEB_DEV DualOf<S> operator+(double b, DualOf<S> a) { return mkd<S>(a.v + b, a.d); }
template <class S> DEB_DEV DualOf<S> operator-(DualOf<S> a, double b) { return mkd<S>(a.v - b, a.d); }
template <class S> DEB_DEV DualOf<S> operator-(double b, DualOf<S> a) { return mkd<S>(b - a.v, -a.d); }
template <class S> DEB_DEV DualOf<S> operator-(DualOf<S> a) { return mkd<S>(-a.v, -a.d); }
template <class S> DEB_DEV DualOf<S> operator*(DualOf<S> a, double b) { return mkd<S>(a.v * b, a.d * b); }
template <class S> DEB_DEV DualOf<S> operator*(double b, DualOf<S> a) { return mkd<S>(a.v * b, a.d * b); }
template <class S> DEB_DEV DualOf<S> operator/(DualOf<S> a, double b) { return mkd<S>(a.v / b, a.d / b); }
template <class S> DEB_DEV DualOf<S> operator/(double b, DualOf<S> a) { S q = b / a.v; return mkd<S>(q, -(q * a.d) / a.v); }
template <class S> DEB_DEV DualOf<S> dsqrt(DualOf<S> a) { S s = dsqrt(a.v); return mkd<S>(s, (0.5 * a.d) / s); }
template <class S> DEB_DEV DualOf<S> dexp(DualOf<S> a) { S e = dexp(a.v); return mkd<S>(e, e * a.d); }
template <class S> DEB_DEV DualOf<S> dlog(DualOf<S> a) { return mkd<S>(dlog(a.v), a.d / a.v); }
template <class S> DEB_DEV DualOf<S> drsqrt(DualOf<S> a) { S r = drsqrt(a.v); return mkd<S>(r, (-0.5 * r) * a.d / a.v); }
template <class S> DEB_DEV double val(DualOf<S> a) { return val(a.v); }

// (value, eps-part) -> T
template <class T> DEB_DEV T lift2(double v, double d);
template <> DEB_DEV Dual lift2<Dual>(double v, double d) { return mk(v, d); }
template <> DEB_DEV HD lift2<HD>(double v, double d) { return mkd<Dual>(mk(v, d), mk(0.0, 0.0)); }
template <class T> DEB_DEV T liftD(Dual x) { return lift2<T>(x.v, x.d); }
DEB_DEV double eps_part(Dual x) { return x.d; }
DEB_DEV double eps_part(HD x) { return x.v.d; }

// ---- cosmology scalars and tables with their tangents ------------------------------------------------
struct SplT { Spl p, t; };            // primal arrays and tangent arrays (same layout)
struct CosmoD {                       // one direction: every scalar as (value, d/d eps)
  Dual Omegam, Omegab, OmegaDE, Omegak, grhom, grhog, grhor, Neff, Nmnu, amnu, w0, wa, cs2de, YHe, H0, taumin, As, ns, kp;
  SplT cs2a, xe, lrn, lpn, a_of_tau, xe_of_tau, tau_of_a;
};
DEB_DEV Spl get_spline_from(const double* tables, const Problem& P, int slot, int which) {
  size_t tl = 3 * (size_t)(5 * P.nth + 2 * P.nnu);
  const double* base = tables + (size_t)slot * tl;
  size_t off = 0;
  for (int s = 0; s < which; ++s) off += 3 * (size_t)((s == T_LRHONU || s == T_LPNU) ? P.nnu : P.nth);
  int n = (which == T_LRHONU || which == T_LPNU) ? P.nnu : P.nth;
  Spl r; r.x = base + off; r.y = r.x + n; r.S = r.y + n; r.n = n;
  return r;
}
DEB_DEV void load_cosmo_d(const Problem& P, int c, int tan, CosmoD* dst) {      // dst: shared memory
  const double* s = P.scalars + (size_t)c * NSCAL;
  const size_t slot = (size_t)tan * P.ncosmo + c;
  const double* d = P.d_scalars + slot * NSCAL;
#define DEB_LD(field, idx) dst->field = mk(DEB_LDG(s + idx), DEB_LDG(d + idx))
  DEB_LD(Omegam, S_OMEGAM); DEB_LD(Omegab, S_OMEGAB); DEB_LD(OmegaDE, S_OMEGADE); DEB_LD(Omegak, S_OMEGAK);
  DEB_LD(grhom, S_GRHOM); DEB_LD(grhog, S_GRHOG); DEB_LD(grhor, S_GRHOR); DEB_LD(Neff, S_NEFF); DEB_LD(Nmnu, S_NMNU);
  DEB_LD(amnu, S_AMNU); DEB_LD(w0, S_W0); DEB_LD(wa, S_WA); DEB_LD(cs2de, S_CS2DE); DEB_LD(YHe, S_YHE); DEB_LD(H0, S_H0);
  DEB_LD(taumin, S_TAUMIN); DEB_LD(As, S_AS); DEB_LD(ns, S_NS); DEB_LD(kp, S_KP);
#undef DEB_LD
#define DEB_LS(field, w) dst->field.p = get_spline(P, c, w); dst->field.t = get_spline_from(P.d_tables, P, (int)slot, w)
  DEB_LS(cs2a, T_CS2A); DEB_LS(xe, T_XE); DEB_LS(lrn, T_LRHONU); DEB_LS(lpn, T_LPNU); DEB_LS(a_of_tau, T_A_OF_TAU);
  DEB_LS(xe_of_tau, T_XE_OF_TAU); DEB_LS(tau_of_a, T_TAU_OF_A);
#undef DEB_LS
}

// spline value with tangents in the knots, the values, the second derivatives and the abscissa
// (spline_interpolation.py:130-153; the interval comes from primal values)
template <class T>
DEB_DEV T spl_eval_g(const SplT& s, T xn, int hint = -1) {
  const int i = spl_locate(s.p.x, s.p.n, val(xn), hint);
  const T x0 = lift2<T>(DEB_LDG(s.p.x + i), DEB_LDG(s.t.x + i)), x1 = lift2<T>(DEB_LDG(s.p.x + i + 1), DEB_LDG(s.t.x + i + 1));
  const T y0 = lift2<T>(DEB_LDG(s.p.y + i), DEB_LDG(s.t.y + i)), y1 = lift2<T>(DEB_LDG(s.p.y + i + 1), DEB_LDG(s.t.y + i + 1));
  const T S0 = lift2<T>(DEB_LDG(s.p.S + i), DEB_LDG(s.t.S + i)), S1 = lift2<T>(DEB_LDG(s.p.S + i + 1), DEB_LDG(s.t.S + i + 1));
  const T h = x1 - x0;
  const T t = (xn - x0) / h;
  const T A = 1.0 - t, B = t;
  return A * y0 + B * y1 + ((A * A * A - A) * S0 + (B * B * B - B) * S1) * (h * h) / 6.0;
}

// ---- background coefficients (perturbations.py:176-218, background.py:110-121), generic --------------
// (scalars only, so that the struct lives in registers; the per-bin velocities v_i go through shared memory)
template <class T> struct BgG {
  T a, H, opac, k2cs2, pbo, wq1, wq, ca2, gc, gb, gg, gr, gnu, gq;
};
constexpr int BG_NS = 14;             // scalars in BgG
// v_i = 1/sqrt(1 + (a amnu/q_i)^2) as T -> 4 doubles per bin (value, eps, d/da, d2); Dual uses the first two
template <class T> DEB_DEV void vv_store(double* dst, T v);
template <> DEB_DEV void vv_store<Dual>(double* dst, Dual v) { dst[0] = v.v; dst[1] = v.d; }
template <> DEB_DEV void vv_store<HD>(double* dst, HD v) { dst[0] = v.v.v; dst[1] = v.v.d; dst[2] = v.d.v; dst[3] = v.d.d; }
template <class T> DEB_DEV T vv_load(const double* src);
template <> DEB_DEV Dual vv_load<Dual>(const double* src) { return mk(src[0], src[1]); }
template <> DEB_DEV HD vv_load<HD>(const double* src) { return mkd<Dual>(mk(src[0], src[1]), mk(src[2], src[3])); }
template <class T>
DEB_DEV void nu_velocity_g(const CosmoD& c, const NuBins& nb, T a, int i, double* dst) {
  const T aq = a * (liftD<T>(c.amnu) / nb.q[i]);
  vv_store<T>(dst, drsqrt(1.0 + aq * aq));
}
template <class T>
DEB_DEV void compute_bg_g(const CosmoD& c, T a, Dual kD, int hint_th, int hint_nu, BgG<T>& b) {
  const T loga = dlog(a);
  const T inva = 1.0 / a, inva2 = inva * inva;
  const T grhom = liftD<T>(c.grhom), grhog = liftD<T>(c.grhog), grhor = liftD<T>(c.grhor);
  const T Omegab = liftD<T>(c.Omegab), Omegam = liftD<T>(c.Omegam), wa = liftD<T>(c.wa), w0 = liftD<T>(c.w0);
  const T Neff = liftD<T>(c.Neff), Nmnu = liftD<T>(c.Nmnu);
  b.a = a;
  const T cs2 = spl_eval_g<T>(c.cs2a, loga, hint_th) * inva;
  const T kT = liftD<T>(kD);
  b.k2cs2 = (kT * kT) * cs2;
  const T xe = spl_eval_g<T>(c.xe, loga, hint_th);
  const T rhonu = dexp(spl_eval_g<T>(c.lrn, loga, hint_nu));
  const T rhoq = dexp((-3.0 * (1.0 + w0 + wa)) * loga + 3.0 * wa * (a - 1.0));
  b.wq = w0 + wa * (1.0 - a);
  b.wq1 = 1.0 + b.wq;
  b.gc = (grhom * (Omegam - Omegab)) * inva;
  b.gb = (grhom * Omegab) * inva;
  b.gg = grhog * inva2;
  b.gr = (grhor * Neff) * inva2;
  b.gnu = (grhor * Nmnu) * inva2;
  b.gq = (grhom * liftD<T>(c.OmegaDE)) * rhoq * (a * a);
  const T grho = (grhom * Omegam) * inva + (grhog + grhor * (Neff + Nmnu * rhonu)) * inva2 + b.gq + grhom * liftD<T>(c.Omegak);
  b.H = dsqrt(grho * (1.0 / 3.0));
  b.ca2 = b.wq + (wa * a) / (3.0 * (b.wq1 + 1e-6));
  const T H0 = liftD<T>(c.H0);
  const T akthom = (AKTHOM_RHS * (1.0 - liftD<T>(c.YHe))) * Omegab * H0 * H0;
  b.opac = xe * akthom * inva2;
  b.pbo = ((4.0 / 3.0) * grhog / (grhom * Omegab)) * inva * b.opac;
}
// a BgG<Dual> to / from 2 BG_NS doubles of shared memory (the base-point coefficients are per step, not per stage)
#define DEB_BG_FIELDS(X) X(a, 0) X(H, 1) X(opac, 2) X(k2cs2, 3) X(pbo, 4) X(wq1, 5) X(wq, 6) X(ca2, 7) X(gc, 8) X(gb, 9) X(gg, 10) \
  X(gr, 11) X(gnu, 12) X(gq, 13)
DEB_DEV void bg_store(double* dst, const BgG<Dual>& b) {
#define DEB_X(f, i) dst[2 * i] = b.f.v; dst[2 * i + 1] = b.f.d;
  DEB_BG_FIELDS(DEB_X)
#undef DEB_X
}
DEB_DEV void bg_load(const double* src, BgG<Dual>& b) {
#define DEB_X(f, i) b.f = mk(src[2 * i], src[2 * i + 1]);
  DEB_BG_FIELDS(DEB_X)
#undef DEB_X
}

// state accessors: u[e] as T
struct StateVD { const double* v; const double* d; };        // (value, eps-part) from two vectors
struct StateV0 { const double* v; };                          // constant state (eps-part zero)
template <class T> DEB_DEV T sget(const StateVD& s, int e) { return lift2<T>(s.v[e], s.d[e]); }
template <class T> DEB_DEV T sget(const StateV0& s, int e) { return lift2<T>(s.v[e], 0.0); }

// metric sources (perturbations.py:229-261), generic
template <class T> struct MetricG { T hp, ep, al, f1; };
template <class T, class St>
DEB_DEV void compute_metric_g(const Problem& P, const CosmoD& c, const NuBins& nb, const BgG<T>& b, const double* vvs, const St& u,
                              Dual kD, MetricG<T>& mt) {
  const int nq = P.nq, iq0 = P.iq0, n = P.n;
  const T eta = sget<T>(u, 2), dc = sget<T>(u, 3), tc = sget<T>(u, 4), db = sget<T>(u, 5), tb = sget<T>(u, 6);
  const T dg = sget<T>(u, 7), tg = sget<T>(u, 8), dr = sget<T>(u, P.ir), tr = sget<T>(u, P.ir + 1);
  const T dq = sget<T>(u, n - 2), tq = sget<T>(u, n - 1);
  T drhonu = 0.0 * b.a, dpnu3 = 0.0 * b.a, fnu = 0.0 * b.a;
  for (int i = 0; i < nq; ++i) {
    const T p0 = nb.w[i] * sget<T>(u, iq0 + i);
    const T v = vv_load<T>(vvs + 4 * i);
    drhonu = drhonu + p0 / v;
    dpnu3 = dpnu3 + p0 * v;
    fnu = fnu + nb.w[i] * sget<T>(u, iq0 + nq + i);
  }
  const T k = liftD<T>(kD), k2 = k * k, ik2 = 1.0 / k2;
  const T cs2de = liftD<T>(c.cs2de);
  const T rpt = b.wq1 * b.gq * tq;
  const T dgrho = b.gc * dc + b.gb * db + b.gg * dg + b.gr * dr + b.gnu * drhonu + b.gq * dq;
  const T dgpres3 = (b.gg * dg + b.gr * dr) + b.gnu * dpnu3 + 3.0 * (cs2de * (b.gq * dq)) + (cs2de - b.ca2) * ((9.0 * ik2) * (b.H * rpt));
  const T dgtheta = b.gc * tc + b.gb * tb + (4.0 / 3.0) * (b.gg * tg + b.gr * tr) + b.gnu * (k * fnu) + rpt;
  mt.f1 = -(dgrho + dgpres3) * b.a;
  mt.hp = ((2.0 * k2) * eta + dgrho) / b.H;
  mt.ep = (0.5 * ik2) * dgtheta;
  mt.al = (mt.hp + 6.0 * mt.ep) * (0.5 * ik2);
}

// ---- rows, branch-free -------------------------------------------------------------------------------
// The 17+3nq head rows (eta, fluids, l <= 2 of every hierarchy, dark energy) are evaluated one lane per row from the
// primal path's operator table (CtaConst::hop_*: constant x background slot x state index), with the slots as T;
// the l >= 3 rows of the hierarchies by the generic three-term recurrence over the tail list.  (A first version
// evaluated every element through one switch over the row type: with 32 lanes on ~12 different row types per trip
// the warp executed each case in turn -- 112 k warp-instructions per step.)
// A T lives in shared memory as 4 doubles (value, eps, d/da, d2); Dual uses the first two.
template <class T>
DEB_DEV void fill_slots_g(const CosmoD& c, const BgG<T>& b, Dual kD, double* sl) {      // slots 0 .. SL_KV0-1 (one lane)
  const T k = liftD<T>(kD), k2 = k * k;
  const T cs2de = liftD<T>(c.cs2de);
  const T one = 0.0 * b.a + 1.0;
  vv_store<T>(sl + 4 * SL_ONE, one); vv_store<T>(sl + 4 * SL_H, b.H); vv_store<T>(sl + 4 * SL_OPAC, b.opac);
  vv_store<T>(sl + 4 * SL_PBO, b.pbo); vv_store<T>(sl + 4 * SL_K2CS2, b.k2cs2); vv_store<T>(sl + 4 * SL_WQ1, b.wq1);
  vv_store<T>(sl + 4 * SL_DQD, (cs2de - b.wq) * b.H);
  vv_store<T>(sl + 4 * SL_DQT, b.wq1 * (cs2de - b.ca2) * (b.H * b.H) * (1.0 / k2));
  vv_store<T>(sl + 4 * SL_TQD, (cs2de * k2) / b.wq1);
  vv_store<T>(sl + 4 * SL_TQT, (1.0 - 3.0 * cs2de) * b.H);
  vv_store<T>(sl + 4 * SL_K, k); vv_store<T>(sl + 4 * SL_K2, k2);
}
// per chain: wavenumber k or k v_i, damping opac or 0; lane `ch` < nch.  Also the k v_i slots of the head operator.
template <class T>
DEB_DEV void chain_coeffs_g(const BgG<T>& b, const double* vvs, Dual kD, int ch, double* kcs, double* kaps, double* sl) {
  const T k = liftD<T>(kD);
  T kc = k;
  if (ch >= 3) { kc = k * vv_load<T>(vvs + 4 * (ch - 3)); vv_store<T>(sl + 4 * (SL_KV0 + ch - 3), kc); }
  vv_store<T>(kcs + 4 * ch, kc);
  vv_store<T>(kaps + 4 * ch, ch < 2 ? b.opac : 0.0 * b.a);
}
template <class T, class St>
DEB_DEV T head_row_g(const CtaConst& C, const double* sl, const St& u, int r, const MetricG<T>& mt) {
  T f = (C.hop_chc[r] * vv_load<T>(sl + 4 * C.hop_chs[r])) * mt.hp + C.hop_cec[r] * mt.ep;
#pragma unroll
  for (int t = 0; t < HOP_NT; ++t) {
    const int m = C.hop_meta[t][r];
    f = f + (C.hop_c[t][r] * sget<T>(u, (m >> 8) & 0xfff)) * vv_load<T>(sl + 4 * (m & 0xff));
  }
  return C.htype[r] == R_AHP ? mt.f1 : f;
}
template <class T, class St>
DEB_DEV T tail_row_g(const CtaConst& C, const double* kcs, const double* kaps, const St& u, int info, T invtau, int* e_out,
                     double* trunc_out) {
  const int e = info & 0xfff, chain = (info >> 12) & 0xf, l = info >> 16;
  const int s = C.ch_stride[chain], L = C.ch_lmax[chain];
  const bool isT = (l == L);
  const double clv = isT ? 1.0 : C.cl[l], chv = isT ? 0.0 : C.ch[l];
  const double tr = isT ? (double)(L + 1) : 0.0;
  const T lin = clv * sget<T>(u, e - s) - chv * sget<T>(u, isT ? e : e + s);
  *e_out = e; *trunc_out = tr;
  return vv_load<T>(kcs + 4 * chain) * lin - (vv_load<T>(kaps + 4 * chain) + tr * invtau) * sget<T>(u, e);
}

// ---- prologue with tangents: start time (perturbations.py:630-681) and adiabatic ICs (:526-627) ------------
DEB_DEV Dual aprimeoa_d(const CosmoD& c, Dual a) {
  const Dual loga = dlog(a);
  const Dual rhonu = dexp(spl_eval_g<Dual>(c.lrn, loga));
  const Dual rhoq = dexp((-3.0 * (1.0 + c.w0 + c.wa)) * loga + 3.0 * c.wa * (a - 1.0));
  const Dual grho = c.grhom * c.Omegam / a + (c.grhog + c.grhor * (c.Neff + c.Nmnu * rhonu)) / (a * a)
                  + c.grhom * c.OmegaDE * rhoq * (a * a) + c.grhom * c.Omegak;
  return dsqrt(grho / 3.0);
}
DEB_DEV Dual cond_small_k_d(const CosmoD& c, Dual lt) {
  const Dual tau = dexp(lt);
  const Dual akthom = (AKTHOM_START * (1.0 - c.YHe)) * c.Omegab * c.H0 * c.H0;
  const Dual xe = spl_eval_g<Dual>(c.xe_of_tau, tau);
  const Dual a = spl_eval_g<Dual>(c.a_of_tau, tau);
  const Dual opac = xe * akthom / (a * a);
  const Dual H = aprimeoa_d(c, a);
  return (1.0 / opac) / (1.0 / H) / 0.0004 - 1.0;
}
DEB_DEV Dual cond_large_k_d(const CosmoD& c, Dual lt, Dual k) {
  const Dual a = spl_eval_g<Dual>(c.a_of_tau, dexp(lt));
  return (1.0 / aprimeoa_d(c, a)) / (1.0 / k) / 0.07 - 1.0;
}
// bisection of util.py:365-396: the branches are decided on primal values, the end points carry tangents
DEB_DEV Dual start_time_d(const CosmoD& c, Dual k) {
  Dual res0 = mk(0.0, 0.0), res1 = mk(0.0, 0.0);
  for (int which = 0; which < 2; ++which) {
    Dual xl = dlog(c.taumin), xr = dlog(spl_eval_g<Dual>(c.tau_of_a, mk(0.1, 0.0)));
    double fl = which == 0 ? cond_small_k_d(c, xl).v : cond_large_k_d(c, xl, k).v;
    for (int it = 0; it < 7; ++it) {
      const Dual xm = 0.5 * (xl + xr);
      const double fm = which == 0 ? cond_small_k_d(c, xm).v : cond_large_k_d(c, xm, k).v;
      if (fm * fl > 0) { xl = xm; fl = fm; } else xr = xm;
    }
    if (which == 0) res0 = 0.5 * (xl + xr); else res1 = 0.5 * (xl + xr);
  }
  return dexp(res0.v <= res1.v ? res0 : res1);
}

struct IcScalarsD { Dual a, deltag, thetag, deltar, thetar, shearr, deltaq, thetaq, eta; };
DEB_DEV IcScalarsD ic_scalars_d(const CosmoD& c, Dual tau, Dual k) {
  IcScalarsD s;
  const Dual a = spl_eval_g<Dual>(c.a_of_tau, tau);
  const Dual rn = dexp(spl_eval_g<Dual>(c.lrn, dlog(a)));
  const Dual a2 = a * a, a4 = a2 * a2;
  const Dual rhom = c.grhom * c.Omegam / (a2 * a);
  const Dual rhor = (c.grhog + c.grhor * (c.Neff + c.Nmnu * rn)) / a4;
  const Dual rhonu = c.grhor * (c.Neff + c.Nmnu * rn) / a4;
  const Dual fracb = c.Omegab / c.Omegam, fracnu = rhonu / rhor;
  const Dual om = a * rhom / dsqrt(rhor);
  const double ci = -1.0;
  const Dual kt = k * tau, kt2 = kt * kt;
  s.a = a;
  s.deltag = -kt2 / 3.0 * (1.0 - om * tau / 5.0) * ci;
  s.thetag = -(kt2 * kt) / tau / 36.0 * (1.0 - 3.0 * (1.0 + 5.0 * fracb - fracnu) / 20.0 / (1.0 - fracnu) * om * tau) * ci;
  s.deltar = s.deltag;
  s.thetar = -(kt2 * kt2) / tau / 36.0 / (4.0 * fracnu + 15.0)
           * (4.0 * fracnu + 11.0 + 12.0 - 3.0 * (8.0 * fracnu * fracnu + 50.0 * fracnu + 275.0) / 20.0 / (2.0 * fracnu + 15.0) * tau * om) * ci;
  s.shearr = kt2 / (45.0 + 12.0 * fracnu) * 2.0 * (1.0 + (4.0 * fracnu - 5.0) / 4.0 / (2.0 * fracnu + 15.0) * tau * om) * ci;
  const Dual wq = c.w0 + c.wa * (1.0 - a);
  s.deltaq = kt2 / 4.0 * (1.0 + wq) * (4.0 - 3.0 * c.cs2de) / (4.0 - 6.0 * wq + 3.0 * c.cs2de) * ci;
  s.thetaq = (kt2 * kt2) / tau / 4.0 * c.cs2de / (4.0 - 6.0 * wq + 3.0 * c.cs2de) * ci;
  s.eta = ci * (1.0 - kt2 / 12.0 / (15.0 + 4.0 * fracnu)
                * (5.0 + 4.0 * fracnu - (16.0 * fracnu * fracnu + 280.0 * fracnu + 325.0) / 10.0 / (2.0 * fracnu + 15.0) * tau * om));
  return s;
}
DEB_DEV Dual ic_value_d(const Problem& P, const CosmoD& c, const NuBins& nb, const IcScalarsD& s, int desc, Dual k) {
  const int type = desc & 0xff, chain = desc >> 16;
  switch (type) {
    case R_A: return s.a;
    case R_ETA: return s.eta;
    case R_DC: case R_DB: return 0.75 * s.deltag;
    case R_TB: case R_F1: return s.thetag;
    case R_F0: return s.deltag;
    case R_N0: return s.deltar;
    case R_N1: return s.thetar;
    case R_N2: return s.shearr * 2.0;
    case R_DQ: return s.deltaq;
    case R_TQ: return s.thetaq;
    case R_P0: case R_P1: case R_P2: {
      const int i = chain - 3;
      const Dual aq = s.a * c.amnu / nb.q[i];
      const Dual v = 1.0 / dsqrt(1.0 + aq * aq);
      const double dl = nb.dl[i];
      if (type == R_P0) return (-0.25 * dl) * s.deltar;
      if (type == R_P1) return (-dl) * s.thetar / v / k / 3.0;
      return (-0.5 * dl) * s.shearr;
    }
    default: return mk(0.0, 0.0);
  }
}

// ---- epilogue with tangents: 20 output fields (perturbations.py:374-523) and get_power (:1101-1123) ------
DEB_DEV void convert_outputs_d(const Problem& P, const CosmoD& c, const NuBins& nb, const StateVD& y, Dual k, Dual* out) {
  const int nq = P.nq, iq0 = P.iq0, n = P.n;
  const Dual a = sget<Dual>(y, 0), eta = sget<Dual>(y, 2), dc = sget<Dual>(y, 3), tc = sget<Dual>(y, 4), db = sget<Dual>(y, 5);
  const Dual tb = sget<Dual>(y, 6), dg = sget<Dual>(y, 7), tg = sget<Dual>(y, 8);
  const Dual dr = sget<Dual>(y, P.ir), tr = sget<Dual>(y, P.ir + 1), dq = sget<Dual>(y, n - 2), tq = sget<Dual>(y, n - 1);
  const Dual la = dlog(a);
  const Dual rhonu = dexp(spl_eval_g<Dual>(c.lrn, la)), pnu = dexp(spl_eval_g<Dual>(c.lpn, la));
  Dual drhonu = mk(0.0, 0.0), fnu = mk(0.0, 0.0);
  for (int i = 0; i < nq; ++i) {
    const Dual aq = a * c.amnu / nb.q[i];
    const Dual v = 1.0 / dsqrt(1.0 + aq * aq);
    drhonu = drhonu + nb.w[i] * sget<Dual>(y, iq0 + i) / v;
    fnu = fnu + nb.w[i] * sget<Dual>(y, iq0 + nq + i);
  }
  const Dual deltanu = drhonu / rhonu, thetanu = k * fnu / (rhonu + pnu);
  const Dual wq = c.w0 + c.wa * (1.0 - a);
  const Dual rhoq = dexp((-3.0 * (1.0 + c.w0 + c.wa)) * la + 3.0 * (a - 1.0) * c.wa);
  const Dual a2 = a * a;
  const Dual Omegac = c.Omegam - c.Omegab;
  const Dual rpt = (1.0 + wq) * rhoq * c.grhom * c.OmegaDE * tq * a2;
  const Dual grho = c.grhom * c.Omegam / a + (c.grhog + c.grhor * (c.Neff + c.Nmnu * rhonu)) / a2
                  + c.grhom * c.OmegaDE * rhoq * a2 + c.grhom * c.Omegak;
  const Dual H = dsqrt(grho / 3.0);
  const Dual mat = c.grhom * (Omegac * dc + c.Omegab * db) / a;
  const Dual matth = c.grhom * (Omegac * tc + c.Omegab * tb) / a;
  const Dual dgrho = mat + (c.grhog * dg + c.grhor * (c.Neff * dr + c.Nmnu * drhonu)) / a2 + c.grhom * c.OmegaDE * dq * rhoq * a2;
  const Dual dgtheta = matth + 4.0 / 3.0 * (c.grhog * tg + c.Neff * c.grhor * tr) / a2 + c.Nmnu * c.grhor * k * fnu / a2 + rpt;
  const Dual k2 = k * k;
  const Dual hp = (2.0 * k2 * eta + dgrho) / H, ep = 0.5 * dgtheta / k2, al = (hp + 6.0 * ep) / 2.0 / k2;
  const Dual deltam = (mat + (c.grhor * c.Nmnu * drhonu) / a2) / (c.grhom * c.Omegam / a + (c.grhor * c.Nmnu * rhonu) / a2);
  Dual thetam = (matth + c.Nmnu * c.grhor * k * fnu / a2) / (3.0 * (c.grhom * c.Omegam / a + c.grhor * c.Nmnu * rhonu / a2));
  const Dual deltabc = mat / (c.grhom * c.Omegam / a);
  Dual thetabc = matth / (3.0 * (c.grhom * c.Omegam / a) / a2);
  thetam = thetam + al * k2; thetabc = thetabc + al * k2;
  out[0] = eta; out[1] = ep; out[2] = hp; out[3] = al;
  out[4] = deltam; out[5] = thetam / H; out[6] = deltabc; out[7] = thetabc / H;
  out[8] = dc; out[9] = tc / H; out[10] = db; out[11] = tb / H; out[12] = dg; out[13] = tg / H;
  out[14] = dr; out[15] = tr / H; out[16] = deltanu; out[17] = thetanu / H; out[18] = dq; out[19] = tq / H;
}
DEB_DEV Dual power_d(const CosmoD& c, Dual k, Dual yv) {
  const Dual tilt = dexp((c.ns - 1.0) * dlog(k / c.kp));
  return (2.0 * 9.869604401089358) * c.As * tilt * (1.0 / (k * k * k)) * yv * yv;
}

// ---- per-mode tangent workspace (shared memory, carved after the primal workspace) ---------------------
struct TanWs {
  double* yd;     // tangent of the accepted state [np]
  double* ud;     // tangent of the stage state [np]
  double* rd;     // tangent right-hand side, solved in place -> kdot_i [np]
  double* jad;    // d/d eps of d f/d a at (t0, y0) [np]
  double* cc;     // primal sum_j C_ij/dt k_j of the current stage [np]
  double* kd;     // kdot_1 .. kdot_7 [7 np]
  double* b0;     // base-point coefficients BgG<Dual> at (a0, a0dot) [2 BG_NS]
  double* vv0;    // base-point v_i [4 NQMAX]
  double* vvi;    // stage v_i [4 NQMAX]
  double* sl0;    // base-point operator slots [4 NSLOT]
  double* sli;    // stage operator slots [4 NSLOT]
  double* kc0;    // base-point chain wavenumbers / dampings [4 NCHMAX] each
  double* kap0;
  double* kci;    // stage chain wavenumbers / dampings
  double* kapi;
  CosmoD* cd;     // scalars and tables of this (cosmology, direction)
};
DEB_HD size_t tan_ws_doubles(int np) {
  return (size_t)12 * np + 2 * BG_NS + 8 * NQMAX + 8 * NSLOT + 16 * NCHMAX + (sizeof(CosmoD) + 7) / 8;
}
DEB_DEV void carve_tan(TanWs& T, double* base, int np) {
  T.yd = base; T.ud = T.yd + np; T.rd = T.ud + np; T.jad = T.rd + np; T.cc = T.jad + np; T.kd = T.cc + np;
  T.b0 = T.kd + (size_t)7 * np; T.vv0 = T.b0 + 2 * BG_NS; T.vvi = T.vv0 + 4 * NQMAX;
  T.sl0 = T.vvi + 4 * NQMAX; T.sli = T.sl0 + 4 * NSLOT;
  T.kc0 = T.sli + 4 * NSLOT; T.kap0 = T.kc0 + 4 * NCHMAX; T.kci = T.kap0 + 4 * NCHMAX; T.kapi = T.kci + 4 * NCHMAX;
  T.cd = (CosmoD*)(T.kapi + 4 * NCHMAX);
}

// Everything of tangent stage `st` up to (not including) the linear solve: forms rdot_i in TW.rd.
//   y/yd  : accepted state and its tangent at (t, td)         u/ud : stage state u_i and its tangent
//   ki    : k_i of the primal stage (W.r after the primal solve)
template <class Dummy = void>
DEB_DEV void tan_stage_rhs(const Problem& P, const CtaConst& C, const WarpWs& W, const TanWs& TW, int st, Dual k,
                           double t, double td, double dt, double ddt, const Hints& hint DEB_LANE_PARAM) {
  const int n = P.n, np = P.np, nq = P.nq, nh = P.nh, nch = P.nch;
  const CosmoD& cd = *TW.cd;
  const NuBins& nb = C.nu;
  const double invdt = 1.0 / dt;
  const double wdiag = ddt / (RD_GAMMA * dt * dt);
  const double* ki = W.r();
  const double ki0 = ki[0];
  // ---- u_dot, and the part of rdot that needs no row evaluation:
  //      ddt/(gamma dt^2) k_i + (d/d eps df/da) k_i0 + sum_j C_ij (kdot_j/dt - k_j ddt/dt^2)
  DEB_LANES_BEGIN
    for (int e = lane; e < n; e += 32) {
      const double* k1 = TW.kd + e;
      double v, cs = 0.0;
      switch (st) {
        case 1: v = TW.yd[e]; break;
        case 2: v = TW.yd[e] + RD_A21 * k1[0];
                cs = RD_C21 * k1[0]; break;
        case 3: v = TW.yd[e] + RD_A31 * k1[0] + RD_A32 * k1[np];
                cs = RD_C31 * k1[0] + RD_C32 * k1[np]; break;
        case 4: v = TW.yd[e] + RD_A41 * k1[0] + RD_A42 * k1[np] + RD_A43 * k1[2 * np];
                cs = RD_C41 * k1[0] + RD_C42 * k1[np] + RD_C43 * k1[2 * np]; break;
        case 5: v = TW.yd[e] + RD_A51 * k1[0] + RD_A52 * k1[np] + RD_A53 * k1[2 * np] + RD_A54 * k1[3 * np];
                cs = RD_C51 * k1[0] + RD_C52 * k1[np] + RD_C53 * k1[2 * np] + RD_C54 * k1[3 * np]; break;
        case 6: v = TW.yd[e] + RD_A61 * k1[0] + RD_A62 * k1[np] + RD_A63 * k1[2 * np] + RD_A64 * k1[3 * np] + RD_A65 * k1[4 * np];
                cs = RD_C61 * k1[0] + RD_C62 * k1[np] + RD_C63 * k1[2 * np] + RD_C64 * k1[3 * np] + RD_C65 * k1[4 * np]; break;
        case 7: v = TW.ud[e] + k1[5 * np];
                cs = RD_C71 * k1[0] + RD_C72 * k1[np] + RD_C73 * k1[2 * np] + RD_C74 * k1[3 * np] + RD_C75 * k1[4 * np] + RD_C76 * k1[5 * np]; break;
        default: v = TW.ud[e] + k1[6 * np];
                cs = RD_C81 * k1[0] + RD_C82 * k1[np] + RD_C83 * k1[2 * np] + RD_C84 * k1[3 * np] + RD_C85 * k1[4 * np] + RD_C86 * k1[5 * np] + RD_C87 * k1[6 * np]; break;
      }
      TW.ud[e] = v;
      double r = wdiag * ki[e];
      if (st > 1) r += TW.jad[e] * ki0 + invdt * cs - (ddt * invdt) * TW.cc[e];
      TW.rd[e] = r;
    }
  DEB_LANES_END
  const double* us = st == 1 ? W.y() : W.u();
  const double ci = st == 2 ? RD_CT2 : st == 3 ? RD_CT3 : st == 4 ? RD_CT4 : st == 5 ? RD_CT5 : (st == 1 ? 0.0 : 1.0);
  const double di = st == 1 ? RD_D1 : st == 2 ? RD_D2 : st == 3 ? RD_D3 : st == 4 ? RD_D4 : st == 5 ? RD_D5 : 0.0;
  const Dual tsD = mk(t + ci * dt, td + ci * ddt);
  const Dual t0D = mk(t, td);
  const Dual invts = 1.0 / tsD, invt0 = 1.0 / t0D;
  const double it2 = 1.0 / (t * t);
  const StateV0 sk = {ki};
  const StateVD su = {us, TW.ud};
  BgG<Dual> b0;
  MetricG<Dual> m0;
  if (st == 1) {
    // base-point coefficients (a0, a0dot), once per step: kept in shared memory for the other seven stages
    compute_bg_g<Dual>(cd, mk(W.y()[0], TW.yd[0]), k, hint.th, hint.nu, b0);
    // second-order duals at (t0, y0): eps-part = fdot, (a x eps)-part = d/d eps of df/da
    BgG<HD> bh;
    const HD aH = mkd<Dual>(mk(W.y()[0], TW.yd[0]), mk(1.0, 0.0));
    compute_bg_g<HD>(cd, aH, k, hint.th, hint.nu, bh);
    DEB_LANES_BEGIN
      if (lane == 0) { bg_store(TW.b0, b0); fill_slots_g<Dual>(cd, b0, k, TW.sl0); fill_slots_g<HD>(cd, bh, k, TW.sli); }
      if (lane < nq) { nu_velocity_g<Dual>(cd, nb, b0.a, lane, TW.vv0 + 4 * lane); nu_velocity_g<HD>(cd, nb, aH, lane, TW.vvi + 4 * lane); }
    DEB_LANES_END
    DEB_LANES_BEGIN
      if (lane < nch) {
        chain_coeffs_g<Dual>(b0, TW.vv0, k, lane, TW.kc0, TW.kap0, TW.sl0);
        chain_coeffs_g<HD>(bh, TW.vvi, k, lane, TW.kci, TW.kapi, TW.sli);
      }
    DEB_LANES_END
    compute_metric_g<Dual>(P, cd, nb, b0, TW.vv0, sk, k, m0);
    MetricG<HD> mh;
    compute_metric_g<HD>(P, cd, nb, bh, TW.vvi, su, k, mh);
    const HD invtH = liftD<HD>(invt0);
    DEB_LANES_BEGIN
      if (lane < nh) {
        const int e = C.hidx[lane];
        const HD f = head_row_g<HD>(C, TW.sli, su, lane, mh);
        const Dual g = head_row_g<Dual>(C, TW.sl0, sk, lane, m0);
        TW.jad[e] = f.d.d;
        TW.rd[e] += f.v.d + g.d + f.d.d * ki0;
      }
      if (lane == 0) { const HD f = bh.H * bh.a; TW.jad[0] = f.d.d; TW.rd[0] += f.v.d + f.d.d * ki0; }
      for (int tt = lane; tt < C.ntail; tt += 32) {
        int e; double tr;
        const HD f = tail_row_g<HD>(C, TW.kci, TW.kapi, su, C.tail[tt], invtH, &e, &tr);
        const Dual g = tail_row_g<Dual>(C, TW.kc0, TW.kap0, sk, C.tail[tt], invt0, &e, &tr);
        TW.jad[e] = f.d.d;
        TW.rd[e] += f.v.d + g.d + f.d.d * ki0
                  + di * (ddt * (tr * it2 * W.y()[e]) + dt * tr * (TW.yd[e] * it2 - 2.0 * W.y()[e] * td * it2 / t));
      }
    DEB_LANES_END
    return;
  }
  bg_load(TW.b0, b0);
  BgG<Dual> bi;
  compute_bg_g<Dual>(cd, mk(us[0], TW.ud[0]), k, hint.th, hint.nu, bi);
  DEB_LANES_BEGIN
    if (lane == 0) fill_slots_g<Dual>(cd, bi, k, TW.sli);
    if (lane < nq) nu_velocity_g<Dual>(cd, nb, bi.a, lane, TW.vvi + 4 * lane);
  DEB_LANES_END
  DEB_LANES_BEGIN
    if (lane < nch) chain_coeffs_g<Dual>(bi, TW.vvi, k, lane, TW.kci, TW.kapi, TW.sli);
  DEB_LANES_END
  compute_metric_g<Dual>(P, cd, nb, b0, TW.vv0, sk, k, m0);
  MetricG<Dual> mi;
  compute_metric_g<Dual>(P, cd, nb, bi, TW.vvi, su, k, mi);
  const double dtrunc = st <= 5 ? di : 0.0;
  DEB_LANES_BEGIN
    if (lane < nh) {
      const int e = C.hidx[lane];
      TW.rd[e] += head_row_g<Dual>(C, TW.sli, su, lane, mi).d + head_row_g<Dual>(C, TW.sl0, sk, lane, m0).d;
    }
    if (lane == 0) TW.rd[0] += (bi.H * bi.a).d;
    for (int tt = lane; tt < C.ntail; tt += 32) {
      int e; double tr;
      const Dual f = tail_row_g<Dual>(C, TW.kci, TW.kapi, su, C.tail[tt], invts, &e, &tr);
      const Dual g = tail_row_g<Dual>(C, TW.kc0, TW.kap0, sk, C.tail[tt], invt0, &e, &tr);
      TW.rd[e] += f.d + g.d + dtrunc * (ddt * (tr * it2 * W.y()[e]) + dt * tr * (TW.yd[e] * it2 - 2.0 * W.y()[e] * td * it2 / t));
    }
  DEB_LANES_END
}

}  // namespace deb
