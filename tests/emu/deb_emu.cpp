// deb_emu.cpp -- TEST INFRASTRUCTURE: compiles the CUDA kernel source (deb_core.cuh) as plain C++
// with the 32 lanes of a warp executed by loops, so that the CPU test-suite (-m "not gpu") can run
// the very same per-mode integrator against the oracle.  It is built on demand by tests/conftest.py
// into tests/emu/_build/, is never imported by the discoeb_b200 package and is not a fallback:
// the product path fails loudly when the CUDA library is missing.
#define DEB_CPU_EMU 1
#include <cstdlib>
#include <cstring>
#include <omp.h>
#include <vector>
#include "../../include/discoeb_b200.h"
#include "../../disco-eb_b200/csrc/deb_core.cuh"
#include "../../disco-eb_b200/csrc/deb_team.cuh"
#include "../../disco-eb_b200/csrc/deb_lane.cuh"
#include "../../disco-eb_b200/csrc/deb_background.cuh"
#include "../../disco-eb_b200/csrc/deb_spectra.cuh"
#include "../../disco-eb_b200/csrc/deb_host.inl"

using namespace deb;

static std::vector<double> g_lt_small;       // per call; the harness is single-caller
static void tau_out_host(Problem& P, double* tau_out, double* dtau_out = nullptr) {
  g_lt_small.assign(P.ncosmo, 0.0);
  for (int c = 0; c < P.ncosmo; ++c) {
    Spl s = get_spline(P, c, T_TAU_OF_A);
    if (P.aexp_out) for (int j = 0; j < P.nout; ++j) tau_out[(size_t)c * P.nout + j] = spl_eval(s, P.aexp_out[j]);
    Cosmo cs = load_cosmo(P, c);
    g_lt_small[c] = start_small_k(cs);
    if (dtau_out && P.aexp_out)
      for (int tn = 0; tn < P.ntan; ++tn) {
        SplT st; st.p = s; st.t = get_spline_from(P.d_tables, P, tn * P.ncosmo + c, T_TAU_OF_A);
        for (int j = 0; j < P.nout; ++j)
          dtau_out[((size_t)tn * P.ncosmo + c) * P.nout + j] = spl_eval_g<Dual>(st, mk(P.aexp_out[j], 0.0)).d;
      }
  }
  P.lt_small = g_lt_small.data();
}

// tangent path: one work item per (direction, cosmology, k)
template <int NE>
static void run_all_tan(const Problem& P) {
  CtaConst C;
  std::vector<int> tail(P.np);
  for (int t = 0; t < 32; ++t) init_cta_const(P, C, tail.data(), t, 32);
  std::vector<double> ws(warp_ws_doubles(P.np) + tan_ws_doubles(P.np));
  const int modes = P.ncosmo * P.nk, total = modes * P.ntan;
#pragma omp parallel for schedule(dynamic, 1) firstprivate(ws)
  for (int it = 0; it < total; ++it) {
    WarpWs W;
    carve(W, ws.data(), P.np);
    TanWs TW;
    carve_tan(TW, ws.data() + warp_ws_doubles(P.np), P.np);
    HelpBox box;
    integrate_mode<NE, false, true>(P, C, W, &box, modes - 1 - it % modes, &TW, it / modes);
  }
}
static int dispatch_tan(const Problem& P) {
  int ne = (P.n + 31) / 32;
  if (ne <= 3) run_all_tan<3>(P);
  else if (ne <= 4) run_all_tan<4>(P);
  else if (ne <= 6) run_all_tan<6>(P);
  else if (ne <= 9) run_all_tan<9>(P);
  else if (ne <= 12) run_all_tan<12>(P);
  else return DEB_E_UNSUPPORTED;
  return DEB_OK;
}

template <int NE, bool HELPER>
static void run_all(const Problem& P) {
  CtaConst C;
  std::vector<int> tail(P.np);
  for (int t = 0; t < 32; ++t) init_cta_const(P, C, tail.data(), t, 32);
  std::vector<double> ws(warp_ws_doubles(P.np));
  const int total = P.ncosmo * P.nk;
#pragma omp parallel for schedule(dynamic, 1) firstprivate(ws)
  for (int m = 0; m < total; ++m) {
    WarpWs W;
    carve(W, ws.data(), P.np);
    HelpBox box;
    integrate_mode<NE, HELPER>(P, C, W, &box, total - 1 - m);
  }
}

// DEB_EMU_HELPER=1 selects the two-warp (main + helper) code path, the helper being run inline
template <bool HELPER>
static int dispatch_h(const Problem& P) {
  int ne = (P.n + 31) / 32;
  if (ne <= 3) run_all<3, HELPER>(P);
  else if (ne <= 4) run_all<4, HELPER>(P);
  else if (ne <= 6) run_all<6, HELPER>(P);
  else if (ne <= 9) run_all<9, HELPER>(P);
  else if (ne <= 12) run_all<12, HELPER>(P);
  else return DEB_E_UNSUPPORTED;
  return DEB_OK;
}
// batched variant: the B warps of a batch are played by B OpenMP threads, the cluster barrier by `omp barrier`
template <int NE>
static int run_batched(const Problem& P) {
  CtaConst C;
  std::vector<int> tail(P.np);
  for (int t = 0; t < 32; ++t) init_cta_const(P, C, tail.data(), t, 32);
  const int B = P.batch_size, nb = P.ncosmo * P.nk / B;
  int bad = 0;
  omp_set_dynamic(0);
  for (int b = 0; b < nb; ++b) {
    std::vector<double> shared((size_t)2 * B * 2, 0.0);
#pragma omp parallel num_threads(B)
    {
      if (omp_get_num_threads() != B) {
#pragma omp atomic
        bad += 1;
      } else {
        std::vector<double> ws(warp_ws_doubles(P.np));
        WarpWs W;
        carve(W, ws.data(), P.np);
        BatchCtx bc;
        bc.slots = nullptr; bc.all = shared.data(); bc.B = B; bc.bw = B; bc.ncta = 1; bc.idx = omp_get_thread_num(); bc.parity = 0;
        HelpBox box;
        integrate_mode<NE, false, false, true>(P, C, W, &box, b * B + bc.idx, nullptr, 0, &bc);
      }
    }
  }
  return bad ? DEB_E_UNSUPPORTED : DEB_OK;
}
static int dispatch_batched(const Problem& P) {
  int ne = (P.n + 31) / 32;
  if (ne <= 3) return run_batched<3>(P);
  if (ne <= 4) return run_batched<4>(P);
  if (ne <= 6) return run_batched<6>(P);
  if (ne <= 9) return run_batched<9>(P);
  if (ne <= 12) return run_batched<12>(P);
  return DEB_E_UNSUPPORTED;
}
// DEB_EMU_TEAM=T selects the CTA-per-mode variant (deb_team.cuh) with T warps per mode, the threads run by loops
template <int NE, int TEAM>
static void run_all_team(const Problem& P) {
  CtaConst C;
  std::vector<int> tail(P.np);
  for (int t = 0; t < 32; ++t) init_cta_const(P, C, tail.data(), t, 32);
  std::vector<int> eslot(P.np), epos(P.np), tailpos(P.np);
  for (int t = 0; t < 32; ++t) init_team_const(P, C, eslot.data(), epos.data(), tailpos.data(), t, 32);
  const int segrows = team_segrows(P.lmaxg, P.lmaxgp, P.lmaxr, P.lmaxnu);
  std::vector<double> ws(warp_ws_doubles(P.np) + team_ws_doubles(segrows));
  const int total = P.ncosmo * P.nk;
#pragma omp parallel for schedule(dynamic, 1) firstprivate(ws)
  for (int m = 0; m < total; ++m) {
    WarpWs W;
    carve(W, ws.data(), P.np);
    TeamWs X;
    carve_team(X, ws.data() + warp_ws_doubles(P.np), segrows, eslot.data(), epos.data(), tailpos.data());
    TeamBox box;
    integrate_mode_team<NE, TEAM>(P, C, W, X, box, total - 1 - m);
  }
}
template <int TEAM>
static int dispatch_team(const Problem& P) {
  const int ne = (P.n + 32 * TEAM - 1) / (32 * TEAM);
  if (ne <= 1) run_all_team<1, TEAM>(P);
  else if (ne <= 2) run_all_team<2, TEAM>(P);
  else if (ne <= 3) run_all_team<3, TEAM>(P);
  else if (ne <= 4) run_all_team<4, TEAM>(P);
  else if (ne <= 6) run_all_team<6, TEAM>(P);
  else return DEB_E_UNSUPPORTED;
  return DEB_OK;
}
// DEB_EMU_LANE=1 selects the register-resident chain-lane variant (deb_lane.cuh)
template <int NT>
static void run_all_lane(const Problem& P) {
  CtaConst C;
  std::vector<int> tail(P.np);
  for (int t = 0; t < 32; ++t) init_cta_const(P, C, tail.data(), t, 32);
  std::vector<LaneTab<NT>> tab(1);
  for (int t = 0; t < 32; ++t) init_lane_tab<NT>(P, C, tab[0], t, 32);
  const int total = P.ncosmo * P.nk;
#pragma omp parallel
  {
    std::vector<LaneWs<NT>> ws(1);
    LaneSync sy; sy.cnt = 32; sy.on = 0; sy.every = 1;
#pragma omp for schedule(dynamic, 1)
    for (int m = 0; m < total; ++m) integrate_mode_lane<NT>(P, C, tab[0], ws[0], sy, total - 1 - m);
  }
}
static int dispatch_lane(const Problem& P) {
  if (LN_NSEG * P.nch > 32 || P.nh > 32) return DEB_E_UNSUPPORTED;
  switch (lane_nt(P.lmaxg, P.lmaxgp, P.lmaxr, P.lmaxnu)) {
    case 1: run_all_lane<1>(P); break;
    case 2: run_all_lane<2>(P); break;
    case 3: run_all_lane<3>(P); break;
    case 4: run_all_lane<4>(P); break;
    case 6: run_all_lane<6>(P); break;
    case 8: run_all_lane<8>(P); break;
    default: return DEB_E_UNSUPPORTED;
  }
  return DEB_OK;
}
static int dispatch(const Problem& P) {
  if (P.batch_size > 0) return dispatch_batched(P);
  if (const char* ln = getenv("DEB_EMU_LANE")) { if (ln[0] == '1') return dispatch_lane(P); }
  if (const char* tm = getenv("DEB_EMU_TEAM")) {
    const int T = atoi(tm);
    if (T == 2) return dispatch_team<2>(P);
    if (T == 4) return dispatch_team<4>(P);
    if (T == 8) return dispatch_team<8>(P);
  }
  const char* h = getenv("DEB_EMU_HELPER");
  return (h && h[0] == '1') ? dispatch_h<true>(P) : dispatch_h<false>(P);
}

extern "C" int emu_evolve_host_f64(const deb_dims* dims, const deb_ctrl* ctrl, const double* scalars,
                                   const double* tables, const double* kmodes, const double* aexp_out,
                                   double* y_out, double* pk_out, double* tau_out, int32_t* status,
                                   int32_t* nsteps, int32_t* naccept) {
  Problem P;
  int rc = fill_problem(dims, ctrl, &P);
  if (rc) return rc;
  P.scalars = scalars; P.tables = tables; P.kmodes = kmodes; P.aexp_out = aexp_out;
  P.y_out = y_out; P.pk_out = pk_out; P.status = status; P.nsteps = nsteps; P.naccept = naccept;
  tau_out_host(P, tau_out);
  P.tau_out = tau_out;
  P.mode = 0;
  return dispatch(P);
}

extern "C" int emu_debug_step_host_f64(const deb_dims* dims, const double* scalars, const double* tables,
                                       const double* kmodes, const double* t0, const double* t1,
                                       const double* y0, double* y1, double* yerr) {
  Problem P;
  deb_ctrl ctrl = {1e-4, 1e-4, 0.25, 0.8, 0.0, 20.0, 0.3, 0.9};
  deb_dims d = *dims;
  d.nout = 1;
  int rc = fill_problem(&d, &ctrl, &P);
  if (rc) return rc;
  const int total = P.ncosmo * P.nk;
  std::vector<double> tau_out(P.ncosmo, 1.0);
  std::vector<int32_t> st(total), ns(total);
  P.scalars = scalars; P.tables = tables; P.kmodes = kmodes; P.aexp_out = nullptr;
  tau_out_host(P, tau_out.data());
  P.tau_out = tau_out.data(); P.status = st.data(); P.nsteps = ns.data(); P.naccept = nullptr;
  P.dbg_t0 = t0; P.dbg_t1 = t1; P.dbg_y0 = y0; P.dbg_y1 = y1; P.dbg_err = yerr;
  P.mode = 1;
  return dispatch(P);
}

extern "C" int emu_debug_ics_host_f64(const deb_dims* dims, const double* scalars, const double* tables,
                                      const double* kmodes, const double* aexp_out, double* tau_start, double* y0) {
  Problem P;
  deb_ctrl ctrl = {1e-4, 1e-4, 0.25, 0.8, 0.0, 20.0, 0.3, 0.9};
  int rc = fill_problem(dims, &ctrl, &P);
  if (rc) return rc;
  const int total = P.ncosmo * P.nk;
  std::vector<double> tau_out((size_t)P.ncosmo * P.nout);
  std::vector<int32_t> st(total), ns(total);
  P.scalars = scalars; P.tables = tables; P.kmodes = kmodes; P.aexp_out = aexp_out;
  tau_out_host(P, tau_out.data());
  P.tau_out = tau_out.data(); P.status = st.data(); P.nsteps = ns.data(); P.naccept = nullptr;
  P.dbg_tau_start = tau_start; P.dbg_ics = y0;
  P.mode = 2;
  return dispatch(P);
}

extern "C" int emu_debug_replay_host_f64(const deb_dims* dims, const deb_ctrl* ctrl, const double* scalars,
                                         const double* tables, const double* kmodes, const double* aexp_out,
                                         const double* rp_tnext, const int32_t* rp_keep, const int32_t* rp_n,
                                         int32_t rp_stride, double* y_out, int32_t* nsteps) {
  Problem P;
  int rc = fill_problem(dims, ctrl, &P);
  if (rc) return rc;
  const int total = P.ncosmo * P.nk;
  std::vector<double> tau_out((size_t)P.ncosmo * P.nout);
  std::vector<int32_t> st(total);
  P.scalars = scalars; P.tables = tables; P.kmodes = kmodes; P.aexp_out = aexp_out;
  tau_out_host(P, tau_out.data());
  P.tau_out = tau_out.data(); P.status = st.data(); P.nsteps = nsteps; P.naccept = nullptr; P.y_out = y_out;
  P.rp_tnext = rp_tnext; P.rp_keep = rp_keep; P.rp_n = rp_n; P.rp_stride = rp_stride;
  P.mode = 3;
  return dispatch(P);
}

extern "C" int emu_evolve_tangent_host_f64(const deb_dims* dims, const deb_ctrl* ctrl, const double* scalars,
                                           const double* tables, const double* kmodes, const double* aexp_out,
                                           const double* d_scalars, const double* d_tables, const double* d_kmodes,
                                           double* y_out, double* dy_out,
                                           double* pk_out, double* dpk_out, double* tau_out, double* dtau_out,
                                           int32_t* status, int32_t* nsteps, int32_t* naccept) {
  Problem P;
  int rc = fill_problem(dims, ctrl, &P);
  if (rc) return rc;
  if (P.ntan < 1) return DEB_E_ARG;
  P.scalars = scalars; P.tables = tables; P.kmodes = kmodes; P.aexp_out = aexp_out; P.d_scalars = d_scalars; P.d_tables = d_tables;
  P.y_out = y_out; P.dy_out = dy_out; P.pk_out = pk_out; P.dpk_out = dpk_out; P.status = status; P.nsteps = nsteps; P.naccept = naccept;
  P.d_kmodes = d_kmodes;
  tau_out_host(P, tau_out, dtau_out);
  P.tau_out = tau_out; P.dtau_out = dtau_out;
  P.mode = 0;
  return dispatch_tan(P);
}

extern "C" int emu_debug_replay_tangent_host_f64(const deb_dims* dims, const deb_ctrl* ctrl, const double* scalars,
                                                 const double* tables, const double* kmodes, const double* aexp_out,
                                                 const double* d_scalars, const double* d_tables, const double* d_kmodes,
                                                 const double* rp_tnext,
                                                 const double* rp_dtnext, const int32_t* rp_keep, const int32_t* rp_n,
                                                 int32_t rp_stride, double* y_out, double* dy_out, double* dtau_out,
                                                 int32_t* nsteps) {
  Problem P;
  int rc = fill_problem(dims, ctrl, &P);
  if (rc) return rc;
  if (P.ntan < 1) return DEB_E_ARG;
  const int total = P.ncosmo * P.nk;
  std::vector<double> tau_out((size_t)P.ncosmo * P.nout), dtau((size_t)P.ntan * P.ncosmo * P.nout);
  std::vector<int32_t> st(total);
  P.scalars = scalars; P.tables = tables; P.kmodes = kmodes; P.aexp_out = aexp_out; P.d_scalars = d_scalars; P.d_tables = d_tables;
  tau_out_host(P, tau_out.data(), dtau.data());
  P.tau_out = tau_out.data(); P.dtau_out = dtau.data();
  P.status = st.data(); P.nsteps = nsteps; P.naccept = nullptr; P.y_out = y_out; P.dy_out = dy_out; P.power_idx = -1;
  P.d_kmodes = d_kmodes;
  P.rp_tnext = rp_tnext; P.rp_dtnext = rp_dtnext; P.rp_keep = rp_keep; P.rp_n = rp_n; P.rp_stride = rp_stride;
  P.mode = 3;
  rc = dispatch_tan(P);
  if (dtau_out) memcpy(dtau_out, dtau.data(), dtau.size() * sizeof(double));
  return rc;
}

// table producer (deb_background.cuh) on the CPU: one cosmology per OpenMP thread
extern "C" int emu_background_host_ex_f64(int32_t device, int32_t ncosmo, int32_t nth, const double* bg_in, double* scalars, double* tables,
                                          double* extras, float* kernel_ms);
extern "C" int emu_background_host_f64(int32_t device, int32_t ncosmo, int32_t nth, const double* bg_in, double* scalars, double* tables,
                                       float* kernel_ms) {
  return emu_background_host_ex_f64(device, ncosmo, nth, bg_in, scalars, tables, nullptr, kernel_ms);
}
extern "C" size_t emu_background_extras_len(int32_t nth) { return deb::bg::extras_len(nth); }
extern "C" int emu_background_host_ex_f64(int32_t device, int32_t ncosmo, int32_t nth, const double* bg_in, double* scalars, double* tables,
                                          double* extras, float* kernel_ms) {
  (void)device;
  using namespace deb::bg;
  if (ncosmo < 1 || nth < 16 || nth > NTH_MAX) return DEB_E_ARG;
  double q[NNUQ], w[NNUQ];
  nu_quadrature(q, w);
  const size_t tl = 3 * (size_t)(5 * nth + 2 * NNU);
#pragma omp parallel for schedule(dynamic, 1)
  for (int c = 0; c < ncosmo; ++c) {
    std::vector<BgWork> W(1);
    for (int i = 0; i < DEB_NSCAL; ++i) scalars[(size_t)c * DEB_NSCAL + i] = 0.0;
    background_one(bg_in + (size_t)c * NBGIN, q, w, nth, scalars + (size_t)c * DEB_NSCAL, tables + (size_t)c * tl, W[0], 0, 1,
                   extras ? extras + (size_t)c * extras_len(nth) : nullptr);
  }
  if (kernel_ms) *kernel_ms = 0.0f;
  return DEB_OK;
}

// spectra epilogues (deb_spectra.cuh) on the CPU
extern "C" int emu_spectra_host_f64(int32_t device, int32_t nk, int32_t nmu, const double* y, const double* kmodes, double As, double ns, double kp,
                                    double bias, const double* sg_coef, int32_t sg_window, const double* mu, int32_t ell, double* P0, double* P2,
                                    double* P4, double* Pkmu, double* Ps_delta, double* Ps_theta, double* xi, double* r) {
  (void)device;
  using namespace deb::sp;
  std::vector<double> dm(nk), tm(nk), lPd(nk), lPt(nk), Pd(nk);
  for (int i = 0; i < nk; ++i) point(i, y, kmodes, As, ns, kp, bias, dm.data(), tm.data(), P0, P2, P4, lPd.data(), lPt.data(), Pd.data());
  if (sg_window > 0) for (int i = 0; i < nk; ++i) { Ps_delta[i] = smooth_at(i, nk, lPd.data(), sg_coef, sg_window); Ps_theta[i] = smooth_at(i, nk, lPt.data(), sg_coef, sg_window); }
  if (Pkmu && nmu > 0 && mu)
    for (int i = 0; i < nk; ++i) for (int j = 0; j < nmu; ++j) {
      const double d = sg_window > 0 ? sqrt(Ps_delta[i]) : dm[i], t = sg_window > 0 ? -sqrt(Ps_theta[i]) : tm[i];
      const double v = bias * d - mu[j] * mu[j] * t;
      Pkmu[(size_t)i * nmu + j] = v * v;
    }
  if (xi && r) {
    const double* Pk = sg_window > 0 ? Ps_delta : Pd.data();
    std::vector<Cx> F(nk / 2 + 1);
    for (int m = 0; m <= nk / 2; ++m) F[m] = fftlog_forward(m, nk, kmodes, Pk, ell);
    for (int nn = 0; nn < nk; ++nn) { xi[nk - 1 - nn] = fftlog_backward(nn, nk, kmodes, F.data(), ell); r[nk - 1 - nn] = 2.0 * M_PI / kmodes[nn]; }
  }
  return DEB_OK;
}
