// deb_lane.cuh -- register-resident "chain-lane" variant of the per-mode integrator (one warp per mode).
//
// Same algorithm and factorisation as integrate_mode (deb_core.cuh; reference map in include/discoeb_b200.h), other
// data layout.  integrate_mode spreads the n state variables cyclically over the lanes and keeps every working vector
// in shared memory: each tri-diagonal sweep is 29 dependent steps on 8 of 32 lanes, every hierarchy row decodes its
// (chain, l) descriptor and gathers three neighbours from shared memory -- 25 k warp-instructions per attempted step
// at n = 265, LSU 26 % busy, 4e9 bank-conflict cycles per 4096-mode launch (profiles/r2_v24_k_evolve9x4_*).  Here
//
//   * lane = segment * nch + chain owns NT CONSECUTIVE multipoles l = 3 + segment*NT + j of ONE hierarchy ("chain":
//     photon T, photon pol., massless nu, one massive-nu momentum bin) -- 4 segments x (3 + nq) chains = 32 lanes at
//     nq = 5 -- and lane h additionally owns row h of the "head" (metric, fluids, l <= 2 of every chain; <= 32 rows);
//   * the stage state u, the right-hand side / solution and k_1..k_7 of a lane's own rows live in REGISTERS for the
//     whole step; the l-1 / l+1 neighbours of a hierarchy row are the lane's own registers, except at the two segment
//     ends (one shuffle each per stage);
//   * both sweeps of the bordered solve are segmented by construction: every lane runs the first-order recurrence over
//     its NT rows from a zero carry, a 3-round shuffle chain supplies the true carries, a fix-up with the per-step
//     cumulative multiplier products applies them (NT + 3 + 1 dependent steps instead of 29, no shared-memory traffic
//     for the vector being solved);
//   * tail factors (m, 1/e, g and their cumulative products) and the accepted state are stored TRANSPOSED, row j of
//     lane ln at j*32 + ln: every access is one conflict-free 256-byte row;
//   * the d f/d a column is folded into the row evaluation (x_0 = r_0 / W_00 closes row 0 before any other row is
//     formed), so the solve touches one vector.
//
// Head build, block inverses and the Woodbury correction are those of integrate_mode.  Debug modes 1-3 are supported;
// tangents and shared-step batches keep integrate_mode.  Compiles for the device and, with -DDEB_CPU_EMU, as plain C++
// (tests/emu: test infrastructure).
#pragma once

namespace deb {

constexpr int LN_NSEG = 4;

// Lock-step of the warps of a CTA (each on its own mode).  The step loop is ~170 KB of straight-line code executed
// once per attempted step: 8 warps in 8 different places stream it from L2 eight times (instruction-fetch stalls were
// 22-45 % of the issue cycles, profiles/r2_lane_v1/v2_*); warps that pass the same point together share every fetched
// line in the SM's instruction cache.  LN_BAR is a counted CTA barrier at the top of the Jacobian phase and of every
// stage, executed by ALL threads of the CTA every time; `barrier.red.popc` returns how many threads still have work.  A
// warp whose work queue ran dry keeps attending (it arrives at once and sleeps in the barrier) until that count is zero.
struct LaneSync { unsigned cnt; int on; int every; };     // every: a stage barrier at stages 1, 1 + every, ... (1, 2, 4 or 8)
#ifdef DEB_CPU_EMU
#define LN_BAR(SY, stay)
#else
#define LN_BAR(SY, stay) do { if ((SY).on) { unsigned nc_; asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\tbarrier.red.popc.u32 %0, 1, p;\n\t}" \
  : "=r"(nc_) : "r"((unsigned)(stay)) : "memory"); (SY).cnt = nc_; } } while (0)
#endif

#ifdef DEB_CPU_EMU
#define LN_BEGIN for (int lane = 0; lane < 32; ++lane) {
#define LN_END }
#define LN_SHFL(name, src) (name##_all[(src) & 31])
#else
#define LN_BEGIN {
#define LN_END }
#define LN_SHFL(name, src) __shfl_sync(0xffffffffu, name, (src) & 31)
#endif

// Per-warp shared-memory block with a COMPILE-TIME layout: every access is [one base register + immediate] (the first
// version carved pointers from the run-time state size and spent 3 k instructions per step re-deriving them).
template <int NT> struct alignas(16) LaneWs {
  static constexpr int TA = NT * 32;
  static constexpr int NPCAP = 34 + 32 * NT;         // >= n for every layout this NT serves: a + 32 head rows + 32 NT tail rows
  double yt[TA], mt[TA], iet[TA], gt[TA];           // transposed tails: accepted state, backward multipliers, 1/pivot, forward multipliers
  double pct[TA], qct[TA], jat[TA];                 // cumulative multiplier products of a segment; d f/d a of the tail rows
  double y_[NPCAP], u_[NPCAP];                      // reference layout: a, head entries and the l = 3 entries are kept current
  double rh[NHMAX], jah[NHMAX];                     // head right-hand side -> solution; d f/d a of the head rows (by head position)
  double lu_[NHMAX * LDB], gh_[NHMAX], ge_[NHMAX], j1_[NHMAX], qh_[NHMAX], qe_[NHMAX], xb_[NHMAX + 8];
  double kc_[2 * NCHMAX], kap_[2 * NCHMAX], nur_[2 * NQMAX], nup_[2 * NQMAX], sl_[2 * NSLOT], ic_[ICACHE];
  double m2[NCHMAX], sch[NCHMAX];                   // per chain: backward multiplier of its l = 2 (head) row, Schur increment of that row
  double ka0[8];                                    // element 0 (scale factor) of k_1..k_7
  double khs[7 * 32];                               // k_1..k_7 of the head rows: column `lane` is private to the lane that owns the row
  int perm_[NHMAX];
  Cosmo cosmo_;
  // interpolated output state (reference layout): pct/qct/jat are dead at the end of a step, 3 TA >= NPCAP doubles
  DEB_DEV double* o_() { return reinterpret_cast<double*>(this) + 4 * TA; }      // = pct, without its array bound
  DEB_DEV double* tails() { return reinterpret_cast<double*>(this); }               // the seven transposed arrays, contiguous
};
template <int NT> DEB_DEV double* lrow(LaneWs<NT>& W, int r, int lo) { return W.lu_ + (r * LDB - lo); }
// launch-constant coefficient rows of the hierarchy equations, transposed like the tail arrays (one copy per CTA):
// row j of lane ln multiplies  kc (cl u_{l-1} - ch u_{l+1}); the truncation row has (1, 0), rows past lmax (0, 0)
template <int NT> struct alignas(16) LaneTab { double cl[NT * 32], ch[NT * 32]; };
template <int NT> DEB_DEV void init_lane_tab(const Problem& P, const CtaConst& C, LaneTab<NT>& T, int tid, int nthreads) {
  for (int i = tid; i < NT * 32; i += nthreads) {
    const int j = i >> 5, ln = i & 31;
    double cl = 0.0, ch = 0.0;
    if (ln < LN_NSEG * P.nch) {
      const int sg = ln / P.nch, c = ln - sg * P.nch, L = C.ch_lmax[c], l = 3 + sg * NT + j;
      if (l < L) { cl = (double)l / (double)(2 * l + 1); ch = (double)(l + 1) / (double)(2 * l + 1); }
      else if (l == L) { cl = 1.0; ch = 0.0; }
    }
    T.cl[i] = cl; T.ch[i] = ch;
  }
}
// smallest supported rows-per-lane for a launch (the kernels are instantiated for NT in {1, 2, 3, 4, 6, 8})
DEB_HD int lane_nt(int lmaxg, int lmaxgp, int lmaxr, int lmaxnu) {
  int L = lmaxg > lmaxgp ? lmaxg : lmaxgp;
  if (lmaxr > L) L = lmaxr;
  if (lmaxnu > L) L = lmaxnu;
  const int need = (L - 2 + LN_NSEG - 1) / LN_NSEG;
  const int opts[6] = {1, 2, 3, 4, 6, 8};
  for (int i = 0; i < 6; ++i) if (opts[i] >= need) return opts[i];
  return 0;
}

// Rodas5 stage combinations over a lane's own rows: kt[jj][j] (tails), kh[jj] (head row), ka0[jj] (scale factor)
#define LN_COMB_A(st, K0, K1, K2, K3, K4) \
  ((st) == 2 ? RD_A21 * (K0) : (st) == 3 ? RD_A31 * (K0) + RD_A32 * (K1) : (st) == 4 ? RD_A41 * (K0) + RD_A42 * (K1) + RD_A43 * (K2) \
   : (st) == 5 ? RD_A51 * (K0) + RD_A52 * (K1) + RD_A53 * (K2) + RD_A54 * (K3) \
   : RD_A61 * (K0) + RD_A62 * (K1) + RD_A63 * (K2) + RD_A64 * (K3) + RD_A65 * (K4))

template <int NT>
DEB_DEV void integrate_mode_lane(const Problem& P, const CtaConst& C, const LaneTab<NT>& T, LaneWs<NT>& W, LaneSync& SY, int mode_in DEB_LANE_PARAM, const double* stab = nullptr) {
  const int n = P.n, nh = P.nh, nch = P.nch, nq = P.nq;
  const int nhb = nh - 1;
  // mode_in < 0: the warp is out of work and only attends the CTA's lock-step barriers (same barrier instructions as the
  // working warps) until nobody has work left; it runs the prologue of mode 0 into its own workspace and writes nothing
  const bool idle = mode_in < 0;
  const int mode = idle ? 0 : mode_in;
  const int cosmo = mode / P.nk, kidx = mode - cosmo * P.nk;
  const double k = DEB_LDG(P.kmodes + (P.k_per_cosmo ? (size_t)cosmo * P.nk + kidx : (size_t)kidx));
  const double k2 = k * k;
  DEB_LANES_BEGIN
    if (lane == 0) {
      W.cosmo_ = load_cosmo(P, cosmo);
      if (stab) {       // RHS splines (cs2a, xe, log rho_nu: one contiguous 24 KB block of the table set) staged in shared memory
        Cosmo& cc = W.cosmo_;
        cc.cs2a.x = stab; cc.cs2a.y = stab + P.nth; cc.cs2a.S = stab + 2 * P.nth;
        cc.xe.x = stab + 3 * P.nth; cc.xe.y = stab + 4 * P.nth; cc.xe.S = stab + 5 * P.nth;
        cc.lrn.x = stab + 6 * P.nth; cc.lrn.y = stab + 6 * P.nth + P.nnu; cc.lrn.S = stab + 6 * P.nth + 2 * P.nnu;
      }
    }
    { double* z = W.tails(); for (int i = lane; i < 7 * NT * 32; i += 32) z[i] = 0.0; }
    if (lane < NCHMAX) { W.m2[lane] = 0.0; W.sch[lane] = 0.0; }
    if (lane < 8) W.ka0[lane] = 0.0;
  DEB_LANES_END
  const Cosmo& c = W.cosmo_;
  const NuBins& nb = C.nu;
  const double* tout = P.tau_out + (size_t)cosmo * P.nout;

  DEB_REGS(int, pcol, ); DEB_REGS(double, rscale, ); DEB_REGS(unsigned, pkey, ); DEB_REGS(int, pivl, ); DEB_REGS(double, fmul, );
  DEB_REGS(double, pval, ); DEB_REGS(double, s1, ); DEB_REGS(double, s2, ); DEB_REGS(double, s3, ); DEB_REGS(double, s4, );
  DEB_REGS(int, nanflag, );
  // lane constants: tails
  DEB_REGS(int, tch, ); DEB_REGS(int, tsg, ); DEB_REGS(int, nrow, ); DEB_REGS(int, l0, ); DEB_REGS(int, eb, ); DEB_REGS(int, es, );
  DEB_REGS(int, tL, ); DEB_REGS(int, e2, ); DEB_REGS(int, h2, ); DEB_REGS(int, hasnext, ); DEB_REGS(int, jT, );
  // lane constants: head
  DEB_REGS(int, he, );
  DEB_LANES_BEGIN
    DEB_USE(tch); DEB_USE(tsg); DEB_USE(nrow); DEB_USE(l0); DEB_USE(eb); DEB_USE(es); DEB_USE(tL); DEB_USE(e2); DEB_USE(h2); DEB_USE(hasnext); DEB_USE(he); DEB_USE(jT);
    tch = -1; tsg = 0; nrow = 0; l0 = 3; eb = 0; es = 1; tL = 3; e2 = 0; h2 = 0; hasnext = 0; jT = -1;
    if (lane < LN_NSEG * nch) {
      tsg = lane / nch; tch = lane - tsg * nch;
      tL = C.ch_lmax[tch]; es = C.ch_stride[tch];
      l0 = 3 + tsg * NT;
      nrow = tL - l0 + 1; if (nrow > NT) nrow = NT; if (nrow < 0) nrow = 0;
      eb = C.ch_base[tch] + l0 * es;
      e2 = C.ch_base[tch] + 2 * es; h2 = C.ch_h2[tch];
      hasnext = (tsg + 1 < LN_NSEG) && (tL - (l0 + NT) + 1 > 0);
      jT = (tL >= l0 && tL < l0 + NT) ? tL - l0 : -1;     // the truncation row, if this segment holds it
    }
    he = lane < nh ? C.hidx[lane] : 0;
  DEB_LANES_END

  // ---- prologue ----
  double t1 = DEB_LDG(tout);
  double tmin_out = t1;
  for (int j = 1; j < P.nout; ++j) { double tj = DEB_LDG(tout + j); t1 = fmax(t1, tj); tmin_out = fmin(tmin_out, tj); }
  double t, tnext;
  const size_t item = (size_t)mode;
  if (P.mode == 1) {
    t = DEB_LDG(P.dbg_t0 + mode); tnext = DEB_LDG(P.dbg_t1 + mode); t1 = tnext;
    DEB_LANES_BEGIN
      for (int e = lane; e < n; e += 32) W.y_[e] = DEB_LDG(P.dbg_y0 + (size_t)mode * n + e);
    DEB_LANES_END
  } else {
    const double st0 = start_time(c, k, DEB_LDG(P.lt_small + cosmo));
    double tau_start = 0.99 * fmin(tmin_out, st0);
    if (!(st0 == st0)) tau_start = st0;
    IcScalars ics = ic_scalars(c, tau_start, k);
    DEB_LANES_BEGIN
      for (int e = lane; e < n; e += 32) W.y_[e] = ic_value(P, c, nb, ics, elem_desc(P, e), k);
    DEB_LANES_END
    if (P.mode == 2) {
      DEB_LANES_BEGIN
        if (lane == 0) P.dbg_tau_start[item] = tau_start;
        for (int e = lane; e < n; e += 32) P.dbg_ics[item * n + e] = W.y_[e];
      DEB_LANES_END
      return;
    }
    t = tau_start;
    tnext = t + fmin(t / 4.0, 0.5 * (t1 - t));          // dt0 (perturbations.py:756)
    if (tnext > t1 - 1e-10) tnext = t1;
  }
  // the accepted state of the lane's own tail rows: transposed copy
  DEB_LANES_BEGIN
    DEB_USE(nrow); DEB_USE(eb); DEB_USE(es);
#pragma unroll
    for (int j = 0; j < NT; ++j) if (j < nrow) W.yt[j * 32 + lane] = W.y_[eb + j * es];
  DEB_LANES_END

  double lprev = 0.0, lpprev = 0.0;          // log of the inverse scaled errors of the last two accepted steps
  int nsteps = 0, nacc = 0, save_idx = 0, status = 0;
  if (!(t == t) || !(tnext == tnext) || !(t1 == t1)) status = 2;
  Hints hint; hint.th = -1; hint.nu = -1;

  while (idle || (t < t1 && nsteps < P.max_steps && status == 0)) {
    if (P.mode == 3) {
      if (nsteps >= DEB_LDG(P.rp_n + mode)) break;
      tnext = DEB_LDG(P.rp_tnext + (size_t)mode * P.rp_stride + nsteps);
      if (fabs(tnext - t1) <= 1e-12 * fabs(t1)) tnext = t1;
    }
    DEB_REGS(double, kt, [7][NT]);      // k_1..k_7 of the lane's tail rows
    DEB_REGS(double, ut, [NT]);         // stage state of the tail rows
    DEB_REGS(double, bt, [NT]);         // right-hand side -> solution of the tail rows
    DEB_REGS(double, ufirst, ); DEB_REGS(double, ulast, ); DEB_REGS(double, cfirst, ); DEB_REGS(double, clast, );
    DEB_REGS(double, cin, );
    DEB_REGS(double, kcl, ); DEB_REGS(double, kpl, );
    const double dt = tnext - t;
    const double invdt = DEB_RCP(dt);
    const double idg = DEB_RCP(dt * RD_GAMMA);
    const double invt0 = DEB_RCP(t);
    const double gdt = dt * RD_GAMMA;

    // ================= Jacobian pieces at (t, y): stage-1 right-hand side, d f/d a, factorisation =================
    LN_BAR(SY, !idle);
    if (idle && SY.cnt == 0) return;
    double x0piv = 1.0, r0 = 0.0;
    double ci_hh = 0.0, ci_he = 0.0, ci_eh = 0.0, ci_ee = 0.0, jq_h = 0.0, jq_e = 0.0;     // inverse capacitance; a h' row applied to qh, qe
    if (!idle) {
    {
      Bg<Dual> bd;
      compute_bg<Dual>(c, nb, nq, mk(W.y_[0], 1.0), hint, W.ic_, bd);
      DEB_LANES_BEGIN
        if (lane < nch) chain_coeffs_lane<Dual>(c, nb, bd, k, lane, W.y_, P.iq0, W.kc_, W.kap_, W.nur_, W.nup_, W.sl_);
        if (lane == 0) fill_slots<Dual>(c, bd, k, W.sl_);
      DEB_LANES_END
      Metric<Dual> md;
      compute_metric<Dual>(P, c, nb, bd, W.y_, k, W.nur_, W.nup_, md);
      {
        const Dual f0 = bd.H * bd.a;
        r0 = f0.v; x0piv = idg - f0.d;
      }
      const double d1t = (dt * RD_D1) * invt0 * invt0;
      // first boundary exchange: the neighbours of a segment's end rows (state y)
      LN_BEGIN
        DEB_USE(ufirst); DEB_USE(ulast);
        ufirst = W.yt[lane]; ulast = W.yt[(NT - 1) * 32 + lane];
      LN_END
      DEB_LANES_BEGIN
        DEB_USE(tch); DEB_USE(tsg); DEB_USE(tL); DEB_USE(e2); DEB_USE(hasnext); DEB_USE(bt); DEB_USE(jT);
        DEB_USE(ufirst); DEB_USE(ulast);
        if (lane < nh) {
          Dual f = head_row<Dual>(C, W.sl_, W.y_, lane, md);
          W.rh[lane] = f.v; W.jah[lane] = f.d;
        }
        const double um_in = LN_SHFL(ulast, lane - nch), up_in = LN_SHFL(ufirst, lane + nch);
        {
          // rows l >= 3 (perturbations.py:300-303, :307-312, :323-327, :346-360): coefficient rows from the CTA table,
          // (0, 0) past lmax, so no row needs a branch; the truncation row adds (lmax+1)/tau damping and its d/dt
          const int tc = tch >= 0 ? tch : 0;
          const double kcv = W.kc_[tc], kcd = W.kc_[NCHMAX + tc], kpv = W.kap_[tc], kpd = W.kap_[NCHMAX + tc];
          const double um0 = tsg == 0 ? W.y_[e2] : um_in, upN = hasnext ? up_in : 0.0;
          const double trv = (double)(tL + 1);
          double yv[NT];
#pragma unroll
          for (int j = 0; j < NT; ++j) yv[j] = W.yt[j * 32 + lane];
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            const double um = (j == 0) ? um0 : yv[j > 0 ? j - 1 : 0];
            const double up = (j == NT - 1) ? upN : yv[j < NT - 1 ? j + 1 : j];
            const double lin = T.cl[j * 32 + lane] * um - T.ch[j * 32 + lane] * up;
            double f = kcv * lin - kpv * yv[j];
            if (j == jT) f += (d1t - invt0) * trv * yv[j];
            bt[j] = f;
            W.jat[j * 32 + lane] = kcd * lin - kpd * yv[j];
          }
        }
      DEB_LANES_END

      // head-column gradients of h', eta' and of row 1 (value parts only)
      const double H = bd.H.v, a = bd.a.v;
      DEB_LANES_BEGIN
        if (lane < nh) {
          int ty = C.htype[lane], bin = C.hbin[lane];
          double wr = 0.0, wp = 0.0, wt = 0.0, extra = 0.0;
          switch (ty) {
            case R_ETA: extra = 2.0 * k2 / H; break;
            case R_DC: wr = bd.gc.v; break;
            case R_TC: wt = bd.gc.v; break;
            case R_DB: wr = bd.gb.v; break;
            case R_TB: wt = bd.gb.v; break;
            case R_F0: wr = bd.gg.v; wp = bd.gg.v / 3.0; break;
            case R_F1: wt = 4.0 / 3.0 * bd.gg.v; break;
            case R_N0: wr = bd.gr.v; wp = bd.gr.v / 3.0; break;
            case R_N1: wt = 4.0 / 3.0 * bd.gr.v; break;
            case R_P0: { const double vb = W.kc_[3 + bin] / k; wr = bd.gnu.v * nb.w[bin] / vb; wp = bd.gnu.v * nb.w[bin] * vb / 3.0; } break;
            case R_P1: wt = bd.gnu.v * k * nb.w[bin]; break;
            case R_DQ: wr = bd.gq.v; wp = c.cs2de * bd.gq.v; break;
            case R_TQ: wt = bd.wq1.v * bd.gq.v; wp = (c.cs2de - bd.ca2.v) * 3.0 * H * wt / k2; break;
            default: break;
          }
          W.gh_[lane] = wr / H + extra;
          W.ge_[lane] = 0.5 * wt / k2;
          W.j1_[lane] = -(wr + 3.0 * wp) * a;
        }
      DEB_LANES_END

      // ---- tails: pivot recurrence l = L .. 3, one lane per chain (sequential: a continued fraction); the factors
      //      go to the transposed arrays of the (segment, chain) lane that owns the row ----
      DEB_LANES_BEGIN
        if (lane < nch) {
          const int L = C.ch_lmax[lane];
          const double kc = W.kc_[lane], kp = W.kap_[lane];
          double ie = DEB_RCP(idg + kp + (double)(L + 1) * invt0);
          int jr = (L - 3) % NT, sgr = (L - 3) / NT;       // (row in segment, segment) of the chain's row l
          int pos = jr * 32 + sgr * nch + lane;
          W.iet[pos] = ie;
          W.gt[pos] = kc * ie;                             // -W_{L,L-1}/e_L
          W.mt[pos] = 0.0;
          double lower_next = -kc;                         // W_{l+1,l}
          for (int l = L - 1; l >= 3; --l) {
            const double mm = (kc * C.ch[l]) * ie;         // W_{l,l+1} / e_{l+1}
            if (--jr < 0) { jr = NT - 1; pos += (NT - 1) * 32 - nch; } else pos -= 32;
            W.mt[pos] = mm;
            ie = DEB_RCP(idg + kp - mm * lower_next);
            W.iet[pos] = ie;
            lower_next = -kc * C.cl[l];
            W.gt[pos] = -lower_next * ie;
          }
          {
            const double mm = (kc * C.ch[2]) * ie;
            W.m2[lane] = mm;
            W.sch[lane] = -mm * lower_next;                // Schur increment of the head diagonal (l = 2 row)
          }
        }
      DEB_LANES_END
      // cumulative multiplier products of every segment: backward  P_j = prod_{jj >= j} (-m_jj), forward Q_j = prod_{jj <= j} g_jj
      DEB_LANES_BEGIN
        double p = 1.0;
#pragma unroll
        for (int j = NT - 1; j >= 0; --j) { p *= -W.mt[j * 32 + lane]; W.pct[j * 32 + lane] = p; }
        double q = 1.0;
#pragma unroll
        for (int j = 0; j < NT; ++j) { q *= W.gt[j * 32 + lane]; W.qct[j * 32 + lane] = q; }
      DEB_LANES_END

      // ---- head:  W_h = D - chv gh^T - cev ge^T  (+ the a h' row); D block diagonal -> block inverses + Woodbury ----
      DEB_LANES_BEGIN
        DEB_USE(pcol); DEB_USE(rscale);
        pcol = -1; rscale = 1.0;
        if (lane < nhb) {
          const int ty = C.htype[lane], lo = C.blo[lane];
          double* row = lrow(W, lane, lo);
#pragma unroll
          for (int i = 0; i < 8; ++i) row[lo + i] = 0.0;
          double diag = idg;
          if (ty == R_F2 || ty == R_G2 || ty == R_N2 || ty == R_P2) {
            const int chain = ty == R_F2 ? 0 : (ty == R_G2 ? 1 : (ty == R_N2 ? 2 : 3 + C.hbin[lane]));
            diag += W.sch[chain];
          }
          row[lane] = diag;
#pragma unroll
          for (int q = 0; q < HOP_NT; ++q) {
            const int m = C.hop_meta[q][lane], hc = (m >> 20) - 1;
            if (hc >= 0) row[hc] -= C.hop_c[q][lane] * W.sl_[m & 0xff];
          }
        }
      DEB_LANES_END
    }
    // (register-resident Gauss-Jordan of deb_core.cuh on this workspace's row storage)
#undef DEB_GJ_ROWP
#undef DEB_GJ_PERM
#define DEB_GJ_ROWP(r) (lrow(W, (r), C.blo[(r)]) + C.blo[(r)])
#define DEB_GJ_PERM(j) W.perm_[(j)]
    DEB_GJ_ALL(true)
#undef DEB_GJ_ROWP
#undef DEB_GJ_PERM
#define DEB_GJ_ROWP(r) (hrow(W, (r), C.blo[(r)]) + C.blo[(r)])
#define DEB_GJ_PERM(j) W.perm()[(j)]
    DEB_LANES_BEGIN
      W.xb_[lane] = 0.0;
      if (lane < 8) W.xb_[NHMAX + lane] = 0.0;
    DEB_LANES_END
    DEB_LANES_BEGIN
      DEB_USE(pcol); DEB_USE(s1); DEB_USE(s2); DEB_USE(s3); DEB_USE(s4);
      s1 = s2 = s3 = s4 = 0.0;
      if (lane < nhb) {
        const int lo = C.blo[lane], hi = C.bhi[lane];
        const double* row = lrow(W, lane, lo);
        double ah = 0.0, ae = 0.0;
        for (int cc = lo; cc < hi; ++cc) {
          const int pr = W.perm_[cc];
          ah += row[cc] * (C.hop_chc[pr] * W.sl_[C.hop_chs[pr]]);
          ae += row[cc] * C.hop_cec[pr];
        }
        W.qh_[pcol] = ah; W.qe_[pcol] = ae;
        const double ghc = W.gh_[pcol], gec = W.ge_[pcol], j1c = W.j1_[pcol];
        s1 = ghc * ah; s2 = ghc * ae; s3 = gec * ah; s4 = gec * ae;
        W.xb_[lane] = j1c;
      }
    DEB_LANES_END
    {
      const double c_hh = 1.0 - DEB_WARP_SUM(s1), c_he = -DEB_WARP_SUM(s2), c_eh = -DEB_WARP_SUM(s3), c_ee = 1.0 - DEB_WARP_SUM(s4);
      const double idet = DEB_RCP(c_hh * c_ee - c_he * c_eh);
      ci_hh = c_ee * idet; ci_he = -c_he * idet; ci_eh = -c_eh * idet; ci_ee = c_hh * idet;
      DEB_LANES_BEGIN
        DEB_USE(pcol); DEB_USE(s1); DEB_USE(s2);
        s1 = s2 = 0.0;
        if (lane < nhb) { const double j1c = W.xb_[lane]; s1 = j1c * W.qh_[pcol]; s2 = j1c * W.qe_[pcol]; }
      DEB_LANES_END
      jq_h = DEB_WARP_SUM(s1); jq_e = DEB_WARP_SUM(s2);
    }

    }     // !idle
    // ================= 8 stages =================
    double errnorm2 = 0.0;
    double ua = W.y_[0];            // scale factor of the stage state (element 0), uniform over the warp
    double x0 = 0.0;
    double ts = t, invts = invt0;
#pragma unroll 1
    for (int st = 1; st <= 8; ++st) {
      if (SY.every == 1 || (st & (SY.every - 1)) == 1) LN_BAR(SY, !idle);
      if (idle) continue;
      if (st > 1) {
        ts = st == 2 ? t + RD_CT2 * dt : st == 3 ? t + RD_CT3 * dt : st == 4 ? t + RD_CT4 * dt : st == 5 ? t + RD_CT5 * dt : t + dt;
        const double dtd = st == 2 ? dt * RD_D2 : st == 3 ? dt * RD_D3 : st == 4 ? dt * RD_D4 : st == 5 ? dt * RD_D5 : 0.0;
        invts = DEB_RCP(ts);
        const double dtt = dtd * invt0 * invt0;
        // ---- stage state: scale factor (uniform), head row and own tail rows ----
        double ca0;                                      // sum_j C_ij k_j[0]
        {
          const double* a0 = W.ka0;
          switch (st) {
            case 2: ua = W.y_[0] + RD_A21 * a0[0]; ca0 = RD_C21 * a0[0]; break;
            case 3: ua = W.y_[0] + RD_A31 * a0[0] + RD_A32 * a0[1]; ca0 = RD_C31 * a0[0] + RD_C32 * a0[1]; break;
            case 4: ua = W.y_[0] + RD_A41 * a0[0] + RD_A42 * a0[1] + RD_A43 * a0[2]; ca0 = RD_C41 * a0[0] + RD_C42 * a0[1] + RD_C43 * a0[2]; break;
            case 5: ua = W.y_[0] + RD_A51 * a0[0] + RD_A52 * a0[1] + RD_A53 * a0[2] + RD_A54 * a0[3];
                    ca0 = RD_C51 * a0[0] + RD_C52 * a0[1] + RD_C53 * a0[2] + RD_C54 * a0[3]; break;
            case 6: ua = W.y_[0] + RD_A61 * a0[0] + RD_A62 * a0[1] + RD_A63 * a0[2] + RD_A64 * a0[3] + RD_A65 * a0[4];
                    ca0 = RD_C61 * a0[0] + RD_C62 * a0[1] + RD_C63 * a0[2] + RD_C64 * a0[3] + RD_C65 * a0[4]; break;
            case 7: ua = ua + a0[5];
                    ca0 = RD_C71 * a0[0] + RD_C72 * a0[1] + RD_C73 * a0[2] + RD_C74 * a0[3] + RD_C75 * a0[4] + RD_C76 * a0[5]; break;
            default: ua = ua + a0[6];
                    ca0 = RD_C81 * a0[0] + RD_C82 * a0[1] + RD_C83 * a0[2] + RD_C84 * a0[3] + RD_C85 * a0[4] + RD_C86 * a0[5] + RD_C87 * a0[6]; break;
          }
        }
        // own rows: u = y + sum a_ij k_j (stages 7, 8: u + k), c-combination -> bt / rh
        DEB_LANES_BEGIN
          DEB_USE(kt); DEB_USE(ut); DEB_USE(bt); DEB_USE(nrow); DEB_USE(he); DEB_USE(tsg); DEB_USE(eb);
          const double* kh_ = W.khs + lane;
#define kh(i) kh_[(i) * 32]
          DEB_USE(ufirst); DEB_USE(ulast); DEB_USE(cfirst);
          double yv[NT];
#pragma unroll
          for (int j = 0; j < NT; ++j) yv[j] = W.yt[j * 32 + lane];
          const double yh = lane < nh ? W.y_[he] : 0.0;
          double uh, chh;
#define LN_T(expr_u, expr_c) _Pragma("unroll") for (int j = 0; j < NT; ++j) { ut[j] = (expr_u); bt[j] = invdt * (expr_c); }
          switch (st) {
            case 2: LN_T(yv[j] + RD_A21 * kt[0][j], RD_C21 * kt[0][j])
                    uh = yh + RD_A21 * kh(0); chh = RD_C21 * kh(0); break;
            case 3: LN_T(yv[j] + RD_A31 * kt[0][j] + RD_A32 * kt[1][j], RD_C31 * kt[0][j] + RD_C32 * kt[1][j])
                    uh = yh + RD_A31 * kh(0) + RD_A32 * kh(1); chh = RD_C31 * kh(0) + RD_C32 * kh(1); break;
            case 4: LN_T(yv[j] + RD_A41 * kt[0][j] + RD_A42 * kt[1][j] + RD_A43 * kt[2][j], RD_C41 * kt[0][j] + RD_C42 * kt[1][j] + RD_C43 * kt[2][j])
                    uh = yh + RD_A41 * kh(0) + RD_A42 * kh(1) + RD_A43 * kh(2); chh = RD_C41 * kh(0) + RD_C42 * kh(1) + RD_C43 * kh(2); break;
            case 5: LN_T(yv[j] + RD_A51 * kt[0][j] + RD_A52 * kt[1][j] + RD_A53 * kt[2][j] + RD_A54 * kt[3][j],
                         RD_C51 * kt[0][j] + RD_C52 * kt[1][j] + RD_C53 * kt[2][j] + RD_C54 * kt[3][j])
                    uh = yh + RD_A51 * kh(0) + RD_A52 * kh(1) + RD_A53 * kh(2) + RD_A54 * kh(3);
                    chh = RD_C51 * kh(0) + RD_C52 * kh(1) + RD_C53 * kh(2) + RD_C54 * kh(3); break;
            case 6: LN_T(yv[j] + RD_A61 * kt[0][j] + RD_A62 * kt[1][j] + RD_A63 * kt[2][j] + RD_A64 * kt[3][j] + RD_A65 * kt[4][j],
                         RD_C61 * kt[0][j] + RD_C62 * kt[1][j] + RD_C63 * kt[2][j] + RD_C64 * kt[3][j] + RD_C65 * kt[4][j])
                    uh = yh + RD_A61 * kh(0) + RD_A62 * kh(1) + RD_A63 * kh(2) + RD_A64 * kh(3) + RD_A65 * kh(4);
                    chh = RD_C61 * kh(0) + RD_C62 * kh(1) + RD_C63 * kh(2) + RD_C64 * kh(3) + RD_C65 * kh(4); break;
            case 7: LN_T(ut[j] + kt[5][j], RD_C71 * kt[0][j] + RD_C72 * kt[1][j] + RD_C73 * kt[2][j] + RD_C74 * kt[3][j] + RD_C75 * kt[4][j] + RD_C76 * kt[5][j])
                    uh = (lane < nh ? W.u_[he] : 0.0) + kh(5);
                    chh = RD_C71 * kh(0) + RD_C72 * kh(1) + RD_C73 * kh(2) + RD_C74 * kh(3) + RD_C75 * kh(4) + RD_C76 * kh(5); break;
            default: LN_T(ut[j] + kt[6][j], RD_C81 * kt[0][j] + RD_C82 * kt[1][j] + RD_C83 * kt[2][j] + RD_C84 * kt[3][j] + RD_C85 * kt[4][j] + RD_C86 * kt[5][j] + RD_C87 * kt[6][j])
                    uh = (lane < nh ? W.u_[he] : 0.0) + kh(6);
                    chh = RD_C81 * kh(0) + RD_C82 * kh(1) + RD_C83 * kh(2) + RD_C84 * kh(3) + RD_C85 * kh(4) + RD_C86 * kh(5) + RD_C87 * kh(6); break;
          }
#undef LN_T
#undef kh
          cfirst = chh;                                   // parked: the head row is formed after the metric sources
          ufirst = ut[0]; ulast = ut[NT - 1];
          if (lane < nh) W.u_[he] = uh;
          if (lane == 0) W.u_[0] = ua;
          if (nrow > 0 && tsg == 0) W.u_[eb] = ut[0];     // the l = 3 entries feed the l = 2 head rows
        DEB_LANES_END
        // ---- background at the stage's scale factor, chain coefficients, metric sources ----
        Bg<double> b;
        compute_bg<double>(c, nb, nq, ua, hint, W.ic_, b);
        DEB_LANES_BEGIN
          if (lane < nch) chain_coeffs_lane<double>(c, nb, b, k, lane, W.u_, P.iq0, W.kc_, W.kap_, W.nur_, W.nup_, W.sl_);
          if (lane == 0) fill_slots<double>(c, b, k, W.sl_);
        DEB_LANES_END
        Metric<double> mt;
        compute_metric<double>(P, c, nb, b, W.u_, k, W.nur_, W.nup_, mt);
        // row 0 is closed: x0 first, the d f/d a column then rides on every other row
        x0 = (b.H * b.a + invdt * ca0) / x0piv;
        // ---- rows: f(ts, u) + dt d_i dT + sum_j C_ij/dt k_j + (d f/d a) x0 ----
        DEB_LANES_BEGIN
          DEB_USE(tch); DEB_USE(tsg); DEB_USE(tL); DEB_USE(e2); DEB_USE(hasnext); DEB_USE(bt); DEB_USE(ut); DEB_USE(jT);
          DEB_USE(ufirst); DEB_USE(ulast); DEB_USE(cfirst);
          if (lane < nh) W.rh[lane] = head_row<double>(C, W.sl_, W.u_, lane, mt) + invdt * cfirst + W.jah[lane] * x0;
          const double um_in = LN_SHFL(ulast, lane - nch), up_in = LN_SHFL(ufirst, lane + nch);
          {
            const int tc = tch >= 0 ? tch : 0;
            const double kcv = W.kc_[tc], kpv = W.kap_[tc];
            const double um0 = tsg == 0 ? W.u_[e2] : um_in, upN = hasnext ? up_in : 0.0;
            const double trv = (double)(tL + 1);
#pragma unroll
            for (int j = 0; j < NT; ++j) {
              const double um = (j == 0) ? um0 : ut[j > 0 ? j - 1 : 0];
              const double up = (j == NT - 1) ? upN : ut[j < NT - 1 ? j + 1 : j];
              const double lin = T.cl[j * 32 + lane] * um - T.ch[j * 32 + lane] * up;
              double f = kcv * lin - kpv * ut[j];
              if (j == jT) f += trv * (dtt * W.yt[j * 32 + lane] - invts * ut[j]);
              bt[j] = bt[j] + f + W.jat[j * 32 + lane] * x0;
            }
          }
        DEB_LANES_END
      } else {
        // stage 1: the right-hand side was formed with the Jacobian pieces; fold the a column
        x0 = r0 / x0piv;
        DEB_LANES_BEGIN
          DEB_USE(bt); DEB_USE(nrow);
          if (lane < nh) W.rh[lane] += W.jah[lane] * x0;
#pragma unroll
          for (int j = 0; j < NT; ++j) if (j < nrow) bt[j] += W.jat[j * 32 + lane] * x0;
        DEB_LANES_END
      }

      // ---- solve W x = r: backward sweep (segments from a zero carry, carry chain, fix-up) ----
      LN_BEGIN
        DEB_USE(bt); DEB_USE(cfirst); DEB_USE(cin);
        double loc = 0.0;
#pragma unroll
        for (int j = NT - 1; j >= 0; --j) { loc = bt[j] - W.mt[j * 32 + lane] * loc; bt[j] = loc; }
        cfirst = loc; cin = 0.0;
      LN_END
#pragma unroll
      for (int rd = 1; rd < LN_NSEG; ++rd) {
        LN_BEGIN
          DEB_USE(bt); DEB_USE(cfirst); DEB_USE(cin); DEB_USE(tsg); DEB_USE(tch); DEB_USE(hasnext);
          const double v = LN_SHFL(cfirst, lane + nch);
          if (tch >= 0 && tsg == LN_NSEG - 1 - rd && hasnext) { cin = v; cfirst = bt[0] + W.pct[lane] * v; }
        LN_END
      }
      DEB_LANES_BEGIN
        DEB_USE(bt); DEB_USE(cin); DEB_USE(cfirst); DEB_USE(tch); DEB_USE(tsg); DEB_USE(h2); DEB_USE(nrow);
#pragma unroll
        for (int j = 0; j < NT; ++j) bt[j] = bt[j] + W.pct[j * 32 + lane] * cin;
        if (tch >= 0 && tsg == 0 && nrow > 0) W.rh[h2] -= W.m2[tch] * bt[0];      // b'_2 = b_2 - m_2 b'_3
      DEB_LANES_END
      // head: p = D^-1 b (block inverses), the rank-2 Woodbury correction and the a h' row
      DEB_LANES_BEGIN
        W.xb_[lane] = lane < nhb ? W.rh[W.perm_[lane]] : 0.0;
      DEB_LANES_END
      DEB_LANES_BEGIN
        DEB_USE(pcol); DEB_USE(s1); DEB_USE(s2); DEB_USE(s3); DEB_USE(pval);
        s1 = s2 = s3 = 0.0; pval = 0.0;
        if (lane < nhb) {
          const int lo = C.blo[lane];
          const double* rb = lrow(W, lane, lo) + lo;
          const double* xb = W.xb_ + lo;
          double ra[8], xa[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) { ra[i] = rb[i]; xa[i] = xb[i]; }
          const double acc = ((ra[0] * xa[0] + ra[1] * xa[1]) + (ra[2] * xa[2] + ra[3] * xa[3]))
                           + ((ra[4] * xa[4] + ra[5] * xa[5]) + (ra[6] * xa[6] + ra[7] * xa[7]));
          pval = acc;
          s1 = W.gh_[pcol] * acc; s2 = W.ge_[pcol] * acc; s3 = W.j1_[pcol] * acc;
        }
      DEB_LANES_END
      {
        const double th = DEB_WARP_SUM(s1), te = DEB_WARP_SUM(s2), ta = DEB_WARP_SUM(s3);
        const double sh = ci_hh * th + ci_he * te, se = ci_eh * th + ci_ee * te;
        DEB_LANES_BEGIN
          DEB_USE(pcol); DEB_USE(pval);
          if (lane < nhb) W.rh[pcol] = pval + W.qh_[pcol] * sh + W.qe_[pcol] * se;
          else if (lane == nhb) W.rh[nhb] = (W.rh[nhb] + ta + jq_h * sh + jq_e * se) * gdt;
        DEB_LANES_END
      }
      // forward sweep: x_l = b'_l / e_l + g_l x_{l-1}
      LN_BEGIN
        DEB_USE(bt); DEB_USE(clast); DEB_USE(cin); DEB_USE(tch); DEB_USE(tsg); DEB_USE(h2); DEB_USE(nrow);
        double loc = 0.0;
#pragma unroll
        for (int j = 0; j < NT; ++j) { loc = bt[j] * W.iet[j * 32 + lane] + W.gt[j * 32 + lane] * loc; bt[j] = loc; }
        cin = (tch >= 0 && tsg == 0 && nrow > 0) ? W.rh[h2] : 0.0;                 // x_2 of the chain (head solution)
        clast = loc + W.qct[(NT - 1) * 32 + lane] * cin;
      LN_END
#pragma unroll
      for (int rd = 1; rd < LN_NSEG; ++rd) {
        LN_BEGIN
          DEB_USE(bt); DEB_USE(clast); DEB_USE(cin); DEB_USE(tsg); DEB_USE(tch); DEB_USE(nrow);
          const double v = LN_SHFL(clast, lane - nch);
          if (tch >= 0 && tsg == rd && nrow > 0) { cin = v; clast = bt[NT - 1] + W.qct[(NT - 1) * 32 + lane] * v; }
        LN_END
      }
      DEB_LANES_BEGIN
        DEB_USE(bt); DEB_USE(cin); DEB_USE(kt);
#pragma unroll
        for (int j = 0; j < NT; ++j) bt[j] = bt[j] + W.qct[j * 32 + lane] * cin;
        const double xh = lane < nh ? W.rh[lane] : 0.0;
        // ---- keep k_st ----
        switch (st) {
#define LN_K(i) _Pragma("unroll") for (int j = 0; j < NT; ++j) kt[i][j] = bt[j]; W.khs[(i) * 32 + lane] = xh; if (lane == 0) W.ka0[i] = x0;
          case 1: LN_K(0) break;
          case 2: LN_K(1) break;
          case 3: LN_K(2) break;
          case 4: LN_K(3) break;
          case 5: LN_K(4) break;
          case 6: LN_K(5) break;
          case 7: LN_K(6) break;
          default: break;
#undef LN_K
        }
      DEB_LANES_END
    }
    if (idle) continue;
    // y1 = u + k8 (candidate; error estimate = k8): tails stay in registers, head and a go to u_
    DEB_LANES_BEGIN
      DEB_USE(nanflag); DEB_USE(ut); DEB_USE(bt); DEB_USE(he);
      nanflag = 0;
#pragma unroll
      for (int j = 0; j < NT; ++j) { const double v = ut[j] + bt[j]; ut[j] = v; nanflag |= (v != v); }
      if (lane < nh) { const double v = W.u_[he] + W.rh[lane]; W.u_[he] = v; nanflag |= (v != v); }
      if (lane == 0) { const double v = ua + x0; W.u_[0] = v; nanflag |= (v != v); }
    DEB_LANES_END

    if (P.mode == 1) {
      DEB_LANES_BEGIN
        DEB_USE(ut); DEB_USE(bt); DEB_USE(nrow); DEB_USE(eb); DEB_USE(es); DEB_USE(he);
#pragma unroll
        for (int j = 0; j < NT; ++j) if (j < nrow) { P.dbg_y1[item * n + eb + j * es] = ut[j]; P.dbg_err[item * n + eb + j * es] = bt[j]; }
        if (lane < nh) { P.dbg_y1[item * n + he] = W.u_[he]; P.dbg_err[item * n + he] = W.rh[lane]; }
        if (lane == 0) { P.dbg_y1[item * n] = W.u_[0]; P.dbg_err[item * n] = x0; }
      DEB_LANES_END
      return;
    }

    // ================= error norm, PID controller (diffrax semantics) =================
    {
      const bool anynan = DEB_ANY(nanflag) != 0;
      const double ik2 = DEB_RCP(k2);
#define LN_ERRC(e, ev_, w) { double y0v = W.y_[e], y1v = anynan ? y0v : W.u_[e], ev = (ev_); if (ev != ev) ev = INFINITY; \
        double sc = ev / (P.atol + fmax(fabs(y0v), fabs(y1v)) * P.rtol) * (w); errnorm2 += sc * sc; }
      LN_ERRC(0, x0, 1.0) LN_ERRC(2, W.rh[0], k2) LN_ERRC(3, W.rh[1], 1.0) LN_ERRC(5, W.rh[3], 1.0) LN_ERRC(6, W.rh[4], ik2) LN_ERRC(7, W.rh[5], 1.0)
#undef LN_ERRC
      DEB_SYNC();
    }
    const double E = sqrt(errnorm2 / 6.0);
    const bool keep = (P.mode == 3) ? (DEB_LDG(P.rp_keep + (size_t)mode * P.rp_stride + nsteps) != 0) : (E < 1.0);
    double inv = 1.0 / E;
    // inv^c1 * inv_prev^c2 * inv_pprev^c3 (PID law) through ONE exp: the logs of the stored inverse errors are carried
    // along (E = 0 or inf keep their limits through pow)
    double fpid;
    const double linv = (inv > 0.0 && !isinf(inv)) ? DEB_LOG(inv) : 0.0;
    if (inv > 0.0 && !isinf(inv)) fpid = DEB_EXP(P.c1 * linv + P.c2 * lprev + P.c3 * lpprev);
    else fpid = (P.c1 != 0.0 ? DEB_POW(inv, P.c1) : 1.0) * DEB_EXP(P.c2 * lprev + P.c3 * lpprev);
    double fac = fmin(fmax(P.safety * fpid, keep ? 1.0 : P.factormin), P.factormax);
    if (!(fac == fac)) fac = NAN;
    const double dtn = dt * fac;
    const double lcur = (inv == 0.0 || isinf(inv)) ? 0.0 : linv;      // log of what diffrax stores (1 for E = 0 or inf)
    ++nsteps;
    if (keep) {
      ++nacc;
      // SaveAt(ts): linear interpolation inside the accepted step
      while (save_idx < P.nout && DEB_LDG(tout + save_idx) <= tnext) {
        const double tt = DEB_LDG(tout + save_idx);
        const double coeff = (tnext == t) ? 0.0 : (tt - t) / (tnext - t);
        const size_t obase = ((size_t)mode * P.nout + save_idx);
        if (P.return_full) {
          DEB_LANES_BEGIN
            DEB_USE(ut); DEB_USE(nrow); DEB_USE(eb); DEB_USE(es); DEB_USE(he);
#pragma unroll
            for (int j = 0; j < NT; ++j) if (j < nrow) { const double y0v = W.yt[j * 32 + lane]; P.y_out[obase * n + eb + j * es] = y0v + coeff * (ut[j] - y0v); }
            if (lane < nh) P.y_out[obase * n + he] = W.y_[he] + coeff * (W.u_[he] - W.y_[he]);
            if (lane == 0) P.y_out[obase * n] = W.y_[0] + coeff * (W.u_[0] - W.y_[0]);
          DEB_LANES_END
        } else {
          DEB_LANES_BEGIN
            DEB_USE(he);
            if (lane < nh) W.o_()[he] = W.y_[he] + coeff * (W.u_[he] - W.y_[he]);
            if (lane == 0) W.o_()[0] = W.y_[0] + coeff * (W.u_[0] - W.y_[0]);
          DEB_LANES_END
          DEB_LANE0_BEGIN
            double o20[20];
            convert_outputs(P, c, nb, W.o_(), k, o20);
            const bool has_pk = P.pk_out && P.power_idx >= 0;
            double pkv = 0.0;
            if (has_pk) { const double yv = o20[P.power_idx]; pkv = 2.0 * 9.869604401089358 * c.As * DEB_POW(k / c.kp, c.ns - 1.0) * DEB_POW(k, -3.0) * yv * yv; }
            store_fields(P, mode, save_idx, o20, has_pk, pkv);
          DEB_LANE0_END
        }
        ++save_idx;
      }
      DEB_LANES_BEGIN
        DEB_USE(ut); DEB_USE(nrow); DEB_USE(tsg); DEB_USE(eb); DEB_USE(he);
#pragma unroll
        for (int j = 0; j < NT; ++j) if (j < nrow) W.yt[j * 32 + lane] = ut[j];
        if (nrow > 0 && tsg == 0) W.y_[eb] = ut[0];
        if (lane < nh) W.y_[he] = W.u_[he];
        if (lane == 0) W.y_[0] = W.u_[0];
      DEB_LANES_END
      lpprev = lprev; lprev = lcur;
      t = fmin(tnext, t1);
    }
    double tn = t + dtn;
    if (tn > t1 - 1e-10) tn = keep ? t1 : t + 0.5 * (t1 - t);
    tnext = tn;
    if (!(tnext == tnext) || isinf(tnext)) status = 2;
  }
  if (status == 0 && t < t1) status = 1;
  DEB_LANE0_BEGIN
    store_status(P, mode, status, nsteps, nacc);
  DEB_LANE0_END
}

}  // namespace deb
