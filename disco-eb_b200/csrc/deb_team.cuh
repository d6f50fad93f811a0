// deb_team.cuh -- CTA-per-mode ("team") variant of the per-mode integrator for launches that cannot fill the GPU.
//
// integrate_mode (deb_core.cuh) runs one mode on ONE warp: with 512 modes on 148 SMs every warp has a scheduler to
// itself and the time of the launch is the latency of its slowest mode, 8 stages x (serial solve + element-wise
// work) per step.  62 % of a step's instructions are element-parallel (stage combinations, hierarchy rows l >= 3,
// the d f/d a column, k_i stores).  Here a mode gets a team of TEAM warps (one CTA):
//   * every thread owns the elements e = tid + 32 TEAM j of the stage vectors k_1..k_7 (registers) and evaluates
//     the tail rows tt = tid + 32 TEAM j;
//   * warp 0 runs what is serial: metric sources, the head rows, the bordered solve (tail sweeps, block inverses,
//     Woodbury correction); the d f/d a column is folded into the sweeps and the head gather;
//   * warp 1 is also the "helper": the scale factor of the NEXT stage is known as soon as x_0 = r_0 / W_00 is, so
//     it evaluates the background scalars of stage i+1 while warp 0 solves stage i; at the start of a step it
//     factors the hierarchy tails while warp 0 builds the head;
//   * three CTA barriers per stage separate the phases.
// Same arithmetic, expression by expression, as the two-warp variant of integrate_mode (reference map: see
// include/discoeb_b200.h and the comments in deb_core.cuh).  Debug modes 1-3 are supported, tangents and shared-step
// batches are not (they keep the one-warp kernels).
#pragma once

namespace deb {

struct TeamBg {                 // what warp 1 posts for one stage (double-buffered by stage parity)
  double a_req;
  double bgs[10];               // H, gc, gb, gg, gr, gnu, gq, wq1, ca2, wq at a_req
  double vv[2 * NQMAX];         // v_i, 1/v_i
  double kc[NCHMAX], kap[NCHMAX];
  double sl[2 * NSLOT];         // operator slots at a_req (value part used)
};
constexpr int TEAM_NSEG = 4;    // segments per hierarchy tail in the sweeps (TEAM_NSEG * NCHMAX = 32 lanes)
constexpr int TEAM_KF = TEAM_NSEG * NCHMAX + 8;
// what the Jacobian evaluation needs of the background at a = y_0, as (value, d/da): posted by warp 1 one step ahead
// (the scale factor of the accepted state, u_0 + x_0 of stage 8, is known before stage 8 is solved)
struct TeamJac {
  Bg<Dual> bd;
  double kc[2 * NCHMAX], kap[2 * NCHMAX];
  double sl[2 * NSLOT];
  double vvd[4 * NQMAX];        // v_i (value, d/da), 1/v_i (value, d/da)
};
struct TeamBox {
  TeamBg bg[2];
  TeamJac jac[2];               // [jcur]: at the state of the current step; [jcur ^ 1]: at the candidate of this step
  double k, x0, x0piv;          // x0 = r_0 / W_00 of the current stage (element 0 of k_i), W_00
  double ic2[ICACHE];           // warp 1's own spline interval cache
  double ka0[8];                // element 0 of k_1..k_7
  double kf[TEAM_KF];           // forward-sweep carry of (segment, chain) = slot; slot 32 stays 0 (head elements)
  double t3[NCHMAX];            // sweep warp -> warp 0: the swept b'_3 of every chain
  double fe[32], fa[32];        // sweep warp -> warp 0: end value and multiplier product of every forward segment
};
// Segmented sweeps work on a TRANSPOSED copy of every tail (l >= 3) quantity: row i of segment lane `ln` lives at
// i * TEAM_ROW + ln, so that the 32 lanes of a sweep step touch 32 consecutive doubles (the natural layout puts the three
// stride-1 hierarchies 32 doubles apart: every sweep access was an 8-way bank conflict and the sweeps were half of the
// solve).  The odd row pitch keeps the element-wise phases (consecutive l of one chain) conflict-free too.  The right-hand
// side / solution of tail elements exists ONLY there (rt); head elements stay in W.r().
constexpr int TEAM_ROW = 33;
struct TeamWs {
  double* rt;                   // [segrows * TEAM_ROW] right-hand side -> solution, tail elements
  double* mt;                   // backward multipliers m_l (0 on the truncation row)
  double* iet;                  // inverse pivots 1/e_l
  double* gt;                   // forward multipliers g_l
  double* mct;                  // backward: prod of -m over the rows of the segment above and including l
  double* gct;                  // forward:  prod of  g over the rows of the segment up to and including l
  const int* eslot;             // [np] element -> slot of kf (32: not a tail element)
  const int* epos;              // [np] element -> position in the transposed arrays (-1: not a tail element)
  const int* tailpos;           // [ntail] tail row tt -> position of its element
  int segrows;                  // rows of the transposed arrays = longest segment
};
DEB_HD int team_segrows(int lmaxg, int lmaxgp, int lmaxr, int lmaxnu) {
  int L = lmaxg > lmaxgp ? lmaxg : lmaxgp;
  if (lmaxr > L) L = lmaxr;
  if (lmaxnu > L) L = lmaxnu;
  return (L - 2 + TEAM_NSEG - 1) / TEAM_NSEG;
}
DEB_HD size_t team_ws_doubles(int segrows) { return (size_t)6 * segrows * TEAM_ROW; }
DEB_DEV void carve_team(TeamWs& X, double* base, int segrows, const int* eslot, const int* epos, const int* tailpos) {
  const size_t a = (size_t)segrows * TEAM_ROW;
  X.rt = base; X.mt = X.rt + a; X.iet = X.mt + a; X.gt = X.iet + a; X.mct = X.gt + a; X.gct = X.mct + a;
  X.eslot = eslot; X.epos = epos; X.tailpos = tailpos; X.segrows = segrows;
}
// rows [lo, hi) of chain c's tail (l = 3 .. L) that segment s sweeps
DEB_DEV void team_segment(const CtaConst& C, int c, int s, int* lo, int* hi) {
  const int L = C.ch_lmax[c], len = (L - 2 + TEAM_NSEG - 1) / TEAM_NSEG;
  int a = 3 + s * len, b = a + len;
  if (b > L + 1) b = L + 1;
  if (a > b) a = b;
  *lo = a; *hi = b;
}
// position of row l of chain c in the transposed arrays
DEB_DEV int team_pos(const CtaConst& C, int nch, int c, int l) {
  const int L = C.ch_lmax[c], len = (L - 2 + TEAM_NSEG - 1) / TEAM_NSEG;
  const int sg = (l - 3) / len;
  return (l - 3 - sg * len) * TEAM_ROW + sg * nch + c;
}
DEB_DEV void init_team_const(const Problem& P, const CtaConst& C, int* eslot, int* epos, int* tailpos, int tid, int nthreads) {
  for (int e = tid; e < P.np; e += nthreads) {
    int slot = TEAM_NSEG * NCHMAX, pos = -1;
    const int d = elem_desc(P, e);
    if (d >= 0) {
      const int ty = d & 0xff, l = (d >> 8) & 0xff, c = d >> 16;
      if (ty == R_GEN || ty == R_TRUNC) {
        const int L = C.ch_lmax[c], len = (L - 2 + TEAM_NSEG - 1) / TEAM_NSEG;
        slot = ((l - 3) / len) * P.nch + c;
        pos = team_pos(C, P.nch, c, l);
      }
    }
    eslot[e] = slot; epos[e] = pos;
  }
  for (int tt = tid; tt < C.ntail; tt += nthreads) {
    const int info = C.tail[tt];
    tailpos[tt] = team_pos(C, P.nch, (info >> 12) & 0xf, info >> 16);
  }
}

#ifdef DEB_CPU_EMU
#define DEB_TID_PARAM
#define DEB_T_BEGIN for (int tid = 0; tid < NT; ++tid) {
#define DEB_T_END }
#define DEB_T_BAR()
#define DEB_T_BAR_BUT1()
#define DEB_NB_ARRIVE(id, nthr)
#define DEB_NB_SYNC(id, nthr)
#define DEB_TREGS(type, name, dims) type name##_all[NT] dims
#define DEB_TUSE(name) auto& name = name##_all[tid]
#define DEB_IF_WARP(w)
#define DEB_DUO_TOP()
#define DEB_DUO_STAGE(st)
#define DEB_T_OR(name) ([&]() { int a_ = 0; for (int l_ = 0; l_ < NT; ++l_) a_ |= (name##_all[l_] != 0); return a_; }())
#else
#define DEB_TID_PARAM , const int tid
#define DEB_T_BEGIN {
#define DEB_T_END }
// (DUO: two teams share one CTA -- deb_team.cu, k_evolve_duo --, so a team's barriers are named ones, ids bb .. bb+4)
#define DEB_T_BAR() do { if (DUO) asm volatile("bar.sync %0, %1;" ::"r"(bb), "n"(32 * TEAM) : "memory"); else __syncthreads(); } while (0)
// barrier of the team without warp 1 (named barrier 1): warp 1 evaluates the next stage's background from the first
// barrier of a stage to the last and must not hold up the solve
#define DEB_T_BAR_BUT1() do { if (TEAM >= 3) { if (wid != 1) DEB_NB_SYNC(1, 32 * TEAM - 32); } else DEB_T_BAR(); } while (0)
// producer/consumer hand-off between two warps (PTX named barriers): the producer arrives and goes on, the consumer waits
#define DEB_NB_ARRIVE(id, nthr) do { if (DUO) asm volatile("bar.arrive %0, %1;" ::"r"(bb + (id)), "n"(nthr) : "memory"); \
                                     else asm volatile("bar.arrive %0, %1;" ::"n"(id), "n"(nthr) : "memory"); } while (0)
#define DEB_NB_SYNC(id, nthr) do { if (DUO) asm volatile("bar.sync %0, %1;" ::"r"(bb + (id)), "n"(nthr) : "memory"); \
                                   else asm volatile("bar.sync %0, %1;" ::"n"(id), "n"(nthr) : "memory"); } while (0)
#define DEB_TREGS(type, name, dims) type name dims
#define DEB_TUSE(name)
#define DEB_IF_WARP(w) if (wid == (w))
#define DEB_T_OR(name) (DUO ? team_or_named(name, bb, 32 * TEAM) : __syncthreads_or(name))
// the cross-team rendezvous of a two-team CTA (warp 0 of each team; see DuoSync)
#define DEB_DUO_TOP() do { if (DUO) { DEB_IF_WARP(0) duo->top(lane); } } while (0)
#define DEB_DUO_STAGE(st) do { if (DUO) { DEB_IF_WARP(0) duo->stage(st, lane); } } while (0)
#endif

// element e of the stage vector just solved: x0 for e = 0, otherwise r plus the deferred forward-sweep carry
// (own element j of the thread: orp[j] points at its right-hand side / solution slot, ogp[j] at its cumulative forward
//  multiplier -- a zero for head elements --, osl[j] is its carry slot)
#define DEB_KV(e) ((e) == 0 ? x0 : *orp[j] + *ogp[j] * box.kf[osl[j]])
// measurement builds only (-DDEB_TEAM_TIMING): g_team_skip knocks phases out (results are then wrong on purpose) so
// that the drop in cycles per step says what a phase contributes to the critical path (tools/team_timing.py)
#if defined(DEB_TEAM_TIMING) && !defined(DEB_CPU_EMU)
#define DEB_KEEP(bit) (!(g_team_skip & (bit)))
#else
#define DEB_KEEP(bit) true
#endif
// optional phase timing of warp 0 (build with -DDEB_TEAM_TIMING; tools/team_timing.py reads the counters)
#if defined(DEB_TEAM_TIMING) && !defined(DEB_CPU_EMU)
#define DEB_TICK(slot) do { const long long now_ = clock64(); if (tid == 0 && kidx == P.nk - 1) atomicAdd((unsigned long long*)&g_team_timing[slot], (unsigned long long)(now_ - tick_)); tick_ = now_; } while (0)
#define DEB_TICK_INIT long long tick_ = clock64(); long long tick2_ = tick_; long long tick3_ = tick_;
#define DEB_TICK3(slot) do { const long long now_ = clock64(); if (tid == 0 && kidx == P.nk - 1) atomicAdd((unsigned long long*)&g_team_timing[slot], (unsigned long long)(now_ - tick3_)); tick3_ = now_; } while (0)
#define DEB_TICK3_START tick3_ = clock64();
#define DEB_TICK2(slot) do { const long long now_ = clock64(); if (tid == 0 && kidx == P.nk - 1) atomicAdd((unsigned long long*)&g_team_timing[slot], (unsigned long long)(now_ - tick2_)); tick2_ = now_; } while (0)
#define DEB_TICK2_START tick2_ = clock64();
#else
#define DEB_TICK(slot)
#define DEB_TICK2(slot)
#define DEB_TICK2_START
#define DEB_TICK3(slot)
#define DEB_TICK3_START
#define DEB_TICK_INIT
#endif
#define DEB_FOR_TEAM(body) _Pragma("unroll") for (int j = 0; j < NE; ++j) { const int e = tid + NT * j; if (e < n) { body } }

// Rodas5 coefficients of the stage combinations in constant memory: a 64-bit immediate costs two UMOVs in front of
// every DFMA (60 of the 131 instructions of one stage's combination code, each a dependency stall); operands from
// the constant bank cost nothing
enum { TRD_A21, TRD_A31, TRD_A32, TRD_A41, TRD_A42, TRD_A43, TRD_A51, TRD_A52, TRD_A53, TRD_A54, TRD_A61, TRD_A62, TRD_A63, TRD_A64, TRD_A65, TRD_C21, TRD_C31, TRD_C32, TRD_C41, TRD_C42, TRD_C43, TRD_C51, TRD_C52, TRD_C53, TRD_C54, TRD_C61, TRD_C62, TRD_C63, TRD_C64, TRD_C65, TRD_C71, TRD_C72, TRD_C73, TRD_C74, TRD_C75, TRD_C76, TRD_C81, TRD_C82, TRD_C83, TRD_C84, TRD_C85, TRD_C86, TRD_C87, TRD_N };
#ifdef DEB_CPU_EMU
static const double g_trd[TRD_N] = { RD_A21, RD_A31, RD_A32, RD_A41, RD_A42, RD_A43, RD_A51, RD_A52, RD_A53, RD_A54, RD_A61, RD_A62, RD_A63, RD_A64, RD_A65, RD_C21, RD_C31, RD_C32, RD_C41, RD_C42, RD_C43, RD_C51, RD_C52, RD_C53, RD_C54, RD_C61, RD_C62, RD_C63, RD_C64, RD_C65, RD_C71, RD_C72, RD_C73, RD_C74, RD_C75, RD_C76, RD_C81, RD_C82, RD_C83, RD_C84, RD_C85, RD_C86, RD_C87 };
#else
__constant__ double g_trd[TRD_N] = { RD_A21, RD_A31, RD_A32, RD_A41, RD_A42, RD_A43, RD_A51, RD_A52, RD_A53, RD_A54, RD_A61, RD_A62, RD_A63, RD_A64, RD_A65, RD_C21, RD_C31, RD_C32, RD_C41, RD_C42, RD_C43, RD_C51, RD_C52, RD_C53, RD_C54, RD_C61, RD_C62, RD_C63, RD_C64, RD_C65, RD_C71, RD_C72, RD_C73, RD_C74, RD_C75, RD_C76, RD_C81, RD_C82, RD_C83, RD_C84, RD_C85, RD_C86, RD_C87 };
#endif
#define TRD(name) g_trd[TRD_##name]

// the output conversion runs once per mode and output time: kept out of line, away from the step loop
#ifdef DEB_CPU_EMU
inline void team_convert_outputs(const Problem& P, const Cosmo& c, const NuBins& nb, const double* y, double k, double* out) {
  convert_outputs(P, c, nb, y, k, out);
}
#else
static __device__ __noinline__ void team_convert_outputs(const Problem& P, const Cosmo& c, const NuBins& nb, const double* y, double k, double* out) {
  convert_outputs(P, c, nb, y, k, out);
}
#endif

// background at box.a_req -> chain coefficients, operator slots and the scalars of the metric sources (warp 1)
DEB_DEV void team_helper_compute(const Problem& P, const CtaConst& C, const Cosmo& c, TeamBox& box, TeamBg& o, Hints& hint2 DEB_LANE_PARAM) {
  const double k = box.k;
  Bg<double> b;
  compute_bg<double>(c, C.nu, P.nq, o.a_req, hint2, box.ic2, b);
  DEB_LANES_BEGIN
    if (lane < P.nch) chain_a_lane(c, C.nu, b, k, lane, o.kc, o.kap, o.vv, o.sl);
    if (lane == 0) {
      fill_slots<double>(c, b, k, o.sl);
      o.bgs[0] = b.H; o.bgs[1] = b.gc; o.bgs[2] = b.gb; o.bgs[3] = b.gg; o.bgs[4] = b.gr; o.bgs[5] = b.gnu;
      o.bgs[6] = b.gq; o.bgs[7] = b.wq1; o.bgs[8] = b.ca2; o.bgs[9] = b.wq;
    }
  DEB_LANES_END
}

// the a-only half of chain_coeffs_lane<Dual> (warp 1) ...
DEB_DEV void chain_a_lane_dual(const Cosmo& c, const NuBins& nb, const Bg<Dual>& b, double k, int ch, double* kcA, double* kapA,
                               double* vvd, double* sl) {
  Dual kc = 0.0 * b.a + k;
  if (ch >= 3) {
    const int i = ch - 3;
    Dual aq = b.a * (c.amnu / nb.q[i]);
    Dual s2 = 1.0 + aq * aq;
    Dual v = drsqrt(s2);
    Dual iv = s2 * v;
    kc = v * k;
    vvd[i] = v.v; vvd[NQMAX + i] = v.d; vvd[2 * NQMAX + i] = iv.v; vvd[3 * NQMAX + i] = iv.d;
    sl[SL_KV0 + i] = kc.v; sl[NSLOT + SL_KV0 + i] = kc.d;
  }
  Dual kp = ch < 2 ? b.opac : 0.0 * b.a;
  kcA[ch] = kc.v; kcA[NCHMAX + ch] = kc.d;
  kapA[ch] = kp.v; kapA[NCHMAX + ch] = kp.d;
}
// ... and the half that needs the state (warp 0)
DEB_DEV void nu_moments_lane_dual(const NuBins& nb, const double* vvd, const double* u, int iq0, int i, double* nurA, double* nupA) {
  const double wp0 = nb.w[i] * u[iq0 + i];
  const Dual v = mk(vvd[i], vvd[NQMAX + i]), iv = mk(vvd[2 * NQMAX + i], vvd[3 * NQMAX + i]);
  const Dual tr = iv * wp0, tp = v * wp0;
  nurA[i] = tr.v; nurA[NQMAX + i] = tr.d;
  nupA[i] = tp.v; nupA[NQMAX + i] = tp.d;
}
DEB_DEV void team_jac_compute(const Problem& P, const CtaConst& C, const Cosmo& c, TeamBox& box, TeamJac& o, double a, Hints& hint2 DEB_LANE_PARAM) {
  const double k = box.k;
  Bg<Dual> bd;
  compute_bg<Dual>(c, C.nu, P.nq, mk(a, 1.0), hint2, box.ic2, bd);
  DEB_LANES_BEGIN
    if (lane < P.nch) chain_a_lane_dual(c, C.nu, bd, k, lane, o.kc, o.kap, o.vvd, o.sl);
    if (lane == 0) { fill_slots<Dual>(c, bd, k, o.sl); o.bd = bd; }
  DEB_LANES_END
}

#ifndef DEB_CPU_EMU
// Two teams in one CTA (k_evolve_duo): their serial warps rendezvous at the top of every step and at every `every`-th
// stage so that both run the same stretch of the (170 KB, straight-line) step code at the same time and share its
// instruction-cache lines -- the same remedy as the chain-lane kernel's lock-step, here between two warps only.
// The rendezvous is a pair of counters in shared memory on a common virtual clock (K ticks per step, step tops at
// multiples of K): a team posts the tick it reached and waits until the other has posted at least as much.  A team
// between modes (work fetch, initial conditions) or out of work posts INT_MAX = "do not wait for me" and re-attaches at
// the other's next step top.  No data travels through the counters; they are read and written with shared-memory atomics
// by lane 0 only (compute-sanitizer's racecheck stays clean).
struct DuoSync {
  int* mine; int* other;
  int v, K, every; bool attached;
  __device__ __forceinline__ void post_wait(int lane) {
    if (lane == 0) { atomicExch(mine, v); while (atomicOr(other, 0) < v) { } }
    __syncwarp();
  }
  __device__ __forceinline__ void top(int lane) {
    if (attached) v = (v / K + 1) * K;
    else { const int a = __shfl_sync(0xffffffffu, lane == 0 ? atomicOr(other, 0) : 0, 0); v = ((a == 0x7fffffff ? v : a) + K - 1) / K * K; attached = true; }
    post_wait(lane);
  }
  __device__ __forceinline__ void stage(int st, int lane) { if (every && ((st - 1) & (every - 1)) == 0) { ++v; post_wait(lane); } }
  __device__ __forceinline__ void detach(int lane) { if (lane == 0) atomicExch(mine, 0x7fffffff); attached = false; }
};
// __syncthreads_or over one team of a two-team CTA (named barrier `id`, nthr threads)
static __device__ __forceinline__ int team_or_named(int pred, int id, int nthr) {
  int r;
  asm volatile("{ .reg .pred p, q; setp.ne.s32 p, %1, 0; barrier.red.or.pred q, %2, %3, p; selp.s32 %0, 1, 0, q; }"
               : "=r"(r) : "r"(pred), "r"(id), "r"(nthr) : "memory");
  return r;
}
#define DEB_DUO_PARAM , const int bb, DuoSync* duo
#else
#define DEB_DUO_PARAM
#endif

template <int NE, int TEAM, bool DUO = false>
DEB_DEV void integrate_mode_team(const Problem& P, const CtaConst& C, WarpWs& W, const TeamWs& X, TeamBox& box, int mode DEB_TID_PARAM DEB_DUO_PARAM) {
  constexpr int NT = 32 * TEAM;
#ifndef DEB_CPU_EMU
  const int lane = tid & 31, wid = tid >> 5;
#endif
  const int n = P.n, nh = P.nh, nch = P.nch, nq = P.nq;
  const int nhb = nh - 1;
  const int cosmo = mode / P.nk, kidx = mode - cosmo * P.nk;
  const double k = DEB_LDG(P.kmodes + (P.k_per_cosmo ? (size_t)cosmo * P.nk + kidx : (size_t)kidx));
  const double k2 = k * k;
  DEB_T_BEGIN
    if (tid == 0) { *W.cosmo() = load_cosmo(P, cosmo); box.k = k; }
    for (int e = tid; e < 6 * X.segrows * TEAM_ROW; e += NT) X.rt[e] = 0.0;        // all six transposed arrays
    if (tid < TEAM_KF) box.kf[tid] = 0.0;
  DEB_T_END
  DEB_T_BAR();
  const Cosmo& c = *W.cosmo();
  const NuBins& nb = C.nu;
  const double* tout = P.tau_out + (size_t)cosmo * P.nout;

  DEB_REGS(int, pcol, );
  DEB_REGS(double, rscale, );
  DEB_REGS(unsigned, pkey, );
  DEB_REGS(int, pivl, );
  DEB_REGS(double, fmul, );
  DEB_REGS(double, pval, );
  DEB_REGS(double, s1, ); DEB_REGS(double, s2, ); DEB_REGS(double, s3, ); DEB_REGS(double, s4, );
  DEB_TREGS(int, nanflag, );
  // lane constants of the segmented sweeps: lane = segment * nch + chain sweeps rows [sLo, sHi) of its chain
  DEB_REGS(int, sSeg, ); DEB_REGS(int, sLen, ); DEB_REGS(int, sI2, ); DEB_REGS(int, sCh, ); DEB_REGS(int, sLo, );
  DEB_LANES_BEGIN
    DEB_USE(sSeg); DEB_USE(sLen); DEB_USE(sI2); DEB_USE(sCh); DEB_USE(sLo);
    sSeg = -1; sLen = 0; sI2 = 0; sCh = 0; sLo = 3;
    if (lane < TEAM_NSEG * nch) {
      sSeg = lane / nch;
      const int ch = lane - sSeg * nch;
      int lo, hi;
      team_segment(C, ch, sSeg, &lo, &hi);
      sLen = hi - lo; sCh = ch; sLo = lo;
      sI2 = C.ch_base[ch] + 2 * C.ch_stride[ch];      // the chain's l = 2 element (a head row)
    }
  DEB_LANES_END
  // where the thread's own elements keep their right-hand side / solution
  DEB_TREGS(double*, orp, [NE]); DEB_TREGS(const double*, ogp, [NE]); DEB_TREGS(int, osl, [NE]);
  DEB_T_BEGIN
    DEB_TUSE(orp); DEB_TUSE(ogp); DEB_TUSE(osl);
#pragma unroll
    for (int j = 0; j < NE; ++j) {
      const int e = tid + NT * j;
      orp[j] = W.r(); ogp[j] = &box.kf[TEAM_NSEG * NCHMAX]; osl[j] = TEAM_NSEG * NCHMAX;
      if (e < n) {
        const int ep = X.epos[e];
        osl[j] = X.eslot[e];
        if (ep >= 0) { orp[j] = X.rt + ep; ogp[j] = X.gct + ep; } else { orp[j] = W.r() + e; }
      }
    }
  DEB_T_END

  // ---- prologue (every thread evaluates the scalars; the state is filled element-wise) ----
  double t1 = DEB_LDG(tout);
  double tmin_out = t1;
  for (int j = 1; j < P.nout; ++j) { double tj = DEB_LDG(tout + j); t1 = fmax(t1, tj); tmin_out = fmin(tmin_out, tj); }
  double t, tnext;
  const size_t item = (size_t)mode;
  if (P.mode == 1) {
    t = DEB_LDG(P.dbg_t0 + mode); tnext = DEB_LDG(P.dbg_t1 + mode); t1 = tnext;
    DEB_T_BEGIN
      for (int e = tid; e < n; e += NT) W.y()[e] = DEB_LDG(P.dbg_y0 + (size_t)mode * n + e);
    DEB_T_END
  } else {
    const double st0 = start_time(c, k, DEB_LDG(P.lt_small + cosmo));
    double tau_start = 0.99 * fmin(tmin_out, st0);
    if (!(st0 == st0)) tau_start = st0;
    IcScalars ics = ic_scalars(c, tau_start, k);
    DEB_T_BEGIN
      for (int e = tid; e < n; e += NT) W.y()[e] = ic_value(P, c, nb, ics, elem_desc(P, e), k);
    DEB_T_END
    if (P.mode == 2) {
      DEB_T_BEGIN
        if (tid == 0) P.dbg_tau_start[item] = tau_start;
        for (int e = tid; e < n; e += NT) P.dbg_ics[item * n + e] = W.y()[e];
      DEB_T_END
      DEB_T_BAR();
      return;
    }
    t = tau_start;
    tnext = t + fmin(t / 4.0, 0.5 * (t1 - t));          // dt0 (perturbations.py:756)
    if (tnext > t1 - 1e-10) tnext = t1;
  }
  DEB_T_BAR();

  double inv_prev = 1.0, inv_pprev = 1.0, linv_prev = 0.0;
  int nsteps = 0, nacc = 0, save_idx = 0, status = 0;
  if (!(t == t) || !(tnext == tnext) || !(t1 == t1)) status = 2;
  Hints hint2; hint2.th = -1; hint2.nu = -1;       // warp 1 (all background evaluations)
  int jcur = 0;
  DEB_IF_WARP(TEAM > 1 ? 1 : 0) {
    team_jac_compute(P, C, c, box, box.jac[0], W.y()[0], hint2 DEB_LANE_ARG);
  }
  DEB_T_BAR();

  while (t < t1 && nsteps < P.max_steps && status == 0) {
    DEB_DUO_TOP();
    if (P.mode == 3) {
      if (nsteps >= DEB_LDG(P.rp_n + mode)) break;
      tnext = DEB_LDG(P.rp_tnext + (size_t)mode * P.rp_stride + nsteps);
      if (fabs(tnext - t1) <= 1e-12 * fabs(t1)) tnext = t1;
    }
    DEB_TREGS(double, ks, [7][NE]);    // stage vectors k_1..k_7 of the thread's own elements
    const double dt = tnext - t;
    const double invdt = DEB_RCP(dt);
    const double idg = DEB_RCP(dt * RD_GAMMA);      // diagonal of W = I/(gamma dt) - J
    const double invt0 = DEB_RCP(t);
    const double gdt = dt * RD_GAMMA;

    // ================= Jacobian pieces at (t, y) =================
    // (the background at a = y_0, the chain coefficients and the operator slots were posted by warp 1 already)
    DEB_TICK_INIT
    const TeamJac& J = box.jac[jcur];
    // row 0 is closed (a' = H a): x_0 of stage 1 is known at once, and with it the d f/d a column of every row
    const Dual f00 = J.bd.H * J.bd.a;
    const double x0piv = idg - f00.d;                // W_00 = 1/(gamma dt) - d(H a)/da
    double x0 = f00.v / x0piv;
    DEB_T_BEGIN
      if (tid == 0) { W.r()[0] = f00.v; W.ja()[0] = f00.d; }
    DEB_T_END
    DEB_IF_WARP(0) {
      const Bg<Dual> bd = J.bd;
      DEB_LANES_BEGIN
        if (lane >= 3 && lane < nch) nu_moments_lane_dual(nb, J.vvd, W.y(), P.iq0, lane - 3, W.nur(), W.nup());
      DEB_LANES_END
      Metric<Dual> md;
      compute_metric<Dual>(P, c, nb, bd, W.y(), k, W.nur(), W.nup(), md);
      const double H = bd.H.v, a = bd.a.v;
      DEB_LANES_BEGIN
        if (lane < nh) {                      // head rows: f -> r (stage-1 right-hand side), d f/d a -> ja
          const int e = C.hidx[lane];
          Dual f = head_row<Dual>(C, J.sl, W.y(), lane, md);
          W.r()[e] = f.v + f.d * x0; W.ja()[e] = f.d;
        }
        if (lane < nh) {                      // head-column gradients of h', eta' and of row 1 (value parts only)
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
            case R_P0: { const double vb = J.kc[3 + bin] / k; wr = bd.gnu.v * nb.w[bin] / vb; wp = bd.gnu.v * nb.w[bin] * vb / 3.0; } break;
            case R_P1: wt = bd.gnu.v * k * nb.w[bin]; break;
            case R_DQ: wr = bd.gq.v; wp = c.cs2de * bd.gq.v; break;
            case R_TQ: wt = bd.wq1.v * bd.gq.v; wp = (c.cs2de - bd.ca2.v) * 3.0 * H * wt / k2; break;
            default: break;
          }
          W.gh()[lane] = wr / H + extra;
          W.ge()[lane] = 0.5 * wt / k2;
          W.j1()[lane] = -(wr + 3.0 * wp) * a;
        }
      DEB_LANES_END
    }
    DEB_IF_WARP(TEAM > 1 ? 1 : 0) {
      // ---- tails: pivot-free backward elimination l = L .. 3, factors into the transposed arrays.  Only the pivot
      //      recurrence e_l = d_l + w_l / e_{l+1}, w_l = W_{l,l+1} W_{l+1,l}, is sequential (one lane per chain: a DFMA and
      //      a reciprocal per row); the products w_l before it and the multipliers m_l, g_l with their cumulative
      //      products after it are made by one lane per (segment, chain).
      DEB_LANES_BEGIN
        DEB_USE(sSeg); DEB_USE(sLen); DEB_USE(sCh); DEB_USE(sLo);
        if (sSeg >= 0 && DEB_KEEP(128)) {
          const int L = C.ch_lmax[sCh];
          const double kc = J.kc[sCh];
#pragma unroll 1
          for (int i = 0; i < sLen; ++i) {
            const int l = sLo + i;
            const double clx = (l + 1 >= L) ? 1.0 : C.cl[l + 1];          // W_{l+1,l} = -kc cl_{l+1} (truncation row: -kc)
            X.mct[i * TEAM_ROW + lane] = l < L ? (kc * C.ch[l]) * (kc * clx) : 0.0;      // (scratch: overwritten below)
          }
        }
      DEB_LANES_END
      DEB_LANES_BEGIN
        if (lane < nch && DEB_KEEP(128)) {
          const int L = C.ch_lmax[lane], len = (L - 2 + TEAM_NSEG - 1) / TEAM_NSEG;
          const double kc = J.kc[lane], base = idg + J.kap[lane];
          double ie = DEB_RCP(base + (double)(L + 1) * invt0);
          int sg = (L - 3) / len, i = L - 3 - sg * len;            // position of row L
          int pos = i * TEAM_ROW + sg * nch + lane;
          X.iet[pos] = ie;
#pragma unroll 1
          for (int l = L - 1; l >= 3; --l) {
            if (--i < 0) { i = len - 1; --sg; }
            pos = i * TEAM_ROW + sg * nch + lane;
            ie = DEB_RCP(base + X.mct[pos] * ie);
            X.iet[pos] = ie;
          }
          const int i2 = C.ch_base[lane] + 2 * C.ch_stride[lane];
          const double mm = (kc * C.ch[2]) * ie;
          W.m()[i2] = mm;
          W.ie()[i2] = mm * (kc * (3 >= L ? 1.0 : C.cl[3]));     // Schur increment for the head diagonal (l = 2 row)
        }
      DEB_LANES_END
      DEB_LANES_BEGIN
        DEB_USE(sSeg); DEB_USE(sLen); DEB_USE(sCh); DEB_USE(sLo);
        if (sSeg >= 0 && DEB_KEEP(128)) {
          const int L = C.ch_lmax[sCh];
          const double kc = J.kc[sCh];
          double acc = 1.0;
#pragma unroll 1
          for (int i = sLen - 1; i >= 0; --i) {          // m_l = W_{l,l+1} / e_{l+1} and its products from the top of the segment
            const int l = sLo + i;
            const double ie_up = l >= L ? 0.0 : (i + 1 < sLen ? X.iet[(i + 1) * TEAM_ROW + lane] : X.iet[lane + nch]);
            const double mm = l >= L ? 0.0 : (kc * C.ch[l]) * ie_up;
            X.mt[i * TEAM_ROW + lane] = mm;
            acc = -mm * acc;
            X.mct[i * TEAM_ROW + lane] = acc;
          }
          acc = 1.0;
#pragma unroll 1
          for (int i = 0; i < sLen; ++i) {               // g_l = -W_{l,l-1} / e_l and its products from the bottom
            const int l = sLo + i;
            const double gg = (kc * (l >= L ? 1.0 : C.cl[l])) * X.iet[i * TEAM_ROW + lane];
            X.gt[i * TEAM_ROW + lane] = gg;
            acc = gg * acc;
            X.gct[i * TEAM_ROW + lane] = acc;
          }
        }
      DEB_LANES_END
    }
    {
      // tail rows of f and d f/d a, two rows per trip
      const double d1t = (dt * RD_D1) * invt0 * invt0;
      // (warp 0 is busy with the head and warp 1 with the tail factors: the rows go to the other warps)
      constexpr int TJ0 = TEAM >= 3 ? 64 : 0, NTJ = NT - TJ0;
      DEB_T_BEGIN
        if (tid >= TJ0)
        for (int tt = tid - TJ0; tt < C.ntail; tt += 2 * NTJ) {
          int e0, e1 = 0; double tr0, tr1 = 0.0;
          const bool two = tt + NTJ < C.ntail;
          Dual f0 = tail_row<Dual>(C, J.kc, J.kap, W.y(), C.tail[tt], invt0, &e0, &tr0);
          Dual f1 = mk(0.0, 0.0);
          if (two) f1 = tail_row<Dual>(C, J.kc, J.kap, W.y(), C.tail[tt + NTJ], invt0, &e1, &tr1);
          X.rt[X.tailpos[tt]] = f0.v + d1t * tr0 * W.y()[e0] + f0.d * x0; W.ja()[e0] = f0.d;
          if (two) { X.rt[X.tailpos[tt + NTJ]] = f1.v + d1t * tr1 * W.y()[e1] + f1.d * x0; W.ja()[e1] = f1.d; }
        }
      DEB_T_END
    }
    DEB_TICK(0);
    DEB_T_BAR();                                      // r (incl. the d f/d a column), ja, tail factors complete
    DEB_TICK(1);

    double ci_hh = 0.0, ci_he = 0.0, ci_eh = 0.0, ci_ee = 0.0, jq_h = 0.0, jq_e = 0.0;
    DEB_IF_WARP(0) {
      // ---- head:  W_h = D - chv gh^T - cev ge^T  (+ the a h' row); D block diagonal -> block inverses + Woodbury
      DEB_LANES_BEGIN
        DEB_USE(pcol); DEB_USE(rscale);
        pcol = -1; rscale = 1.0;
        if (lane < nhb) {
          const int ty = C.htype[lane], lo = C.blo[lane];
          double* row = hrow(W, lane, lo);
#pragma unroll
          for (int i = 0; i < 8; ++i) row[lo + i] = 0.0;
          double diag = idg;
          if (ty == R_F2 || ty == R_G2 || ty == R_N2 || ty == R_P2) {      // Schur complement of the chain tail
            const int chain = ty == R_F2 ? 0 : (ty == R_G2 ? 1 : (ty == R_N2 ? 2 : 3 + C.hbin[lane]));
            diag += W.ie()[C.ch_base[chain] + 2 * C.ch_stride[chain]];
          }
          row[lane] = diag;
#pragma unroll
          for (int q = 0; q < HOP_NT; ++q) {
            const int m = C.hop_meta[q][lane], hc = (m >> 20) - 1;
            if (hc >= 0) row[hc] -= C.hop_c[q][lane] * J.sl[m & 0xff];
          }
        }
      DEB_LANES_END
      DEB_GJ_ALL(DEB_KEEP(64))
      DEB_LANES_BEGIN
        W.xb()[lane] = 0.0;
      DEB_LANES_END
      DEB_LANES_BEGIN
        DEB_USE(pcol); DEB_USE(s1); DEB_USE(s2); DEB_USE(s3); DEB_USE(s4);
        s1 = s2 = s3 = s4 = 0.0;
        if (lane < nhb) {
          const int lo = C.blo[lane], hi = C.bhi[lane];
          const double* row = hrow(W, lane, lo);
          double ah = 0.0, ae = 0.0;
          for (int cc = lo; cc < hi; ++cc) {
            const int pr = W.perm()[cc];
            ah += row[cc] * (C.hop_chc[pr] * J.sl[C.hop_chs[pr]]);
            ae += row[cc] * C.hop_cec[pr];
          }
          W.qh()[pcol] = ah; W.qe()[pcol] = ae;
          const double ghc = W.gh()[pcol], gec = W.ge()[pcol], j1c = W.j1()[pcol];
          s1 = ghc * ah; s2 = ghc * ae; s3 = gec * ah; s4 = gec * ae;
          W.xb()[lane] = j1c;
        }
      DEB_LANES_END
      {
        const double c_hh = 1.0 - DEB_WARP_SUM(s1), c_he = -DEB_WARP_SUM(s2), c_eh = -DEB_WARP_SUM(s3), c_ee = 1.0 - DEB_WARP_SUM(s4);
        const double idet = DEB_RCP(c_hh * c_ee - c_he * c_eh);
        ci_hh = c_ee * idet; ci_he = -c_he * idet; ci_eh = -c_eh * idet; ci_ee = c_hh * idet;
        DEB_LANES_BEGIN
          DEB_USE(pcol); DEB_USE(s1); DEB_USE(s2);
          s1 = s2 = 0.0;
          if (lane < nhb) { const double j1c = W.xb()[lane]; s1 = j1c * W.qh()[pcol]; s2 = j1c * W.qe()[pcol]; }
        DEB_LANES_END
        jq_h = DEB_WARP_SUM(s1); jq_e = DEB_WARP_SUM(s2);
      }
    }

    DEB_TICK(2);
    // ================= 8 stages =================
    double errnorm2 = 0.0;
#pragma unroll 1
    for (int st = 1; st <= 8; ++st) {
      DEB_DUO_STAGE(st);
      DEB_TICK3_START
      if (st > 1) {
        TeamBg& cur = box.bg[st & 1];                  // posted by warp 1 during stage st-1
        const double ts = st == 2 ? t + RD_CT2 * dt : st == 3 ? t + RD_CT3 * dt : st == 4 ? t + RD_CT4 * dt
                        : st == 5 ? t + RD_CT5 * dt : t + dt;
        const double dtd = st == 2 ? dt * RD_D2 : st == 3 ? dt * RD_D3 : st == 4 ? dt * RD_D4 : st == 5 ? dt * RD_D5 : 0.0;
        // ---- own elements: keep k_{st-1}, stage state u, C-combinations -> r; thread 0 closes row 0: x_0 of this stage ----
#define DEB_ROW0 if (e == 0) { W.r()[0] += cur.bgs[0] * cur.a_req; W.u()[0] = cur.a_req; }
        DEB_T_BEGIN
          DEB_TUSE(ks); DEB_TUSE(orp); DEB_TUSE(ogp); DEB_TUSE(osl);
          if (DEB_KEEP(1))
          switch (st) {
            case 2: DEB_FOR_TEAM(ks[0][j] = DEB_KV(e);
                                 W.u()[e] = W.y()[e] + TRD(A21) * ks[0][j];
                                 *orp[j] = invdt * (TRD(C21) * ks[0][j]); DEB_ROW0) break;
            case 3: DEB_FOR_TEAM(ks[1][j] = DEB_KV(e);
                                 W.u()[e] = W.y()[e] + TRD(A31) * ks[0][j] + TRD(A32) * ks[1][j];
                                 *orp[j] = invdt * (TRD(C31) * ks[0][j] + TRD(C32) * ks[1][j]); DEB_ROW0) break;
            case 4: DEB_FOR_TEAM(ks[2][j] = DEB_KV(e);
                                 W.u()[e] = W.y()[e] + TRD(A41) * ks[0][j] + TRD(A42) * ks[1][j] + TRD(A43) * ks[2][j];
                                 *orp[j] = invdt * (TRD(C41) * ks[0][j] + TRD(C42) * ks[1][j] + TRD(C43) * ks[2][j]); DEB_ROW0) break;
            case 5: DEB_FOR_TEAM(ks[3][j] = DEB_KV(e);
                                 W.u()[e] = W.y()[e] + TRD(A51) * ks[0][j] + TRD(A52) * ks[1][j] + TRD(A53) * ks[2][j] + TRD(A54) * ks[3][j];
                                 *orp[j] = invdt * (TRD(C51) * ks[0][j] + TRD(C52) * ks[1][j] + TRD(C53) * ks[2][j] + TRD(C54) * ks[3][j]); DEB_ROW0) break;
            case 6: DEB_FOR_TEAM(ks[4][j] = DEB_KV(e);
                                 W.u()[e] = W.y()[e] + TRD(A61) * ks[0][j] + TRD(A62) * ks[1][j] + TRD(A63) * ks[2][j] + TRD(A64) * ks[3][j] + TRD(A65) * ks[4][j];
                                 *orp[j] = invdt * (TRD(C61) * ks[0][j] + TRD(C62) * ks[1][j] + TRD(C63) * ks[2][j] + TRD(C64) * ks[3][j] + TRD(C65) * ks[4][j]); DEB_ROW0) break;
            case 7: DEB_FOR_TEAM(ks[5][j] = DEB_KV(e);
                                 W.u()[e] = W.u()[e] + ks[5][j];
                                 *orp[j] = invdt * (TRD(C71) * ks[0][j] + TRD(C72) * ks[1][j] + TRD(C73) * ks[2][j] + TRD(C74) * ks[3][j] + TRD(C75) * ks[4][j] + TRD(C76) * ks[5][j]); DEB_ROW0) break;
            default: DEB_FOR_TEAM(ks[6][j] = DEB_KV(e);
                                 W.u()[e] = W.u()[e] + ks[6][j];
                                 *orp[j] = invdt * (TRD(C81) * ks[0][j] + TRD(C82) * ks[1][j] + TRD(C83) * ks[2][j] + TRD(C84) * ks[3][j] + TRD(C85) * ks[4][j] + TRD(C86) * ks[5][j] + TRD(C87) * ks[6][j]); DEB_ROW0) break;
          }
        DEB_T_END
#undef DEB_ROW0
        DEB_TICK(3);
        DEB_T_BAR();
        DEB_TICK(4);                                   // u, r (C-combinations; row 0 complete) visible
        x0 = W.r()[0] / x0piv;                         // every thread: the same bits
        // ---- f(ts, u) + dt d_i dT + (d f/d a) x0 added onto r:  warp 0 the head rows, warps >= 2 the tail rows ----
        const double invts = DEB_RCP(ts);
        const double dtt = dtd * invt0 * invt0;
        DEB_IF_WARP(0) {
          Bg<double> b;
          b.a = cur.a_req;
          b.H = cur.bgs[0]; b.gc = cur.bgs[1]; b.gb = cur.bgs[2]; b.gg = cur.bgs[3]; b.gr = cur.bgs[4]; b.gnu = cur.bgs[5];
          b.gq = cur.bgs[6]; b.wq1 = cur.bgs[7]; b.ca2 = cur.bgs[8]; b.wq = cur.bgs[9];
          b.opac = cur.sl[SL_OPAC]; b.pbo = cur.sl[SL_PBO]; b.cs2 = 0.0;
          DEB_LANES_BEGIN
            if (lane >= 3 && lane < nch) nu_moments_lane(nb, cur.vv, W.u(), P.iq0, lane - 3, W.nur(), W.nup());
          DEB_LANES_END
          Metric<double> mt;
          mt.hp = mt.ep = mt.al = mt.f1 = 0.0;
          if (DEB_KEEP(8)) compute_metric<double>(P, c, nb, b, W.u(), k, W.nur(), W.nup(), mt);
          DEB_LANES_BEGIN
            if (lane < nh && DEB_KEEP(8)) {
              const int e = C.hidx[lane];
              const double rv = W.r()[e] + head_row<double>(C, cur.sl, W.u(), lane, mt);
              W.r()[e] = rv + W.ja()[e] * x0;
            }
          DEB_LANES_END
        }
        constexpr int TS0 = TEAM >= 3 ? 64 : (TEAM == 2 ? 32 : 0), NTS = NT - TS0;
        DEB_T_BEGIN
          if (tid >= TS0 && DEB_KEEP(2))
          for (int tt = tid - TS0; tt < C.ntail; tt += 2 * NTS) {
            int e0, e1 = 0; double tr0, tr1 = 0.0, f1 = 0.0, r1 = 0.0, y1 = 0.0;
            const bool two = tt + NTS < C.ntail;
            const double f0 = tail_row<double>(C, cur.kc, cur.kap, W.u(), C.tail[tt], invts, &e0, &tr0);
            const int p0 = X.tailpos[tt];
            int p1 = 0;
            const double r0 = X.rt[p0], y0 = W.y()[e0];
            if (two) { f1 = tail_row<double>(C, cur.kc, cur.kap, W.u(), C.tail[tt + NTS], invts, &e1, &tr1); p1 = X.tailpos[tt + NTS]; r1 = X.rt[p1]; y1 = W.y()[e1]; }
            X.rt[p0] = r0 + f0 + dtt * tr0 * y0 + W.ja()[e0] * x0;
            if (two) X.rt[p1] = r1 + f1 + dtt * tr1 * y1 + W.ja()[e1] * x0;
          }
        DEB_T_END
      }

      // ---- warp 1: background scalars of stage st+1 at a = y_0 + sum_j a_{st+1,j} k_{j,0}  (k_{st,0} = x0) ----
      if (st < 8) {
        DEB_IF_WARP(TEAM > 1 ? 1 : 0) {
          TeamBg& nxt = box.bg[(st + 1) & 1];
          DEB_LANES_BEGIN
            if (lane == 0) {
              box.ka0[st - 1] = x0;
              const double* ka = box.ka0;
              double an;
              switch (st) {
                case 1: an = W.y()[0] + RD_A21 * x0; break;
                case 2: an = W.y()[0] + RD_A31 * ka[0] + RD_A32 * x0; break;
                case 3: an = W.y()[0] + RD_A41 * ka[0] + RD_A42 * ka[1] + RD_A43 * x0; break;
                case 4: an = W.y()[0] + RD_A51 * ka[0] + RD_A52 * ka[1] + RD_A53 * ka[2] + RD_A54 * x0; break;
                case 5: an = W.y()[0] + RD_A61 * ka[0] + RD_A62 * ka[1] + RD_A63 * ka[2] + RD_A64 * ka[3] + RD_A65 * x0; break;
                default: an = W.u()[0] + x0; break;        // stages 7 and 8: u + k
              }
              nxt.a_req = an;
            }
          DEB_LANES_END
          if (DEB_KEEP(4)) team_helper_compute(P, C, c, box, nxt, hint2 DEB_LANE_ARG);
        }
      } else {
        // stage 8: the candidate's scale factor y1_0 = u_0 + x0 is known -> background of the next step's Jacobian
        DEB_IF_WARP(TEAM > 1 ? 1 : 0) {
          if (DEB_KEEP(4)) team_jac_compute(P, C, c, box, box.jac[jcur ^ 1], W.u()[0] + x0, hint2 DEB_LANE_ARG);
        }
      }
      DEB_TICK(5);
      // ---- solve W x = r in place.  x_0 and its column are already in.  Every hierarchy tail is swept in TEAM_NSEG
      //      segments at once (lane = segment * nch + chain): local recurrences from zero, a carry chain over the
      //      segments, then x_l = local_l + (product of multipliers) * carry.  The sweeps need the tail rows only, so
      //      the LAST warp runs them while warp 0 is still busy with the metric sources and the head rows (stage 1: the
      //      block inverses): backward sweep + carries, then the forward recurrences from zero, which do not depend on
      //      the head either.  Warp 0 takes the swept l = 3 values into the head solve and finishes with the forward
      //      carry chain; the forward carries are applied by whoever reads the result (DEB_KV).
      constexpr int WS = TEAM - 1;                     // the sweep warp
      if (st > 1 && TEAM >= 4) {                       // tail rows (warps 2 .. TEAM-1) complete before the sweeps start
        DEB_IF_WARP(WS) { DEB_NB_SYNC(2, 32 * (TEAM - 2)); }
#ifndef DEB_CPU_EMU
        else if (wid >= 2) DEB_NB_ARRIVE(2, 32 * (TEAM - 2));
#endif
      } else if (st > 1) {
        DEB_T_BAR_BUT1();
      }
      DEB_IF_WARP(WS) {
        DEB_REGS(double, sT, ); DEB_REGS(double, sE, ); DEB_REGS(double, sA, ); DEB_REGS(double, sK, );
        DEB_LANES_BEGIN          // backward sweep: b'_l = b_l - m_l b'_{l+1}
          DEB_USE(sT); DEB_USE(sE); DEB_USE(sA); DEB_USE(sK); DEB_USE(sSeg); DEB_USE(sLen);
          sE = 0.0; sA = 1.0; sK = 0.0;
          if (sLen > 0 && DEB_KEEP(16)) {
            double* rp = X.rt + ((sLen - 1) * TEAM_ROW + lane);
            const double* mp = X.mt + ((sLen - 1) * TEAM_ROW + lane);
            double p = 0.0;
            double rn = *rp, mn = *mp;
            for (int i = sLen - 1; i > 0; --i) {
              const double rc = rn, mc = mn;
              rn = *(rp - TEAM_ROW); mn = *(mp - TEAM_ROW);
              p = rc - mc * p;
              *rp = p;
              rp -= TEAM_ROW; mp -= TEAM_ROW;
            }
            p = rn - mn * p;
            *rp = p;
            sE = p; sA = X.mct[lane];
          }
          sT = sE;
        DEB_LANES_END
        for (int sg = TEAM_NSEG - 2; sg >= 0; --sg) {        // carry chain: the true b' just above segment sg
          DEB_LANES_BEGIN
            DEB_USE(sT); DEB_USE(sE); DEB_USE(sA); DEB_USE(sK); DEB_USE(sSeg);
            const double kin = DEB_SHFL(sT, (lane + nch) & 31);
            if (sSeg == sg) { sK = kin; sT = sE + sA * kin; }
          DEB_LANES_END
        }
        DEB_LANES_BEGIN
          DEB_USE(sT);
          if (lane < nch) box.t3[lane] = sT;           // b'_3 of the chain, for its l = 2 row in the head
        DEB_LANES_END
        if (TEAM >= 4) DEB_NB_ARRIVE(3, 64);
        DEB_LANES_BEGIN          // forward recurrences from zero: x_l = b'_l/e_l + g_l x_{l-1}
          DEB_USE(sK); DEB_USE(sSeg); DEB_USE(sLen);
          const double kb = sK;          // backward carry of this lane's segment
          double x = 0.0, a = 1.0;
          if (sLen > 0 && DEB_KEEP(16)) {
            double* rp = X.rt + lane;
            const double* ip = X.iet + lane;
            const double* gp = X.gt + lane;
            const double* cp = X.mct + lane;
            double cn = (*rp + *cp * kb) * *ip, gn = *gp;
            for (int i = 0; i < sLen - 1; ++i) {
              const double cc = cn, gc = gn;
              cn = (*(rp + TEAM_ROW) + *(cp + TEAM_ROW) * kb) * *(ip + TEAM_ROW); gn = *(gp + TEAM_ROW);
              x = cc + gc * x;
              *rp = x;
              rp += TEAM_ROW; ip += TEAM_ROW; gp += TEAM_ROW; cp += TEAM_ROW;
            }
            x = cn + gn * x;
            *rp = x;
            a = X.gct[(sLen - 1) * TEAM_ROW + lane];
          }
          box.fe[lane] = x; box.fa[lane] = a;
        DEB_LANES_END
        if (TEAM >= 4) DEB_NB_ARRIVE(4, 64);
      }
      DEB_TICK(6);
      DEB_IF_WARP(0) {
        if (TEAM >= 4) DEB_NB_SYNC(3, 64);             // the swept tails are in
        DEB_TICK2_START
        DEB_LANES_BEGIN
          DEB_USE(sI2);
          if (lane < nch) W.r()[sI2] = W.r()[sI2] - W.m()[sI2] * box.t3[lane];        // l = 2 (a head row): b'_2 = b_2 - m_2 b'_3
        DEB_LANES_END
        DEB_TICK2(11);
        // head: p = D^-1 b (block inverses), then the rank-2 Woodbury correction and the a h' row
        if (DEB_KEEP(32)) {
        DEB_LANES_BEGIN
          W.xb()[lane] = lane < nhb ? W.r()[C.hidx[W.perm()[lane]]] : 0.0;
          if (lane < 8) W.xb()[NHMAX + lane] = 0.0;
        DEB_LANES_END
        DEB_LANES_BEGIN
          DEB_USE(pcol); DEB_USE(s1); DEB_USE(s2); DEB_USE(s3); DEB_USE(pval);
          s1 = s2 = s3 = 0.0; pval = 0.0;
          if (lane < nhb) {
            const int lo = C.blo[lane];
            const double* rb = hrow(W, lane, lo) + lo;
            const double* xb = W.xb() + lo;
            double ra[8], xa[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { ra[i] = rb[i]; xa[i] = xb[i]; }
            const double acc = ((ra[0] * xa[0] + ra[1] * xa[1]) + (ra[2] * xa[2] + ra[3] * xa[3]))
                             + ((ra[4] * xa[4] + ra[5] * xa[5]) + (ra[6] * xa[6] + ra[7] * xa[7]));
            pval = acc;
            s1 = W.gh()[pcol] * acc; s2 = W.ge()[pcol] * acc; s3 = W.j1()[pcol] * acc;
          }
        DEB_LANES_END
        DEB_TICK2(12);
        {
          const double th = DEB_WARP_SUM(s1), te = DEB_WARP_SUM(s2), ta = DEB_WARP_SUM(s3);
          const double sh = ci_hh * th + ci_he * te, se = ci_eh * th + ci_ee * te;
          DEB_LANES_BEGIN
            DEB_USE(pcol); DEB_USE(pval);
            if (lane < nhb) W.r()[C.hidx[pcol]] = pval + W.qh()[pcol] * sh + W.qe()[pcol] * se;
            else if (lane == nhb) W.r()[1] = (W.r()[1] + ta + jq_h * sh + jq_e * se) * gdt;     // a h' (state 1): closed row
          DEB_LANES_END
        }
        }
        DEB_TICK2(13);
        if (TEAM >= 4) DEB_NB_SYNC(4, 64);             // the forward recurrences are in
        // forward carry chain: the true x just below every segment (x_2 of the head for segment 0)
        DEB_REGS(double, fT, ); DEB_REGS(double, fK, );
        DEB_LANES_BEGIN
          DEB_USE(fT); DEB_USE(fK); DEB_USE(sSeg); DEB_USE(sI2);
          fK = 0.0; fT = 0.0;
          if (sSeg == 0) { fK = W.r()[sI2]; fT = box.fe[lane] + box.fa[lane] * fK; }
        DEB_LANES_END
        for (int sg = 1; sg < TEAM_NSEG; ++sg) {
          DEB_LANES_BEGIN
            DEB_USE(fT); DEB_USE(fK); DEB_USE(sSeg);
            const double kin = DEB_SHFL(fT, (lane - nch) & 31);
            if (sSeg == sg) { fK = kin; fT = box.fe[lane] + box.fa[lane] * kin; }
          DEB_LANES_END
        }
        DEB_LANES_BEGIN
          DEB_USE(fK); DEB_USE(sSeg);
          if (sSeg >= 0) box.kf[lane] = fK;
        DEB_LANES_END
        DEB_TICK2(14);
      }
      DEB_TICK(7);
      DEB_T_BAR();                                     // r, kf = k_st (element 0: x0), warp 1's results posted
      DEB_TICK(8);
      DEB_TICK3(16 + st);
    }

    // y1 = u + k8 -> u ; err = k8 -> r
    DEB_T_BEGIN
      DEB_TUSE(nanflag); DEB_TUSE(orp); DEB_TUSE(ogp); DEB_TUSE(osl);
      nanflag = 0;
      DEB_FOR_TEAM(const double kv = DEB_KV(e); W.r()[e] = kv;
                   const double y1v = W.u()[e] + kv; W.u()[e] = y1v; nanflag |= (y1v != y1v);)
    DEB_T_END
    bool anynan = DEB_T_OR(nanflag) != 0;              // (a barrier)

    if (P.mode == 1) {
      DEB_T_BEGIN
        for (int e = tid; e < n; e += NT) { P.dbg_y1[item * n + e] = W.u()[e]; P.dbg_err[item * n + e] = W.r()[e]; }
      DEB_T_END
      DEB_T_BAR();
      return;
    }

    // ================= error norm, PID controller (diffrax semantics) =================
    {
      const double ik2 = DEB_RCP(k2);
      // components 0, 2, 3, 5, 6, 7 with weights 1, k^2, 1, 1, 1/k^2, 1 (perturbations.py:701-724): one lane each (every
      // warp of the team does the same), summed in the reference's order so that all threads hold the same bits
      DEB_REGS(double, esq, );
      DEB_LANES_BEGIN
        DEB_USE(esq);
        esq = 0.0;
        if (lane < 6) {
          const int q = lane, e = q == 0 ? 0 : (q < 3 ? q + 1 : q + 2);
          const double w = q == 1 ? k2 : (q == 4 ? ik2 : 1.0);
          double y0v = W.y()[e], y1v = anynan ? y0v : W.u()[e], ev = W.r()[e];
          if (ev != ev) ev = INFINITY;
          const double sc = ev / (P.atol + fmax(fabs(y0v), fabs(y1v)) * P.rtol) * w;
          esq = sc * sc;
        }
      DEB_LANES_END
#pragma unroll
      for (int q = 0; q < 6; ++q) errnorm2 += DEB_SHFL(esq, q);
    }
    DEB_T_BAR();          // y, u, r are rewritten below (output sampling, accepted state)
    const double E = sqrt(errnorm2 / 6.0);
    const bool keep = (P.mode == 3) ? (DEB_LDG(P.rp_keep + (size_t)mode * P.rp_stride + nsteps) != 0) : (E < 1.0);
    double inv = 1.0 / E;
    const bool inv_regular = inv > 0.0 && !isinf(inv);
    const double linv = inv_regular ? DEB_LOG(inv) : 0.0;        // (inv = 0 or inf is reset to 1 below: log = 0)
    double f1 = P.c1 != 0.0 ? (inv_regular ? DEB_EXP(P.c1 * linv) : DEB_POW(inv, P.c1)) : 1.0;
    double f2 = P.c2 != 0.0 ? DEB_EXP(P.c2 * linv_prev) : 1.0;        // linv_prev = log(inv_prev), kept from the step that set it
    double f3 = P.c3 != 0.0 ? DEB_EXP(P.c3 * DEB_LOG(inv_pprev)) : 1.0;
    double fac = fmin(fmax(P.safety * f1 * f2 * f3, keep ? 1.0 : P.factormin), P.factormax);
    if (!(fac == fac)) fac = NAN;
    const double dtn = dt * fac;
    if (inv == 0.0 || isinf(inv)) inv = 1.0;
    ++nsteps;
    if (keep) {
      ++nacc;
      // SaveAt(ts): linear interpolation inside the accepted step
      while (save_idx < P.nout && DEB_LDG(tout + save_idx) <= tnext) {
        const double tt = DEB_LDG(tout + save_idx);
        const double coeff = (tnext == t) ? 0.0 : (tt - t) / (tnext - t);
        const size_t obase = ((size_t)mode * P.nout + save_idx);
        if (P.return_full) {
          DEB_T_BEGIN
            for (int e = tid; e < n; e += NT) P.y_out[obase * n + e] = W.y()[e] + coeff * (W.u()[e] - W.y()[e]);
          DEB_T_END
        } else {
          DEB_T_BEGIN
            for (int e = tid; e < n; e += NT) W.r()[e] = W.y()[e] + coeff * (W.u()[e] - W.y()[e]);
          DEB_T_END
          DEB_T_BAR();
          DEB_T_BEGIN
            if (tid == 0) {
              double o20[20];
              team_convert_outputs(P, c, nb, W.r(), k, o20);
              const bool has_pk = P.pk_out && P.power_idx >= 0;
              double pkv = 0.0;
              if (has_pk) { const double yv = o20[P.power_idx]; pkv = 2.0 * 9.869604401089358 * c.As * DEB_POW(k / c.kp, c.ns - 1.0) * DEB_POW(k, -3.0) * yv * yv; }
              store_fields(P, mode, save_idx, o20, has_pk, pkv);
            }
          DEB_T_END
          DEB_T_BAR();
        }
        ++save_idx;
      }
      DEB_T_BEGIN
        DEB_FOR_TEAM(W.y()[e] = W.u()[e];)
      DEB_T_END
      inv_pprev = inv_prev; inv_prev = inv;
      linv_prev = linv;
      t = fmin(tnext, t1);
      jcur ^= 1;                 // the background posted during stage 8 is the accepted state's
    }
    DEB_T_BAR();          // the accepted state is visible before the next Jacobian evaluation
    double tn = t + dtn;
    if (tn > t1 - 1e-10) {
      if (keep) tn = t1; else tn = t + 0.5 * (t1 - t);
    }
    tnext = tn;
    if (!(tnext == tnext) || isinf(tnext)) status = 2;
    DEB_TICK(9);
#if defined(DEB_TEAM_TIMING) && !defined(DEB_CPU_EMU)
    if (tid == 0 && kidx == P.nk - 1) atomicAdd((unsigned long long*)&g_team_timing[15], 1ull);
#endif
  }
  if (status == 0 && t < t1) status = 1;
  DEB_T_BEGIN
    if (tid == 0) {
      store_status(P, mode, status, nsteps, nacc);
    }
  DEB_T_END
}

}  // namespace deb
