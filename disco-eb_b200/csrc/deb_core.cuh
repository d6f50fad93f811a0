// deb_core.cuh -- per-mode Einstein-Boltzmann integrator, one WARP per k-mode.
//
// The algorithm is the reference's (Rodas5Transformed + diffrax PID loop, see include/discoeb_b200.h
// for the file:line map) but organised for a Blackwell SM:
//   * 32 lanes own the n state variables cyclically (element e -> lane e%32); the 7 stage vectors
//     k_1..k_7 live in registers, the stage state / right-hand side in shared memory;
//   * the Jacobian is never formed: W = I/(gamma dt) - J is split into
//       - the scale-factor column (row 0 is closed, so x_0 is solved first and moved to the rhs),
//       - 3+nq tri-diagonal hierarchy tails (l >= 3) eliminated by a pivot-free continued-fraction
//         sweep whose Schur complement touches one diagonal entry of the head per chain,
//       - a dense "head" (metric, fluids, l <= 2 of every hierarchy, a h') of 17+3nq <= 32 unknowns
//         factored by LU with partial pivoting, one lane per row;
//   * d f/d a comes from a forward-dual evaluation of the same RHS code (what jax.jacfwd does).
//
// This header compiles for the device (nvcc) and, with -DDEB_CPU_EMU, as plain C++ in which the
// 32 lanes are executed by loops.  The emulation build is TEST INFRASTRUCTURE only (tests/emu):
// it lets the CPU test-suite run the very same source against the oracle; the Python package
// never loads it.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#ifdef DEB_CPU_EMU
#define DEB_DEV inline
#define DEB_HD inline
#define DEB_LDG(p) (*(p))
#define DEB_LANES_BEGIN for (int lane = 0; lane < 32; ++lane) {
#define DEB_LANES_END }
#define DEB_LANE0_BEGIN {
#define DEB_LANE0_END }
#define DEB_SYNC()
#define DEB_REGS(type, name, dims) type name##_all[32] dims
#define DEB_USE(name) auto& name = name##_all[lane]
#define DEB_SHFL(name, src) (name##_all[(src)])
#define DEB_ANY(name) ([&]() { int a_ = 0; for (int l_ = 0; l_ < 32; ++l_) a_ |= (name##_all[l_] != 0); return a_; }())
// lane holding the largest 32-bit key (lowest lane among ties)
#define DEB_ARGMAX_U32(name) ([&]() { unsigned m_ = name##_all[0]; int a_ = 0; for (int l_ = 1; l_ < 32; ++l_) if (name##_all[l_] > m_) { m_ = name##_all[l_]; a_ = l_; } return a_; }())
#define DEB_RSQRT(x) (1.0 / sqrt(x))
#define DEB_RCP(x) (1.0 / (x))
#define DEB_BAR_ARRIVE(id)
#define DEB_BAR_SYNC(id)
// argmax restricted to the lanes of `mask` (every lane passes the mask of its own group)
#define DEB_ARGMAX_U32_IN(name, mask) ([&]() { unsigned m_ = 0; int a_ = -1; for (int l_ = 0; l_ < 32; ++l_) if (((mask) >> l_) & 1u) { if (a_ < 0 || name##_all[l_] > m_) { m_ = name##_all[l_]; a_ = l_; } } return a_; }())
#define DEB_WARP_SUM(name) ([&]() { double s_ = 0.0; for (int l_ = 0; l_ < 32; ++l_) s_ += name##_all[l_]; return s_; }())
#else
#define DEB_DEV __device__ __forceinline__
#define DEB_HD __host__ __device__ __forceinline__
#ifndef DEB_LDG
#define DEB_LDG(p) __ldg(p)
#endif
#define DEB_LANES_BEGIN {
#define DEB_LANES_END } __syncwarp();
#define DEB_LANE0_BEGIN if (lane == 0) {
#define DEB_LANE0_END } __syncwarp();
#define DEB_SYNC() __syncwarp()
#define DEB_REGS(type, name, dims) type name dims
#define DEB_USE(name)
#define DEB_SHFL(name, src) __shfl_sync(0xffffffffu, name, (src))
#define DEB_ANY(name) __any_sync(0xffffffffu, name)
#define DEB_ARGMAX_U32(name) (__ffs(__ballot_sync(0xffffffffu, (name) == __reduce_max_sync(0xffffffffu, (name)))) - 1)
#define DEB_RSQRT(x) rsqrt(x)
#define DEB_RCP(x) __drcp_rn(x)      // correctly rounded reciprocal: same bits as 1.0/x, shorter sequence
// named barriers of the two-warp (main + helper) variant: 64 threads, producer arrives, consumer syncs
#define DEB_BAR_ARRIVE(id) asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory")
#define DEB_BAR_SYNC(id) asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory")
#define DEB_ARGMAX_U32_IN(name, mask) (__ffs(__ballot_sync(0xffffffffu, (name) == __reduce_max_sync((mask), (name))) & (mask)) - 1)
#define DEB_WARP_SUM(name) deb::warp_sum(name)
#endif

// exp/log/pow go through these names so that a translation unit can route them to out-of-line copies
// (deb_team.cu: the step loop must fit the instruction cache; each inlined exp/log is 1-2 KB, pow 5 KB)
#ifndef DEB_EXP
#define DEB_EXP(x) exp(x)
#define DEB_LOG(x) log(x)
#define DEB_POW(x, y) pow(x, y)
#endif

namespace deb {

constexpr int NSCAL = 24;
// learned work list (team kernel): header magic and the largest launch it is kept for
#define DEB_ORDER_MAGIC 0x0EB0DE55
#define DEB_ORDER_MAX 2048
// status[] before a mode has been integrated (include/discoeb_b200.h: 0 ok, 1 max_steps, 2 non-finite, 3 not processed)
#define DEB_STATUS_UNPROCESSED 3
enum { S_OMEGAM = 0, S_OMEGAB, S_OMEGADE, S_OMEGAK, S_GRHOM, S_GRHOG, S_GRHOR, S_NEFF, S_NMNU, S_AMNU,
       S_W0, S_WA, S_CS2DE, S_YHE, S_H0, S_TAUMIN, S_AS, S_NS, S_KP };
enum { T_CS2A = 0, T_XE, T_LRHONU, T_LPNU, T_A_OF_TAU, T_XE_OF_TAU, T_TAU_OF_A, NSPLINE };

constexpr int NQMAX = 5;          // head must fit one lane per row: 17 + 3 nq <= 32
constexpr int NHMAX = 32;
constexpr int NCHMAX = 3 + NQMAX;
constexpr int LMAXCAP = 64;
constexpr int HOP_NT = 5;        // local couplings per head row (max: F2, G2)
constexpr int ICACHE = 20;        // doubles of the per-mode spline interval cache
constexpr int NSLOT = 12 + NQMAX; // background scalars the head rows are built from (see HeadOp)
constexpr int LDB = 9;            // head block rows: at most 8 entries (largest block), odd stride = conflict-free across lanes

// row types (element descriptors)
enum RowType : int { R_A = 0, R_AHP, R_ETA, R_DC, R_TC, R_DB, R_TB, R_F0, R_F1, R_F2, R_G0, R_G1, R_G2,
                     R_N0, R_N1, R_N2, R_P0, R_P1, R_P2, R_DQ, R_TQ, R_GEN, R_TRUNC };

// background-scalar slots of the head operator
enum Slot : int { SL_ONE = 0, SL_H, SL_OPAC, SL_PBO, SL_K2CS2, SL_WQ1, SL_DQD, SL_DQT, SL_TQD, SL_TQT, SL_K, SL_K2, SL_KV0 };
// SL_DQD = (cs2_Q - w_Q) H   SL_DQT = (1+w_Q)(cs2_Q - ca2_Q) H^2/k^2   SL_TQD = cs2_Q k^2/(1+w_Q)   SL_TQT = (1 - 3 cs2_Q) H

// Rodas5 coefficients, transformed form (ode_integrators_stiff.py:622-687)
#define RD_GAMMA 0.19
#define RD_A21 2.0
#define RD_A31 3.040894194418781
#define RD_A32 1.041747909077569
#define RD_A41 2.576417536461461
#define RD_A42 1.622083060776640
#define RD_A43 -0.9089668560264532
#define RD_A51 2.760842080225597
#define RD_A52 1.446624659844071
#define RD_A53 -0.3036980084553738
#define RD_A54 0.2877498600325443
#define RD_A61 -14.09640773051259
#define RD_A62 6.925207756232704
#define RD_A63 -41.47510893210728
#define RD_A64 2.343771018586405
#define RD_A65 24.13215229196062
#define RD_C21 -10.31323885133993
#define RD_C31 -21.04823117650003
#define RD_C32 -7.234992135176716
#define RD_C41 32.22751541853323
#define RD_C42 -4.943732386540191
#define RD_C43 19.44922031041879
#define RD_C51 -20.69865579590063
#define RD_C52 -8.816374604402768
#define RD_C53 1.260436877740897
#define RD_C54 -0.7495647613787146
#define RD_C61 -46.22004352711257
#define RD_C62 -17.49534862857472
#define RD_C63 -289.6389582892057
#define RD_C64 93.60855400400906
#define RD_C65 318.3822534212147
#define RD_C71 34.20013733472935
#define RD_C72 -14.15535402717690
#define RD_C73 57.82335640988400
#define RD_C74 25.83362985412365
#define RD_C75 1.408950972071624
#define RD_C76 -6.551835421242162
#define RD_C81 42.57076742291101
#define RD_C82 -13.80770672017997
#define RD_C83 93.98938432427124
#define RD_C84 18.77919633714503
#define RD_C85 -31.58359187223370
#define RD_C86 -6.685968952921985
#define RD_C87 -5.810979938412932
#define RD_CT2 0.38
#define RD_CT3 0.3878509998321533
#define RD_CT4 0.4839718937873840
#define RD_CT5 0.4570477008819580
#define RD_D1 0.19
#define RD_D2 -0.1823079225333714636
#define RD_D3 -0.319231832186874912
#define RD_D4 0.3449828624725343
#define RD_D5 -0.377417564392089818

#define AKTHOM_RHS 2.3038921003709498e-9   // perturbations.py:215
#define AKTHOM_START 2.3048e-9             // perturbations.py:648

// ---------------------------------------------------------------------------------------------
// forward dual number (value, d/da): the seed is the scale factor y[0]
// ---------------------------------------------------------------------------------------------
struct alignas(16) D2 { double x, y; };   // one 128-bit shared-memory access
struct Dual { double v, d; };
DEB_DEV Dual mk(double v, double d) { Dual r; r.v = v; r.d = d; return r; }
DEB_DEV Dual operator+(Dual a, Dual b) { return mk(a.v + b.v, a.d + b.d); }
DEB_DEV Dual operator-(Dual a, Dual b) { return mk(a.v - b.v, a.d - b.d); }
DEB_DEV Dual operator*(Dual a, Dual b) { return mk(a.v * b.v, a.d * b.v + a.v * b.d); }
DEB_DEV Dual operator/(Dual a, Dual b) { double q = a.v / b.v; return mk(q, (a.d - q * b.d) / b.v); }
DEB_DEV Dual operator+(Dual a, double b) { return mk(a.v + b, a.d); }
DEB_DEV Dual operator+(double b, Dual a) { return mk(a.v + b, a.d); }
DEB_DEV Dual operator-(Dual a, double b) { return mk(a.v - b, a.d); }
DEB_DEV Dual operator-(double b, Dual a) { return mk(b - a.v, -a.d); }
DEB_DEV Dual operator-(Dual a) { return mk(-a.v, -a.d); }
DEB_DEV Dual operator*(Dual a, double b) { return mk(a.v * b, a.d * b); }
DEB_DEV Dual operator*(double b, Dual a) { return mk(a.v * b, a.d * b); }
DEB_DEV Dual operator/(Dual a, double b) { return mk(a.v / b, a.d / b); }
DEB_DEV Dual operator/(double b, Dual a) { double ia = DEB_RCP(a.v), q = b * ia; return mk(q, -q * a.d * ia); }
DEB_DEV Dual dsqrt(Dual a) { double s = sqrt(a.v); return mk(s, 0.5 * a.d / s); }
DEB_DEV Dual dexp(Dual a) { double e = DEB_EXP(a.v); return mk(e, e * a.d); }
DEB_DEV Dual dlog(Dual a) { return mk(DEB_LOG(a.v), a.d / a.v); }
DEB_DEV double dsqrt(double a) { return sqrt(a); }
DEB_DEV double dexp(double a) { return DEB_EXP(a); }
DEB_DEV double dlog(double a) { return DEB_LOG(a); }
DEB_DEV double val(double a) { return a; }
DEB_DEV double val(Dual a) { return a.v; }
DEB_DEV double der(double) { return 0.0; }
DEB_DEV double der(Dual a) { return a.d; }
// upper 32 bits of |x|: a monotone integer key for pivot selection (ties within 2^-20 are equivalent pivots)
DEB_DEV unsigned hi32abs(double x) {
#ifdef DEB_CPU_EMU
  unsigned long long b; double ax = fabs(x); memcpy(&b, &ax, 8); return (unsigned)(b >> 32);
#else
  return (unsigned)__double2hiint(fabs(x));
#endif
}
#ifndef DEB_CPU_EMU
DEB_DEV double warp_sum(double v) {      // xor butterfly: every lane ends with the same bits
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif
DEB_DEV Dual drsqrt(Dual a) { double r = DEB_RSQRT(a.v); return mk(r, -0.5 * r * a.d / a.v); }
DEB_DEV double drsqrt(double a) { return DEB_RSQRT(a); }

// ---------------------------------------------------------------------------------------------
// natural cubic spline lookup (spline_interpolation.py:130-153)
// ---------------------------------------------------------------------------------------------
struct Spl { const double* x; const double* y; const double* S; int n; };

// idx = clip(searchsorted_left(x, xn) - 1, 0, n-2); `hint` (>=0) makes it a local walk
DEB_DEV int spl_locate(const double* x, int n, double xn, int hint) {
  int i;
  if (hint < 0) {
    int lo = 0, hi = n;               // first index with x[idx] >= xn
    while (lo < hi) { int mid = (lo + hi) >> 1; if (DEB_LDG(x + mid) < xn) lo = mid + 1; else hi = mid; }
    i = lo - 1;
    if (i < 0) i = 0;
    if (i > n - 2) i = n - 2;
    return i;
  }
  i = hint;
  if (i > n - 2) i = n - 2;
  while (i < n - 2 && DEB_LDG(x + i + 1) < xn) ++i;
  while (i > 0 && DEB_LDG(x + i) >= xn) --i;
  return i;
}

DEB_DEV double spl_eval_at(const Spl& s, int i, double xn) {
  double x0 = DEB_LDG(s.x + i), x1 = DEB_LDG(s.x + i + 1);
  double h = x1 - x0;
  double t = (xn - x0) / h;
  double A = 1.0 - t, B = t;
  return A * DEB_LDG(s.y + i) + B * DEB_LDG(s.y + i + 1)
       + ((A * A * A - A) * DEB_LDG(s.S + i) + (B * B * B - B) * DEB_LDG(s.S + i + 1)) * (h * h) / 6.0;
}
DEB_DEV Dual spl_eval_at(const Spl& s, int i, Dual xn) {
  double x0 = DEB_LDG(s.x + i), x1 = DEB_LDG(s.x + i + 1);
  double h = x1 - x0;
  double t = (xn.v - x0) / h;
  double A = 1.0 - t, B = t;
  double y0 = DEB_LDG(s.y + i), y1 = DEB_LDG(s.y + i + 1), S0 = DEB_LDG(s.S + i), S1 = DEB_LDG(s.S + i + 1);
  double v = A * y0 + B * y1 + ((A * A * A - A) * S0 + (B * B * B - B) * S1) * (h * h) / 6.0;
  double dvdt = (y1 - y0) + (-(3.0 * A * A - 1.0) * S0 + (3.0 * B * B - 1.0) * S1) * (h * h) / 6.0;
  return mk(v, dvdt / h * xn.d);
}
DEB_DEV double spl_eval(const Spl& s, double xn) { return spl_eval_at(s, spl_locate(s.x, s.n, xn, -1), xn); }

// ---------------------------------------------------------------------------------------------
// launch-wide problem description
// ---------------------------------------------------------------------------------------------
struct Problem {
  int ncosmo, nk, nout, n, nh, nch, np;          // np: n rounded up to even (smem vector stride)
  int lmaxg, lmaxgp, lmaxr, lmaxnu, nq, nth, nnu;
  int max_steps, return_full, k_per_cosmo, power_idx;
  int ig, igp, ir, iq0;
  double rtol, atol, c1, c2, c3, factormax, factormin, safety;
  const double* scalars; const double* tables; const double* kmodes; const double* aexp_out;
  const double* tau_out;                         // [ncosmo, nout] (filled by the tau_out pre-kernel)
  const double* lt_small;                        // [ncosmo] log tau of the k-independent start-time root (pre-kernel)
  double* y_out; double* pk_out;
  int* status; int* nsteps; int* naccept;
  // debug single-step mode
  const double* dbg_t0; const double* dbg_t1; const double* dbg_y0; double* dbg_y1; double* dbg_err;
  double* dbg_tau_start; double* dbg_ics;
  // debug replay mode: follow a prescribed step sequence (oracle trace) instead of the controller
  const double* rp_tnext; const int* rp_keep; const int* rp_n; int rp_stride;
  int mode;                                      // 0 evolve, 1 single step, 2 prologue only, 3 replay
  unsigned int* ticket;                          // work-queue counter
  int batch_size;                                // > 0: shared-step batches of this many consecutive modes (Rodas5Batched)
  // forward tangents (deb_tangent.cuh): ntan directions, seeds laid out like the primal inputs with a leading [ntan]
  int ntan;
  const double* d_scalars; const double* d_tables;
  const double* dtau_out;                        // [ntan, ncosmo, nout] (pre-kernel)
  double* dy_out; double* dpk_out;               // [ntan, ncosmo, nk, nout, 20|n], [ntan, ncosmo, nk, nout]
  const double* d_kmodes;                        // tangent of the wavenumbers [ntan, nk] / [ntan, ncosmo, nk], or NULL (zero)
  const double* rp_dtnext;                       // replay: tangent of the prescribed step ends [ntan, ncosmo*nk, rp_stride]
  const double* dbg_dt0; const double* dbg_dt1; const double* dbg_dy0; double* dbg_dy1;      // single-step mode
  // cost-ordered work list left in the caller's workspace by the previous launch of the same shape (team kernel):
  // header {magic, nmodes, shape hash, ...} + mode ids by descending step count; NULL when the workspace is too small
  int* order_hdr; int shape_hash;
  // sharded launch with peer stores (deb_evolve_sharded_peer_f64): a rank integrates every `out_mul`-th k-mode and writes
  // the 20 fields, P(k) and the status words of local mode kidx to row kidx*out_mul + out_add (of out_nk) of the FULL-SIZE
  // buffers of all npeer ranks -- its own and, through NVLink peer mappings, everybody else's: the epilogue is the gather
  // hybrid launch (small launches in steady state, deb_kernels.cu: launch_evolve): the team kernel takes the `hybrid_split`
  // longest modes of the learned work list (one CTA per SM), the chain-lane kernel the rest on the warps that fit beside;
  // both look at the list's verdict word (ticket[2]) and fall back / exit when the list was rejected
  int hybrid_split;
  unsigned int* ticket2;     // the lane kernel's own queue counter in hybrid launches
  int duo_nosolo;            // two-team CTAs: 1 = every long mode keeps a partner team (measurement switch DEB_DUO_SOLO=0)
  int lockstep;              // chain-lane kernel: warps of a CTA advance stage by stage together (deb_lane.cuh)
  int stage_tables;          // chain-lane kernel, ncosmo == 1: the three RHS splines staged in shared memory by one bulk async copy
  int npeer, out_mul, out_add, out_nk;
  double* y_peer[8]; double* pk_peer[8]; int* st_peer[8]; int* ns_peer[8];
};

// epilogue stores: 20 output fields (+ P(k)) of (mode, save_idx), and the per-mode status words
DEB_DEV void store_fields(const Problem& P, int mode, int save_idx, const double* o20, bool has_pk, double pk) {
  if (P.npeer == 0) {
    const size_t obase = (size_t)mode * P.nout + save_idx;
    for (int q = 0; q < 20; ++q) P.y_out[obase * 20 + q] = o20[q];
    if (has_pk) P.pk_out[obase] = pk;
    return;
  }
  const int cosmo = mode / P.nk, kidx = mode - cosmo * P.nk;
  const size_t obase = ((size_t)cosmo * P.out_nk + (size_t)kidx * P.out_mul + P.out_add) * P.nout + save_idx;
  for (int r = 0; r < P.npeer; ++r) {
    double* y = P.y_peer[r];
    for (int q = 0; q < 20; ++q) y[obase * 20 + q] = o20[q];
    if (has_pk) P.pk_peer[r][obase] = pk;
  }
}
DEB_DEV void store_status(const Problem& P, int mode, int status, int nsteps, int nacc) {
  P.status[mode] = status; P.nsteps[mode] = nsteps;
  if (P.naccept) P.naccept[mode] = nacc;
  if (P.npeer > 0) {
    const int cosmo = mode / P.nk, kidx = mode - cosmo * P.nk;
    const size_t g = (size_t)cosmo * P.out_nk + (size_t)kidx * P.out_mul + P.out_add;
    for (int r = 0; r < P.npeer; ++r) { P.st_peer[r][g] = status; P.ns_peer[r][g] = nsteps; }
  }
}

// momentum bins (background.py:27-38); weights already divided by 7 pi^4/120
struct NuBins { double q[NQMAX], w[NQMAX], dl[NQMAX]; };

// launch-constant tables placed in shared memory once per CTA
struct CtaConst {
  double cl[LMAXCAP];     // l/(2l+1)
  double ch[LMAXCAP];     // (l+1)/(2l+1)
  NuBins nu;
  int hidx[NHMAX];        // head position -> state index
  int htype[NHMAX];       // head position -> row type
  int hbin[NHMAX];        // head position -> momentum bin (P rows)
  int ch_base[NCHMAX];    // chain -> state index of its l=0 element
  int ch_stride[NCHMAX];
  int ch_lmax[NCHMAX];
  int ch_h2[NCHMAX];      // chain -> head position of its l=2 element
  int blo[NHMAX], bhi[NHMAX];   // head position -> [lo, hi) of its diagonal block (AHP: empty)
  // head operator: row r of the RHS restricted to the head is
  //   f_r = sum_t hop_c[t][r] * S[hop_slot] * y[hop_col]  +  hop_chc[r] * S[hop_chs[r]] * h'  +  hop_cec[r] * eta'
  // with S the per-evaluation background scalars (slots below); the same table assembles W_h.
  double hop_c[HOP_NT][NHMAX];
  double hop_chc[NHMAX], hop_cec[NHMAX];
  int hop_meta[HOP_NT][NHMAX];   // slot | state index << 8 | (head column + 1) << 20   (0 = unused term)
  int hop_chs[NHMAX];
  const int* tail;        // tail entry -> element | chain<<12 | ell<<16, elements ascending   [np]
  int ntail;              // number of hierarchy rows with l >= 3
};

DEB_DEV Spl get_spline(const Problem& P, int cosmo, int which) {
  size_t tl = 3 * (size_t)(5 * P.nth + 2 * P.nnu);
  const double* base = P.tables + (size_t)cosmo * tl;
  size_t off = 0;
  for (int s = 0; s < which; ++s) off += 3 * (size_t)((s == T_LRHONU || s == T_LPNU) ? P.nnu : P.nth);
  int n = (which == T_LRHONU || which == T_LPNU) ? P.nnu : P.nth;
  Spl r; r.x = base + off; r.y = r.x + n; r.S = r.y + n; r.n = n;
  return r;
}

// element descriptor: type | ell << 8 | chain << 16
DEB_DEV int elem_desc(const Problem& P, int e) {
  const int nq = P.nq;
  if (e >= P.n) return -1;
  if (e == 0) return R_A;
  if (e == 1) return R_AHP;
  if (e == 2) return R_ETA;
  if (e == 3) return R_DC;
  if (e == 4) return R_TC;
  if (e == 5) return R_DB;
  if (e == 6) return R_TB;
  if (e == P.n - 2) return R_DQ;
  if (e == P.n - 1) return R_TQ;
  int chain, l, L;
  if (e < P.igp) { chain = 0; l = e - P.ig; L = P.lmaxg; }
  else if (e < P.ir) { chain = 1; l = e - P.igp; L = P.lmaxgp; }
  else if (e < P.iq0) { chain = 2; l = e - P.ir; L = P.lmaxr; }
  else { int o = e - P.iq0; l = o / nq; chain = 3 + (o - l * nq); L = P.lmaxnu; }
  int type;
  if (l == L) type = R_TRUNC;
  else if (l >= 3) type = R_GEN;
  else if (chain == 0) type = R_F0 + l;
  else if (chain == 1) type = R_G0 + l;
  else if (chain == 2) type = R_N0 + l;
  else type = R_P0 + l;
  return type | (l << 8) | (chain << 16);
}

DEB_DEV void init_cta_const(const Problem& P, CtaConst& C, int* tail, int tid, int nthreads) {
  if (tid == 0) {
    int nt = 0;
    for (int e = 0; e < P.n; ++e) {
      const int d = elem_desc(P, e), ty = d & 0xff;
      if (ty == R_GEN || ty == R_TRUNC) tail[nt++] = e | ((d >> 16) << 12) | (((d >> 8) & 0xff) << 16);
    }
    C.ntail = nt; C.tail = tail;
  }
  for (int l = tid; l < LMAXCAP; l += nthreads) {
    C.cl[l] = (double)l / (double)(2 * l + 1);
    C.ch[l] = (double)(l + 1) / (double)(2 * l + 1);
  }
  if (tid == 0) {
    const int nq = P.nq;
    const double q3[3] = {0.913201, 3.37517, 7.79184}, k3[3] = {0.0687359, 3.31435, 2.29911};
    const double q4[4] = {0.7, 2.62814, 5.90428, 12.0}, k4[4] = {0.0200251, 1.84539, 3.52736, 0.289427};
    const double q5[5] = {0.583165, 2.0, 4.0, 7.26582, 13.0}, k5[5] = {0.0081201, 0.689407, 2.8063, 2.05156, 0.12681};
    for (int i = 0; i < nq; ++i) {
      double q = nq == 3 ? q3[i] : (nq == 4 ? q4[i] : q5[i]);
      double kw = nq == 3 ? k3[i] : (nq == 4 ? k4[i] : k5[i]);
      double dl = -q / (1.0 + DEB_EXP(-q));
      C.nu.q[i] = q; C.nu.dl[i] = dl;
      C.nu.w[i] = kw / (-0.25 * dl) / 5.682196976983475;
    }
    int h = 0;
    auto put = [&](int e, int type, int bin) { C.hidx[h] = e; C.htype[h] = type; C.hbin[h] = bin; ++h; };
    // Head order = the diagonal blocks of the local Jacobian: everything else couples only through the two
    // metric scalars h' and eta' (a rank-2 term) -- {eta} {dc,tc} {db,tb,F0-2,G0-2} {N0-2} {psi0-2}_i {dq,tq}, a h'
    int b0;
    auto blk = [&](int lo, int hi) { for (int q = lo; q < hi; ++q) { C.blo[q] = lo; C.bhi[q] = hi; } };
    for (int q = 0; q < NHMAX; ++q) { C.blo[q] = 0; C.bhi[q] = 0; }
    b0 = h; put(2, R_ETA, 0); blk(b0, h);
    b0 = h; put(3, R_DC, 0); put(4, R_TC, 0); blk(b0, h);
    b0 = h; put(5, R_DB, 0); put(6, R_TB, 0);
    for (int l = 0; l < 3; ++l) put(P.ig + l, R_F0 + l, 0);
    C.ch_h2[0] = h - 1;
    for (int l = 0; l < 3; ++l) put(P.igp + l, R_G0 + l, 0);
    C.ch_h2[1] = h - 1; blk(b0, h);
    b0 = h; for (int l = 0; l < 3; ++l) put(P.ir + l, R_N0 + l, 0);
    C.ch_h2[2] = h - 1; blk(b0, h);
    for (int i = 0; i < nq; ++i) {
      b0 = h; for (int l = 0; l < 3; ++l) put(P.iq0 + l * nq + i, R_P0 + l, i);
      C.ch_h2[3 + i] = h - 1; blk(b0, h);
    }
    b0 = h; put(P.n - 2, R_DQ, 0); put(P.n - 1, R_TQ, 0); blk(b0, h);
    put(1, R_AHP, 0);
    // ---- head operator (perturbations.py:259-369 restricted to the l <= 2 rows) ----
    for (int r = 0; r < NHMAX; ++r) {
      for (int t = 0; t < HOP_NT; ++t) { C.hop_c[t][r] = 0.0; C.hop_meta[t][r] = 0; }
      C.hop_chc[r] = 0.0; C.hop_cec[r] = 0.0; C.hop_chs[r] = SL_ONE;
    }
    auto hpos = [&](int e) { for (int q = 0; q < h; ++q) if (C.hidx[q] == e) return q; return -1; };
    for (int r = 0; r < h; ++r) {
      int nt = 0;
      auto term = [&](double cst, int slot, int e) {
        C.hop_c[nt][r] = cst; C.hop_meta[nt][r] = slot | (e << 8) | ((hpos(e) + 1) << 20); ++nt;
      };
      const int e = C.hidx[r], ty = C.htype[r], bin = C.hbin[r];
      const int ig = P.ig, igp = P.igp, ir = P.ir;
      switch (ty) {
        case R_ETA: C.hop_cec[r] = 1.0; break;
        case R_DC: term(-1.0, SL_ONE, 4); C.hop_chc[r] = -0.5; break;
        case R_TC: term(-1.0, SL_H, 4); break;
        case R_DB: term(-1.0, SL_ONE, 6); C.hop_chc[r] = -0.5; break;
        case R_TB: term(-1.0, SL_H, 6); term(1.0, SL_K2CS2, 5); term(1.0, SL_PBO, 8); term(-1.0, SL_PBO, 6); break;
        case R_F0: term(-4.0 / 3.0, SL_ONE, ig + 1); C.hop_chc[r] = -2.0 / 3.0; break;
        case R_F1: term(0.25, SL_K2, ig); term(-0.5, SL_K2, ig + 2); term(-1.0, SL_OPAC, ig + 1); term(1.0, SL_OPAC, 6); break;
        case R_F2: term(8.0 / 15.0, SL_ONE, ig + 1); term(-0.6, SL_K, ig + 3); term(-0.9, SL_OPAC, ig + 2);
                   term(0.1, SL_OPAC, igp); term(0.1, SL_OPAC, igp + 2); C.hop_chc[r] = 4.0 / 15.0; C.hop_cec[r] = 1.6; break;
        case R_G0: term(-1.0, SL_K, igp + 1); term(-0.5, SL_OPAC, igp); term(0.5, SL_OPAC, ig + 2); term(0.5, SL_OPAC, igp + 2); break;
        case R_G1: term(1.0 / 3.0, SL_K, igp); term(-2.0 / 3.0, SL_K, igp + 2); term(-1.0, SL_OPAC, igp + 1); break;
        case R_G2: term(0.4, SL_K, igp + 1); term(-0.6, SL_K, igp + 3); term(-0.9, SL_OPAC, igp + 2);
                   term(0.1, SL_OPAC, ig + 2); term(0.1, SL_OPAC, igp); break;
        case R_N0: term(-4.0 / 3.0, SL_ONE, ir + 1); C.hop_chc[r] = -2.0 / 3.0; break;
        case R_N1: term(0.25, SL_K2, ir); term(-0.5, SL_K2, ir + 2); break;
        case R_N2: term(8.0 / 15.0, SL_ONE, ir + 1); term(-0.6, SL_K, ir + 3); C.hop_chc[r] = 4.0 / 15.0; C.hop_cec[r] = 1.6; break;
        case R_P0: term(-1.0, SL_KV0 + bin, e + nq); C.hop_chc[r] = C.nu.dl[bin] / 6.0; break;
        case R_P1: term(1.0 / 3.0, SL_KV0 + bin, e - nq); term(-2.0 / 3.0, SL_KV0 + bin, e + nq); break;
        case R_P2: term(0.4, SL_KV0 + bin, e - nq); term(-0.6, SL_KV0 + bin, e + nq);
                   C.hop_chc[r] = -C.nu.dl[bin] / 15.0; C.hop_cec[r] = -0.4 * C.nu.dl[bin]; break;
        case R_DQ: term(-1.0, SL_WQ1, P.n - 1); term(-3.0, SL_DQD, P.n - 2); term(-9.0, SL_DQT, P.n - 1);
                   C.hop_chc[r] = -0.5; C.hop_chs[r] = SL_WQ1; break;
        case R_TQ: term(-1.0, SL_TQT, P.n - 1); term(1.0, SL_TQD, P.n - 2); break;
        default: break;      // R_AHP: f = mt.f1, dense Jacobian row j1
      }
    }
    C.ch_base[0] = P.ig;  C.ch_stride[0] = 1; C.ch_lmax[0] = P.lmaxg;
    C.ch_base[1] = P.igp; C.ch_stride[1] = 1; C.ch_lmax[1] = P.lmaxgp;
    C.ch_base[2] = P.ir;  C.ch_stride[2] = 1; C.ch_lmax[2] = P.lmaxr;
    for (int i = 0; i < nq; ++i) { C.ch_base[3 + i] = P.iq0 + i; C.ch_stride[3 + i] = nq; C.ch_lmax[3 + i] = P.lmaxnu; }
  }
}

// per-mode constants
struct Cosmo {
  double Omegam, Omegab, OmegaDE, Omegak, grhom, grhog, grhor, Neff, Nmnu, amnu, w0, wa, cs2de, YHe, H0, taumin;
  double As, ns, kp;
  double Omegac, akthom, rq_exp;
  Spl cs2a, xe, lrn, lpn, a_of_tau, xe_of_tau, tau_of_a;
};
DEB_DEV Cosmo load_cosmo(const Problem& P, int c) {
  const double* s = P.scalars + (size_t)c * NSCAL;
  Cosmo o;
  o.Omegam = DEB_LDG(s + S_OMEGAM); o.Omegab = DEB_LDG(s + S_OMEGAB); o.OmegaDE = DEB_LDG(s + S_OMEGADE);
  o.Omegak = DEB_LDG(s + S_OMEGAK); o.grhom = DEB_LDG(s + S_GRHOM); o.grhog = DEB_LDG(s + S_GRHOG);
  o.grhor = DEB_LDG(s + S_GRHOR); o.Neff = DEB_LDG(s + S_NEFF); o.Nmnu = DEB_LDG(s + S_NMNU);
  o.amnu = DEB_LDG(s + S_AMNU); o.w0 = DEB_LDG(s + S_W0); o.wa = DEB_LDG(s + S_WA); o.cs2de = DEB_LDG(s + S_CS2DE);
  o.YHe = DEB_LDG(s + S_YHE); o.H0 = DEB_LDG(s + S_H0); o.taumin = DEB_LDG(s + S_TAUMIN);
  o.As = DEB_LDG(s + S_AS); o.ns = DEB_LDG(s + S_NS); o.kp = DEB_LDG(s + S_KP);
  o.Omegac = o.Omegam - o.Omegab;
  o.akthom = AKTHOM_RHS * (1.0 - o.YHe) * o.Omegab * o.H0 * o.H0;
  o.rq_exp = -3.0 * (1.0 + o.w0 + o.wa);
  o.cs2a = get_spline(P, c, T_CS2A); o.xe = get_spline(P, c, T_XE); o.lrn = get_spline(P, c, T_LRHONU);
  o.lpn = get_spline(P, c, T_LPNU); o.a_of_tau = get_spline(P, c, T_A_OF_TAU);
  o.xe_of_tau = get_spline(P, c, T_XE_OF_TAU); o.tau_of_a = get_spline(P, c, T_TAU_OF_A);
  return o;
}

// ---------------------------------------------------------------------------------------------
// per-warp shared-memory workspace (carved out of dynamic shared memory by the kernel)
// ---------------------------------------------------------------------------------------------
struct WarpWs {
  double* y_;   // accepted state at tprev [np]
  double* u_;   // stage state [np]
  double* r_;   // stage right-hand side, solved in place -> k_i [np]
  double* m_;   // tail backward multipliers W_{l,l+1}/e_{l+1} [np]
  double* ie_;   // tail inverse pivots 1/e_l [np]
  double* g_;   // tail forward multipliers -W_{l,l-1}/e_l [np]
  double* ja_;   // d f / d a at (t0, y0) [np]
  double* lu_;   // head matrix -> its inverse [NHMAX*LDB], row r holds its block's columns lo..hi-1
  double* gh_;   // d h'/d y_c over head columns [NHMAX]
  double* ge_;   // d eta'/d y_c [NHMAX]
  double* j1_;   // d f_1/d y_c (the a h' row) [NHMAX]
  double* qh_;   // D^-1 (coefficient of h' per row) [NHMAX]
  double* qe_;   // D^-1 (coefficient of eta' per row) [NHMAX]
  double* xb_;   // gather buffer of the head solve [NHMAX + 8], the last 8 stay zero
  double* kc_;   // chain wavenumber k or k v_i (value, d/da) [2*NCHMAX]
  double* kap_;   // chain damping opac or 0 (value, d/da) [2*NCHMAX]
  double* nur_;   // w_i psi0_i / v_i (value, d/da) [2*NQMAX]
  double* nup_;   // w_i psi0_i v_i (value, d/da) [2*NQMAX]
  double* sl_;   // background-scalar slots of the head operator (value, d/da) [2*NSLOT]
  double* ic_;   // spline interval cache [ICACHE]
  int* perm_;    // pivot row of elimination step j [NHMAX]
  Cosmo* cosmo_; // per-mode constants and table pointers
  DEB_DEV double* y() const { return y_; }
  DEB_DEV double* u() const { return u_; }
  DEB_DEV double* r() const { return r_; }
  DEB_DEV double* m() const { return m_; }
  DEB_DEV double* ie() const { return ie_; }
  DEB_DEV double* g() const { return g_; }
  DEB_DEV double* ja() const { return ja_; }
  DEB_DEV double* lu() const { return lu_; }
  DEB_DEV double* gh() const { return gh_; }
  DEB_DEV double* ge() const { return ge_; }
  DEB_DEV double* j1() const { return j1_; }
  DEB_DEV double* qh() const { return qh_; }
  DEB_DEV double* qe() const { return qe_; }
  DEB_DEV double* xb() const { return xb_; }
  DEB_DEV double* kc() const { return kc_; }
  DEB_DEV double* kap() const { return kap_; }
  DEB_DEV double* nur() const { return nur_; }
  DEB_DEV double* nup() const { return nup_; }
  DEB_DEV double* sl() const { return sl_; }
  DEB_DEV double* ic() const { return ic_; }
  DEB_DEV int* perm() const { return perm_; }
  DEB_DEV Cosmo* cosmo() const { return cosmo_; }
};
DEB_HD size_t warp_ws_doubles(int np) {
  return (size_t)7 * np + NHMAX * LDB + 6 * NHMAX + 8 + 4 * NCHMAX + 4 * NQMAX + 2 * NSLOT + ICACHE + NHMAX / 2 + 2 + (sizeof(Cosmo) + 7) / 8;
}
DEB_DEV void carve(WarpWs& W, double* base, int np) {
  W.y_ = base; W.u_ = W.y_ + np; W.r_ = W.u_ + np; W.m_ = W.r_ + np; W.ie_ = W.m_ + np; W.g_ = W.ie_ + np; W.ja_ = W.g_ + np;
  W.lu_ = W.ja_ + np; W.gh_ = W.lu_ + NHMAX * LDB; W.ge_ = W.gh_ + NHMAX; W.j1_ = W.ge_ + NHMAX;
  W.qh_ = W.j1_ + NHMAX; W.qe_ = W.qh_ + NHMAX; W.xb_ = W.qe_ + NHMAX;
  W.kc_ = W.xb_ + NHMAX + 8; W.kap_ = W.kc_ + 2 * NCHMAX; W.nur_ = W.kap_ + 2 * NCHMAX; W.nup_ = W.nur_ + 2 * NQMAX;
  W.sl_ = W.nup_ + 2 * NQMAX;
  W.ic_ = W.sl_ + 2 * NSLOT;
  W.perm_ = (int*)(W.ic_ + ICACHE);
  W.cosmo_ = (Cosmo*)(W.ic_ + ICACHE + NHMAX / 2 + 2);
}

// row r of the block-diagonal head matrix, addressable by ABSOLUTE head column lo <= c < hi
DEB_DEV double* hrow(const WarpWs& W, int r, int lo) { return W.lu() + (r * LDB - lo); }

// Gauss-Jordan with partial pivoting on all diagonal blocks of the head at once (lane = row, <= 8 columns per block), used
// by the team and chain-lane kernels (expects W, C, nhb, lane and the lane registers pcol, rscale, pkey, pivl, fmul in scope; the
// rows were assembled in shared memory, hrow).  The rows live in REGISTERS for the eight pivot steps and the pivot row
// travels by shuffles: a step is one dependency chain (key, 8-way search, reciprocal, update) without the load - store -
// load round trips through shared memory between steps (team kernel: 10.8 k -> 8.3 k cycles per Jacobian).  Same
// operations on the same operands as the in-memory form: the factors are bit-identical.  The pivot row is left unscaled
// until the end; afterwards S[:, j] is column perm[j] of the accumulated row operations: x_j = (S b_perm)[perm[j]].
#define DEB_GJ_ROWP(r) (hrow(W, (r), C.blo[(r)]) + C.blo[(r)])      /* first of the 8 slots of row r (deb_lane.cuh redefines both) */
#define DEB_GJ_PERM(j) W.perm()[(j)]
#define DEB_GJ_ELIM(i, BT) if ((i) != (BT)) { const double p_ = DEB_SHFL(g##i, pivl); if (gact) g##i = g##i - fmul * p_; }
#define DEB_GJ_STEP(BT) \
      DEB_LANES_BEGIN \
        DEB_USE(pcol); DEB_USE(pkey); DEB_USE(g##BT); \
        const int lo = C.blo[lane], hi = C.bhi[lane]; \
        pkey = (lane < nhb && lo + BT < hi && pcol < 0) ? hi32abs(g##BT) + 1u : 0u; \
      DEB_LANES_END \
      DEB_LANES_BEGIN \
        DEB_USE(pcol); DEB_USE(rscale); DEB_USE(pkey); DEB_USE(pivl); DEB_USE(fmul); DEB_USE(gact); DEB_USE(g##BT); \
        const int lo = C.blo[lane], hi = C.bhi[lane], j = lo + BT; \
        /* pivot = first row of the block holding the largest key: a scan of the (at most 8) block rows by shuffles */ \
        int piv = lane; unsigned best = 0u; \
        _Pragma("unroll") \
        for (int i = 0; i < 8; ++i) { \
          const int src = (lo + i < hi) ? lo + i : lane; \
          const unsigned kk = DEB_SHFL(pkey, src); \
          if (lo + i < hi && kk > best) { best = kk; piv = lo + i; } \
        } \
        pivl = piv; fmul = 0.0; gact = 0; \
        const double pv = DEB_SHFL(g##BT, piv); \
        if (lane < nhb && j < hi) { \
          const double ipv = DEB_RCP(pv); \
          if (lane == piv) { pcol = j; rscale = ipv; DEB_GJ_PERM(j) = piv; } \
          else { fmul = g##BT * ipv; gact = 1; } \
        } \
      DEB_LANES_END \
      DEB_LANES_BEGIN \
        DEB_USE(pivl); DEB_USE(fmul); DEB_USE(gact); \
        DEB_USE(g0); DEB_USE(g1); DEB_USE(g2); DEB_USE(g3); DEB_USE(g4); DEB_USE(g5); DEB_USE(g6); DEB_USE(g7); \
        const int lo = C.blo[lane], hi = C.bhi[lane]; \
        DEB_GJ_ELIM(0, BT) DEB_GJ_ELIM(1, BT) DEB_GJ_ELIM(2, BT) DEB_GJ_ELIM(3, BT) \
        DEB_GJ_ELIM(4, BT) DEB_GJ_ELIM(5, BT) DEB_GJ_ELIM(6, BT) DEB_GJ_ELIM(7, BT) \
        if (lane < nhb && lo + BT < hi) { if (gact) g##BT = -fmul; else if (lane == pivl) g##BT = 1.0; } \
      DEB_LANES_END
#define DEB_GJ_ALL(cond) \
      DEB_REGS(double, g0, ); DEB_REGS(double, g1, ); DEB_REGS(double, g2, ); DEB_REGS(double, g3, ); \
      DEB_REGS(double, g4, ); DEB_REGS(double, g5, ); DEB_REGS(double, g6, ); DEB_REGS(double, g7, ); \
      DEB_REGS(int, gact, ); \
      DEB_LANES_BEGIN \
        DEB_USE(g0); DEB_USE(g1); DEB_USE(g2); DEB_USE(g3); DEB_USE(g4); DEB_USE(g5); DEB_USE(g6); DEB_USE(g7); \
        g0 = g1 = g2 = g3 = g4 = g5 = g6 = g7 = 0.0; \
        if (lane < nhb) { \
          const double* rb = DEB_GJ_ROWP(lane); \
          g0 = rb[0]; g1 = rb[1]; g2 = rb[2]; g3 = rb[3]; g4 = rb[4]; g5 = rb[5]; g6 = rb[6]; g7 = rb[7]; \
        } \
      DEB_LANES_END \
      if (cond) { \
        DEB_GJ_STEP(0) DEB_GJ_STEP(1) DEB_GJ_STEP(2) DEB_GJ_STEP(3) DEB_GJ_STEP(4) DEB_GJ_STEP(5) DEB_GJ_STEP(6) DEB_GJ_STEP(7) \
      } \
      /* scale the rows; right-hand sides of the two Woodbury solves are gathered in pivot order below */ \
      DEB_LANES_BEGIN \
        DEB_USE(rscale); \
        DEB_USE(g0); DEB_USE(g1); DEB_USE(g2); DEB_USE(g3); DEB_USE(g4); DEB_USE(g5); DEB_USE(g6); DEB_USE(g7); \
        if (lane < nhb) { \
          double* rb = DEB_GJ_ROWP(lane); \
          rb[0] = g0 * rscale; rb[1] = g1 * rscale; rb[2] = g2 * rscale; rb[3] = g3 * rscale; \
          rb[4] = g4 * rscale; rb[5] = g5 * rscale; rb[6] = g6 * rscale; rb[7] = g7 * rscale; \
        } \
      DEB_LANES_END

// background coefficients at scale factor a (perturbations.py:176-218, background.py:110-121)
template <class T> struct Bg {
  T a, H, opac, cs2, pbo, wq1, wq, ca2;
  T gc, gb, gg, gr, gnu, gq;      // grhom Oc/a, grhom Ob/a, grhog/a^2, grhor Neff/a^2, grhor Nmnu/a^2, grhom ODE rhoQ a^2
};
struct Hints { int th, nu; };

// The RHS reads three splines of log a (cs2a and xe on one knot set, log rho_nu on another).  a(tau) is
// monotone, so the active interval changes every few dozen evaluations: its knot data, 1/h and h^2/6 are
// cached per mode in shared memory and refreshed only when log a leaves the interval.
//   thermo: [0] x0 [1] x1 [2] 1/h [3] h^2/6 [4..7] cs2a y0 y1 S0 S1 [8..11] xe y0 y1 S0 S1
//   nu    : [12] x0 [13] x1 [14] 1/h [15] h^2/6 [16..19] log rho_nu y0 y1 S0 S1
struct SplPos { double invh, A, B, cA, cB, dA, dB; };
DEB_DEV void icache_fill(double* ic, const Spl& s0, const Spl* s1, int i) {
  const double x0 = DEB_LDG(s0.x + i), x1 = DEB_LDG(s0.x + i + 1), h = x1 - x0;
  ic[0] = x0; ic[1] = x1; ic[2] = DEB_RCP(h); ic[3] = (h * h) / 6.0;
  ic[4] = DEB_LDG(s0.y + i); ic[5] = DEB_LDG(s0.y + i + 1); ic[6] = DEB_LDG(s0.S + i); ic[7] = DEB_LDG(s0.S + i + 1);
  if (s1) { ic[8] = DEB_LDG(s1->y + i); ic[9] = DEB_LDG(s1->y + i + 1); ic[10] = DEB_LDG(s1->S + i); ic[11] = DEB_LDG(s1->S + i + 1); }
}
// true when interval `i` (already located) still brackets xn under the reference's clip rule
DEB_DEV bool icache_hit(const double* ic, int i, int n, double xn) {
  return i >= 0 && (ic[0] < xn || i == 0) && (xn <= ic[1] || i == n - 2);
}
DEB_DEV SplPos spl_pos(const double* ic, double xn) {
  SplPos p;
  p.invh = ic[2];
  const double t = (xn - ic[0]) * ic[2];
  p.A = 1.0 - t; p.B = t;
  const double h26 = ic[3];
  p.cA = (p.A * p.A * p.A - p.A) * h26; p.cB = (p.B * p.B * p.B - p.B) * h26;
  p.dA = -(3.0 * p.A * p.A - 1.0) * h26; p.dB = (3.0 * p.B * p.B - 1.0) * h26;
  return p;
}
DEB_DEV double spl_at(const double* v, const SplPos& p, double) {       // v = {y0, y1, S0, S1}
  return p.A * v[0] + p.B * v[1] + (p.cA * v[2] + p.cB * v[3]);
}
DEB_DEV Dual spl_at(const double* v, const SplPos& p, Dual xn) {
  return mk(p.A * v[0] + p.B * v[1] + (p.cA * v[2] + p.cB * v[3]), ((v[1] - v[0]) + (p.dA * v[2] + p.dB * v[3])) * p.invh * xn.d);
}

// the refill runs once per few dozen evaluations: a translation unit may keep it out of line (DEB_COLD), away from
// the step loop's instruction footprint
#ifndef DEB_COLD
#define DEB_COLD DEB_DEV
#endif
DEB_COLD void icache_refill(const Cosmo& c, double lg, Hints& hint, double* ic, bool miss_th, bool miss_nu) {
  DEB_SYNC();
  if (miss_th) { hint.th = spl_locate(c.cs2a.x, c.cs2a.n, lg, hint.th); icache_fill(ic, c.cs2a, &c.xe, hint.th); }   // cs2a, xe share knots
  if (miss_nu) { hint.nu = spl_locate(c.lrn.x, c.lrn.n, lg, hint.nu); icache_fill(ic + 12, c.lrn, nullptr, hint.nu); }
  DEB_SYNC();
}
template <class T>
DEB_DEV void compute_bg(const Cosmo& c, const NuBins& nb, int nq, T a, Hints& hint, double* ic, Bg<T>& b) {
  T loga = dlog(a);
  const double lg = val(loga);
  // (all lanes take the same branch: the test reads the cache before anyone refills it; the refill writes the
  //  same values from every lane, fenced on both sides)
  const bool miss_th = !icache_hit(ic, hint.th, c.cs2a.n, lg), miss_nu = !icache_hit(ic + 12, hint.nu, c.lrn.n, lg);
  if (miss_th || miss_nu) icache_refill(c, lg, hint, ic, miss_th, miss_nu);
  const SplPos pth = spl_pos(ic, lg);
  const SplPos pnu = spl_pos(ic + 12, lg);
  T inva = 1.0 / a;
  T inva2 = inva * inva;
  b.a = a;
  b.cs2 = spl_at(ic + 4, pth, loga) * inva;
  T xe = spl_at(ic + 8, pth, loga);
  T rhonu = dexp(spl_at(ic + 16, pnu, loga));
  T rhoq = dexp(c.rq_exp * loga + 3.0 * c.wa * (a - 1.0));
  b.wq = c.w0 + c.wa * (1.0 - a);
  b.wq1 = 1.0 + b.wq;
  b.gc = (c.grhom * c.Omegac) * inva;
  b.gb = (c.grhom * c.Omegab) * inva;
  b.gg = c.grhog * inva2;
  b.gr = (c.grhor * c.Neff) * inva2;
  b.gnu = (c.grhor * c.Nmnu) * inva2;
  b.gq = (c.grhom * c.OmegaDE) * rhoq * (a * a);
  T grho = (c.grhom * c.Omegam) * inva + (c.grhog + c.grhor * (c.Neff + c.Nmnu * rhonu)) * inva2
         + b.gq + c.grhom * c.Omegak;
  b.H = dsqrt(grho * (1.0 / 3.0));
  // c_a^2 = w - w'/(3 (1+w+1e-6) H) with w' = -wa H a: the Hubble rate cancels (perturbations.py:211-212)
  b.ca2 = b.wq + (c.wa * a) / (3.0 * (b.wq1 + 1e-6));
  b.opac = xe * c.akthom * inva2;
  b.pbo = (4.0 / 3.0 * c.grhog / (c.grhom * c.Omegab)) * inva * b.opac;
}

// per-chain wavenumber k_c (k, or k v_i with v_i = 1/sqrt(1+(a amnu/q_i)^2)) and damping (opacity for the
// photon chains); executed by lane `ch` < nch.  Layout [value x NCHMAX | d/da x NCHMAX].
template <class T>
DEB_DEV void chain_coeffs_lane(const Cosmo& c, const NuBins& nb, const Bg<T>& b, double k, int ch, const double* u, int iq0,
                               double* kcA, double* kapA, double* nurA, double* nupA, double* sl) {
  T kc = 0.0 * b.a + k;
  if (ch >= 3) {
    const int i = ch - 3;
    T aq = b.a * (c.amnu / nb.q[i]);
    T s2 = 1.0 + aq * aq;
    T v = drsqrt(s2);            // v_i
    T iv = s2 * v;               // 1/v_i without a division
    kc = v * k;
    const double wp0 = nb.w[i] * u[iq0 + i];
    T tr = iv * wp0, tp = v * wp0;            // terms of drhonu and 3 dpnu (nu_perturb, perturbations.py:45-46)
    nurA[i] = val(tr); nurA[NQMAX + i] = der(tr);
    nupA[i] = val(tp); nupA[NQMAX + i] = der(tp);
    sl[SL_KV0 + i] = val(kc); sl[NSLOT + SL_KV0 + i] = der(kc);
  }
  T kp = ch < 2 ? b.opac : 0.0 * b.a;
  kcA[ch] = val(kc); kcA[NCHMAX + ch] = der(kc);
  kapA[ch] = val(kp); kapA[NCHMAX + ch] = der(kp);
}

// Mailbox of the two-warp variant: the main warp posts the scale factor of the NEXT stage as soon as it is
// known (x0 is the first thing a solve produces); the helper warp evaluates everything that depends on a alone
// while the main warp runs the sweeps of the current stage.
struct HelpBox {
  double a_req, k;
  int cmd, seq;                 // cmd: 0 = evaluate, 1 = exit;  seq: bumped per mode (helper resets its cache)
  double bgs[10];               // H, gc, gb, gg, gr, gnu, gq, wq1, ca2, wq
  double vv[2 * NQMAX];         // v_i, 1/v_i
  double ic2[ICACHE];           // the helper's own spline interval cache
  double invts, dtt;            // stage scalars for the helper's share of the tail rows
  int split;
};
enum { BAR_REQ = 1, BAR_RDY = 2, BAR_GO = 3, BAR_DONE = 4 };

// the a-only half of chain_coeffs_lane (helper warp) ...
DEB_DEV void chain_a_lane(const Cosmo& c, const NuBins& nb, const Bg<double>& b, double k, int ch, double* kcA, double* kapA,
                          double* vv, double* sl) {
  double kc = k;
  if (ch >= 3) {
    const int i = ch - 3;
    const double aq = b.a * (c.amnu / nb.q[i]);
    const double s2 = 1.0 + aq * aq;
    const double v = DEB_RSQRT(s2);
    vv[i] = v; vv[NQMAX + i] = s2 * v;
    kc = v * k;
    sl[SL_KV0 + i] = kc;
  }
  kcA[ch] = kc;
  kapA[ch] = ch < 2 ? b.opac : 0.0;
}
// ... and the half that needs the stage state (main warp)
DEB_DEV void nu_moments_lane(const NuBins& nb, const double* vv, const double* u, int iq0, int i, double* nurA, double* nupA) {
  const double wp0 = nb.w[i] * u[iq0 + i];
  nurA[i] = vv[NQMAX + i] * wp0;
  nupA[i] = vv[i] * wp0;
}

template <class T> DEB_DEV T pick(const double* arr, int i);
template <> DEB_DEV double pick<double>(const double* arr, int i) { return arr[i]; }
template <> DEB_DEV Dual pick<Dual>(const double* arr, int i) { return mk(arr[i], arr[NCHMAX + i]); }
template <class T> DEB_DEV T pickq(const double* arr, int i);
template <> DEB_DEV double pickq<double>(const double* arr, int i) { return arr[i]; }
template <> DEB_DEV Dual pickq<Dual>(const double* arr, int i) { return mk(arr[i], arr[NQMAX + i]); }

// metric sources and the three constraint quantities (perturbations.py:229-261)
template <class T> struct Metric { T hp, ep, al, f1; };

template <class T>
DEB_DEV void compute_metric(const Problem& P, const Cosmo& c, const NuBins& nb, const Bg<T>& b, const double* u,
                            double k, const double* nurA, const double* nupA, Metric<T>& mt) {
  const int nq = P.nq, iq0 = P.iq0, n = P.n;
  double eta = u[2], dc = u[3], tc = u[4], db = u[5], tb = u[6], dg = u[7], tg = u[8];
  double dr = u[P.ir], tr = u[P.ir + 1], dq = u[n - 2], tq = u[n - 1];
  T drhonu = 0.0 * b.a, dpnu = 0.0 * b.a;
  double fnu = 0.0;
  for (int i = 0; i < nq; ++i) {
    drhonu = drhonu + pickq<T>(nurA, i);
    dpnu = dpnu + pickq<T>(nupA, i);
    fnu += nb.w[i] * u[iq0 + nq + i];
  }
  const double k2 = k * k, ik2 = DEB_RCP(k2);
  T rpt = b.wq1 * b.gq * tq;
  T dgrho = b.gc * dc + b.gb * db + b.gg * dg + b.gr * dr + b.gnu * drhonu + b.gq * dq;
  // 3 dgpres: the 1/3 of the radiation and neutrino pressure cancels against the 3 of (dgrho + 3 dgpres)
  T dgpres3 = (b.gg * dg + b.gr * dr) + b.gnu * dpnu + 3.0 * (c.cs2de * (b.gq * dq))
            + (c.cs2de - b.ca2) * (9.0 * ik2 * (b.H * rpt));
  T dgtheta = b.gc * tc + b.gb * tb + 4.0 / 3.0 * (b.gg * tg + b.gr * tr) + b.gnu * (k * fnu) + rpt;
  mt.f1 = -(dgrho + dgpres3) * b.a;
  mt.hp = (2.0 * k2 * eta + dgrho) / b.H;
  mt.ep = (0.5 * ik2) * dgtheta;
  mt.al = (mt.hp + 6.0 * mt.ep) * (0.5 * ik2);
}

// background-scalar slots (lane 0) -- what the head rows and the head matrix are built from
template <class T>
DEB_DEV void fill_slots(const Cosmo& c, const Bg<T>& b, double k, double* sl) {
  const double k2 = k * k;
  T one = 0.0 * b.a + 1.0;
  T s[SL_KV0];
  s[SL_ONE] = one; s[SL_H] = b.H; s[SL_OPAC] = b.opac; s[SL_PBO] = b.pbo; s[SL_K2CS2] = k2 * b.cs2; s[SL_WQ1] = b.wq1;
  s[SL_DQD] = (c.cs2de - b.wq) * b.H;
  s[SL_DQT] = b.wq1 * (c.cs2de - b.ca2) * (b.H * b.H) * (1.0 / k2);
  s[SL_TQD] = (c.cs2de * k2) / b.wq1;
  s[SL_TQT] = (1.0 - 3.0 * c.cs2de) * b.H;
  s[SL_K] = one * k; s[SL_K2] = one * k2;
#pragma unroll
  for (int i = 0; i < SL_KV0; ++i) { sl[i] = val(s[i]); sl[NSLOT + i] = der(s[i]); }
}
template <class T> DEB_DEV T picks(const double* sl, int i);
template <> DEB_DEV double picks<double>(const double* sl, int i) { return sl[i]; }
template <> DEB_DEV Dual picks<Dual>(const double* sl, int i) { return mk(sl[i], sl[NSLOT + i]); }

// head row r of the right-hand side from the operator table (branch-free, one lane per row)
template <class T>
DEB_DEV T head_row(const CtaConst& C, const double* sl, const double* u, int r, const Metric<T>& mt) {
  T f = (C.hop_chc[r] * picks<T>(sl, C.hop_chs[r])) * mt.hp + C.hop_cec[r] * mt.ep;
#pragma unroll
  for (int t = 0; t < HOP_NT; ++t) {
    const int m = C.hop_meta[t][r];
    f = f + (C.hop_c[t][r] * u[(m >> 8) & 0xfff]) * picks<T>(sl, m & 0xff);
  }
  return C.htype[r] == R_AHP ? mt.f1 : f;
}

// hierarchy rows with l >= 3 (perturbations.py:300-303, :307-312, :323-327, :346-360), branch-free: the
// truncation row is the generic recurrence with (cl, ch) = (1, 0) and (lmax+1)/tau added to the damping
template <class T>
DEB_DEV T tail_row(const CtaConst& C, const double* kcA, const double* kapA, const double* u, int info, double invtau,
                   int* e_out, double* trunc_out) {
  const int e = info & 0xfff, chain = (info >> 12) & 0xf, l = info >> 16;
  const int s = C.ch_stride[chain], L = C.ch_lmax[chain];
  const bool isT = (l == L);
  const double clv = isT ? 1.0 : C.cl[l], chv = isT ? 0.0 : C.ch[l];
  const double tr = isT ? (double)(L + 1) : 0.0;
  const double lin = clv * u[e - s] - chv * u[isT ? e : e + s];
  T kcv = pick<T>(kcA, chain), kpv = pick<T>(kapA, chain);
  *e_out = e; *trunc_out = tr;
  return kcv * lin - (kpv + tr * invtau) * u[e];
}

// ---------------------------------------------------------------------------------------------
// prologue: start time (perturbations.py:630-681, util.py:365-396)
// ---------------------------------------------------------------------------------------------
DEB_DEV double aprimeoa_plain(const Cosmo& c, double a) {
  double loga = DEB_LOG(a);
  double rhonu = DEB_EXP(spl_eval(c.lrn, loga));
  double rhoq = DEB_EXP(c.rq_exp * loga + 3.0 * c.wa * (a - 1.0));
  double grho = c.grhom * c.Omegam / a + (c.grhog + c.grhor * (c.Neff + c.Nmnu * rhonu)) / (a * a)
              + c.grhom * c.OmegaDE * rhoq * (a * a) + c.grhom * c.Omegak;
  return sqrt(grho / 3.0);
}
DEB_DEV double cond_small_k(const Cosmo& c, double lt) {
  double tau = DEB_EXP(lt);
  double akthom = AKTHOM_START * (1.0 - c.YHe) * c.Omegab * c.H0 * c.H0;
  double xe = spl_eval(c.xe_of_tau, tau);
  double a = spl_eval(c.a_of_tau, tau);
  double opac = xe * akthom / (a * a);
  double H = aprimeoa_plain(c, a);
  return (1.0 / opac) / (1.0 / H) / 0.0004 - 1.0;
}
DEB_DEV double cond_large_k(const Cosmo& c, double lt, double k) {
  double a = spl_eval(c.a_of_tau, DEB_EXP(lt));
  return (1.0 / aprimeoa_plain(c, a)) / (1.0 / k) / 0.07 - 1.0;
}
// util.py:365-396 keeps [mid, right] when f(mid) f(left) > 0.  f(left) is carried along instead of being
// re-evaluated (same function at the same point: identical value), which halves the spline lookups.
DEB_DEV double start_small_k(const Cosmo& c) {          // k-independent root: once per cosmology
  double xl = DEB_LOG(c.taumin), xr = DEB_LOG(spl_eval(c.tau_of_a, 0.1));
  double fl = cond_small_k(c, xl);
  for (int it = 0; it < 7; ++it) {
    const double xm = 0.5 * (xl + xr), fm = cond_small_k(c, xm);
    if (fm * fl > 0) { xl = xm; fl = fm; } else xr = xm;
  }
  return 0.5 * (xl + xr);
}
DEB_DEV double start_time(const Cosmo& c, double k, double lt_small) {
  double xl = DEB_LOG(c.taumin), xr = DEB_LOG(spl_eval(c.tau_of_a, 0.1));
  double fl = cond_large_k(c, xl, k);
  for (int it = 0; it < 7; ++it) {
    const double xm = 0.5 * (xl + xr), fm = cond_large_k(c, xm, k);
    if (fm * fl > 0) { xl = xm; fl = fm; } else xr = xm;
  }
  const double lt_large = 0.5 * (xl + xr);
  if (!(lt_small == lt_small) || !(lt_large == lt_large)) return NAN;
  return DEB_EXP(fmin(lt_small, lt_large));
}

// adiabatic initial conditions (perturbations.py:526-627): value of element e
struct IcScalars { double a, deltag, thetag, deltar, thetar, shearr, deltaq, thetaq, eta; };
DEB_DEV IcScalars ic_scalars(const Cosmo& c, double tau, double k) {
  IcScalars s;
  double a = spl_eval(c.a_of_tau, tau);
  double rn = DEB_EXP(spl_eval(c.lrn, DEB_LOG(a)));
  double a2 = a * a, a4 = a2 * a2;
  double rhom = c.grhom * c.Omegam / (a2 * a);
  double rhor = (c.grhog + c.grhor * (c.Neff + c.Nmnu * rn)) / a4;
  double rhonu = c.grhor * (c.Neff + c.Nmnu * rn) / a4;
  double fracb = c.Omegab / c.Omegam, fracnu = rhonu / rhor;
  double om = a * rhom / sqrt(rhor);
  const double ci = -1.0;
  double kt = k * tau, kt2 = kt * kt;
  s.a = a;
  s.deltag = -kt2 / 3.0 * (1.0 - om * tau / 5.0) * ci;
  s.thetag = -(kt2 * kt) / tau / 36.0 * (1.0 - 3.0 * (1.0 + 5.0 * fracb - fracnu) / 20.0 / (1.0 - fracnu) * om * tau) * ci;
  s.deltar = s.deltag;
  s.thetar = -(kt2 * kt2) / tau / 36.0 / (4.0 * fracnu + 15.0)
           * (4.0 * fracnu + 11.0 + 12.0 - 3.0 * (8.0 * fracnu * fracnu + 50.0 * fracnu + 275.0) / 20.0 / (2.0 * fracnu + 15.0) * tau * om) * ci;
  s.shearr = kt2 / (45.0 + 12.0 * fracnu) * 2.0 * (1.0 + (4.0 * fracnu - 5.0) / 4.0 / (2.0 * fracnu + 15.0) * tau * om) * ci;
  double wq = c.w0 + c.wa * (1.0 - a);
  s.deltaq = kt2 / 4.0 * (1.0 + wq) * (4.0 - 3.0 * c.cs2de) / (4.0 - 6.0 * wq + 3.0 * c.cs2de) * ci;
  s.thetaq = (kt2 * kt2) / tau / 4.0 * c.cs2de / (4.0 - 6.0 * wq + 3.0 * c.cs2de) * ci;
  s.eta = ci * (1.0 - kt2 / 12.0 / (15.0 + 4.0 * fracnu)
                * (5.0 + 4.0 * fracnu - (16.0 * fracnu * fracnu + 280.0 * fracnu + 325.0) / 10.0 / (2.0 * fracnu + 15.0) * tau * om));
  return s;
}
DEB_DEV double ic_value(const Problem& P, const Cosmo& c, const NuBins& nb, const IcScalars& s, int desc, double k) {
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
      int i = chain - 3;
      double aq = s.a * c.amnu / nb.q[i];
      double v = 1.0 / sqrt(1.0 + aq * aq);
      double dl = nb.dl[i];
      if (type == R_P0) return -0.25 * dl * s.deltar;
      if (type == R_P1) return -dl * s.thetar / v / k / 3.0;
      return -0.5 * dl * s.shearr;
    }
    default: return 0.0;
  }
}

// ---------------------------------------------------------------------------------------------
// epilogue: state -> 20 output fields (perturbations.py:374-523), executed by one lane
// ---------------------------------------------------------------------------------------------
DEB_DEV void convert_outputs(const Problem& P, const Cosmo& c, const NuBins& nb, const double* y, double k, double* out) {
  const int nq = P.nq, iq0 = P.iq0, n = P.n;
  double a = y[0], eta = y[2], dc = y[3], tc = y[4], db = y[5], tb = y[6], dg = y[7], tg = y[8];
  double dr = y[P.ir], tr = y[P.ir + 1], dq = y[n - 2], tq = y[n - 1];
  double la = DEB_LOG(a);
  double rhonu = DEB_EXP(spl_eval(c.lrn, la)), pnu = DEB_EXP(spl_eval(c.lpn, la));
  double drhonu = 0.0, fnu = 0.0;
  for (int i = 0; i < nq; ++i) {
    double aq = a * c.amnu / nb.q[i];
    double v = 1.0 / sqrt(1.0 + aq * aq);
    drhonu += nb.w[i] * y[iq0 + i] / v;
    fnu += nb.w[i] * y[iq0 + nq + i];
  }
  double deltanu = drhonu / rhonu, thetanu = k * fnu / (rhonu + pnu);
  double wq = c.w0 + c.wa * (1.0 - a);
  double rhoq = DEB_POW(a, c.rq_exp) * DEB_EXP(3.0 * (a - 1.0) * c.wa);
  double a2 = a * a;
  double rpt = (1.0 + wq) * rhoq * c.grhom * c.OmegaDE * tq * a2;
  double grho = c.grhom * c.Omegam / a + (c.grhog + c.grhor * (c.Neff + c.Nmnu * rhonu)) / a2
              + c.grhom * c.OmegaDE * rhoq * a2 + c.grhom * c.Omegak;
  double H = sqrt(grho / 3.0);
  double mat = c.grhom * (c.Omegac * dc + c.Omegab * db) / a;
  double matth = c.grhom * (c.Omegac * tc + c.Omegab * tb) / a;
  double dgrho = mat + (c.grhog * dg + c.grhor * (c.Neff * dr + c.Nmnu * drhonu)) / a2 + c.grhom * c.OmegaDE * dq * rhoq * a2;
  double dgtheta = matth + 4.0 / 3.0 * (c.grhog * tg + c.Neff * c.grhor * tr) / a2 + c.Nmnu * c.grhor * k * fnu / a2 + rpt;
  double k2 = k * k;
  double hp = (2.0 * k2 * eta + dgrho) / H, ep = 0.5 * dgtheta / k2, al = (hp + 6.0 * ep) / 2.0 / k2;
  double deltam = (mat + (c.grhor * c.Nmnu * drhonu) / a2) / (c.grhom * c.Omegam / a + (c.grhor * c.Nmnu * rhonu) / a2);
  double thetam = (matth + c.Nmnu * c.grhor * k * fnu / a2) / (3.0 * (c.grhom * c.Omegam / a + c.grhor * c.Nmnu * rhonu / a2));
  double deltabc = mat / (c.grhom * c.Omegam / a);
  double thetabc = matth / (3.0 * (c.grhom * c.Omegam / a) / a2);
  thetam += al * k2; thetabc += al * k2;
  out[0] = eta; out[1] = ep; out[2] = hp; out[3] = al;
  out[4] = deltam; out[5] = thetam / H; out[6] = deltabc; out[7] = thetabc / H;
  out[8] = dc; out[9] = tc / H; out[10] = db; out[11] = tb / H; out[12] = dg; out[13] = tg / H;
  out[14] = dr; out[15] = tr / H; out[16] = deltanu; out[17] = thetanu / H; out[18] = dq; out[19] = tq / H;
}

// ---------------------------------------------------------------------------------------------
// the integrator: one call = one (cosmology, k) mode, executed by one warp
// ---------------------------------------------------------------------------------------------
#ifdef DEB_CPU_EMU
#define DEB_LANE_PARAM
#define DEB_LANE_ARG
#else
#define DEB_LANE_PARAM , const int lane
#define DEB_LANE_ARG , lane
#endif

// tail rows tt in [t0, t1) of a stage right-hand side, added onto r (two rows per trip: both evaluated before
// either store); executed by one warp
DEB_DEV void stage_tail_rows(const CtaConst& C, WarpWs& W, int t0, int t1, double invts, double dtt DEB_LANE_PARAM) {
  DEB_LANES_BEGIN
    for (int tt = t0 + lane; tt < t1; tt += 64) {
      int e0, e1 = 0; double tr0, tr1 = 0.0, f1 = 0.0, r1 = 0.0, y1 = 0.0;
      const bool two = tt + 32 < t1;
      const double f0 = tail_row<double>(C, W.kc(), W.kap(), W.u(), C.tail[tt], invts, &e0, &tr0);
      const double r0 = W.r()[e0], y0 = W.y()[e0];
      if (two) { f1 = tail_row<double>(C, W.kc(), W.kap(), W.u(), C.tail[tt + 32], invts, &e1, &tr1); r1 = W.r()[e1]; y1 = W.y()[e1]; }
      W.r()[e0] = r0 + f0 + dtt * tr0 * y0;
      if (two) W.r()[e1] = r1 + f1 + dtt * tr1 * y1;
    }
  DEB_LANES_END
}

// One evaluation by the helper warp: background at box.a_req -> chain coefficients, operator slots, and the
// scalars the main warp's metric sources need.
DEB_DEV void helper_compute(const Problem& P, const CtaConst& C, WarpWs& W, HelpBox& box, Hints& hint2 DEB_LANE_PARAM) {
  const Cosmo& c = *W.cosmo();
  const double k = box.k;
  Bg<double> b;
  compute_bg<double>(c, C.nu, P.nq, box.a_req, hint2, box.ic2, b);
  DEB_LANES_BEGIN
    if (lane < P.nch) chain_a_lane(c, C.nu, b, k, lane, W.kc(), W.kap(), box.vv, W.sl());
    if (lane == 0) {
      fill_slots<double>(c, b, k, W.sl());
      box.bgs[0] = b.H; box.bgs[1] = b.gc; box.bgs[2] = b.gb; box.bgs[3] = b.gg; box.bgs[4] = b.gr; box.bgs[5] = b.gnu;
      box.bgs[6] = b.gq; box.bgs[7] = b.wq1; box.bgs[8] = b.ca2; box.bgs[9] = b.wq;
    }
  DEB_LANES_END
}

#ifndef DEB_CPU_EMU
// helper warp of the two-warp variant: serves requests until told to exit
DEB_DEV void helper_loop(const Problem& P, const CtaConst& C, WarpWs& W, HelpBox& box, const int lane) {
  Hints hint2; hint2.th = -1; hint2.nu = -1;
  int seq = -1;
  for (;;) {
    DEB_BAR_SYNC(BAR_REQ);
    if (box.cmd == 1) break;
    if (box.seq != seq) { seq = box.seq; hint2.th = -1; hint2.nu = -1; }
    helper_compute(P, C, W, box, hint2, lane);
    __syncwarp();
    DEB_BAR_ARRIVE(BAR_RDY);
    DEB_BAR_SYNC(BAR_GO);                  // the main warp has formed the C-combinations of this stage in r
    stage_tail_rows(C, W, box.split, C.ntail, box.invts, box.dtt, lane);
    DEB_BAR_SYNC(BAR_DONE);
  }
}
#endif

// ---------------------------------------------------------------------------------------------
// Batched variant (ode_integrators_stiff.py:846-1010 Rodas5Batched, perturbations.py:786-922): the modes of a batch
// advance with ONE adaptive step size.  Every mode still runs on its own warp; what the batch shares is (i) the start
// time (the smallest of its modes') and (ii) per attempted step the error norm, an RMS over all B x 6 filtered
// components.  Both are all-to-all exchanges of one double per mode: lane 0 of every warp writes its slot, a barrier,
// every warp reads all B slots (lane i <- slots i and i + 32) and reduces them with the same butterfly, so that all warps of
// the batch hold bit-identical values and take identical accept/reject decisions.  A batch is bw warps per CTA
// x ncta CTAs of one thread-block cluster (8 warps per CTA is what the register file allows); remote slots are read
// through distributed shared memory, the barrier is the cluster's.
// ---------------------------------------------------------------------------------------------
struct BatchCtx {
  double* slots;        // this CTA's [2][bw][2] doubles (double-buffered by `parity`): value, flag
  int bw, ncta, B;      // warps per CTA, CTAs per batch, modes per batch
  int idx;              // this warp's position in the batch
  int parity;
#ifdef DEB_CPU_EMU
  double* all;          // emulation: one array [2][B][2] shared by the B threads that play the warps
#endif
};
enum { BATCH_SUM = 0, BATCH_MIN = 1 };
#ifdef DEB_CPU_EMU
inline void batch_exchange(BatchCtx& bc, double v, double f, int op, double* out_v, double* out_f) {
  double* buf = bc.all + (size_t)bc.parity * bc.B * 2;
  buf[2 * bc.idx] = v; buf[2 * bc.idx + 1] = f;
#pragma omp barrier
  // the butterfly of the device code: pairwise tree over 32 lanes (missing lanes hold the neutral element)
  double x[32], g[32];
  const double neutral = op == BATCH_MIN ? INFINITY : 0.0;
  for (int i = 0; i < 32; ++i) {          // lane i holds slots i and i + 32 (B <= 64)
    const double a = i < bc.B ? buf[2 * i] : neutral, b = i + 32 < bc.B ? buf[2 * (i + 32)] : neutral;
    x[i] = op == BATCH_MIN ? ((a != a || b != b) ? NAN : fmin(a, b)) : a + b;
    g[i] = (i < bc.B ? buf[2 * i + 1] : 0.0) + (i + 32 < bc.B ? buf[2 * (i + 32) + 1] : 0.0);
  }
  for (int o = 16; o > 0; o >>= 1)
    for (int i = 0; i < 32; ++i) if ((i & o) == 0) {
      const double a = x[i], b = x[i ^ o];
      const double r = op == BATCH_MIN ? ((a != a || b != b) ? NAN : fmin(a, b)) : a + b;
      x[i] = x[i ^ o] = r;
      const double h = g[i] + g[i ^ o]; g[i] = g[i ^ o] = h;
    }
  *out_v = x[0]; *out_f = g[0];
  bc.parity ^= 1;
}
#else
}  // namespace deb
#include <cooperative_groups.h>
namespace deb {
__device__ __forceinline__ void batch_exchange(BatchCtx& bc, double v, double f, int op, double* out_v, double* out_f) {
  namespace cg = cooperative_groups;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* mine = bc.slots + ((size_t)bc.parity * bc.bw + warp) * 2;
  if (lane == 0) { mine[0] = v; mine[1] = f; }
  if (bc.ncta > 1) cg::this_cluster().sync(); else __syncthreads();
  const double neutral = op == BATCH_MIN ? INFINITY : 0.0;
  double x = neutral, g = 0.0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {             // lane i holds slots i and i + 32 (B <= 64)
    const int slot = lane + 32 * h;
    double a = neutral, fa = 0.0;
    if (slot < bc.B) {
      const int cta = slot / bc.bw, w = slot - cta * bc.bw;
      const double* src = bc.slots + ((size_t)bc.parity * bc.bw + w) * 2;
      if (bc.ncta > 1) src = cg::this_cluster().map_shared_rank(const_cast<double*>(src), cta);
      a = src[0]; fa = src[1];
    }
    if (h == 0) { x = a; g = fa; }
    else { x = op == BATCH_MIN ? ((x != x || a != a) ? NAN : fmin(x, a)) : x + a; g += fa; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double b = __shfl_xor_sync(0xffffffffu, x, o), h = __shfl_xor_sync(0xffffffffu, g, o);
    x = op == BATCH_MIN ? ((x != x || b != b) ? NAN : fmin(x, b)) : x + b;
    g += h;
  }
  *out_v = x; *out_f = g;
  bc.parity ^= 1;       // the other buffer next time: nobody can still be reading it (every warp passed this barrier since)
}
#endif

}  // namespace deb
#include "deb_tangent.cuh"
namespace deb {

// TAN: also carry one forward tangent (direction `tan` of P.ntan) through the same step sequence; the work
// item is then (direction, cosmology, k).  TW is the tangent workspace (nullptr otherwise).
template <int NE, bool HELPER, bool TAN = false, bool BATCH = false>
DEB_DEV void integrate_mode(const Problem& P, const CtaConst& C, WarpWs& W, HelpBox* box, int mode DEB_LANE_PARAM,
                            const TanWs* TW = nullptr, int tan = 0, BatchCtx* BC = nullptr) {
  const int n = P.n, nh = P.nh, nch = P.nch, nq = P.nq;
  const int nhb = nh - 1;            // head unknowns inside diagonal blocks (the last head row is a h')
  const int cosmo = mode / P.nk, kidx = mode - cosmo * P.nk;
  const double k = DEB_LDG(P.kmodes + (P.k_per_cosmo ? (size_t)cosmo * P.nk + kidx : (size_t)kidx));
  const double k2 = k * k;
  DEB_LANE0_BEGIN
    *W.cosmo() = load_cosmo(P, cosmo);
  DEB_LANE0_END
  const Cosmo& c = *W.cosmo();
  const NuBins& nb = C.nu;
#ifdef DEB_CPU_EMU
  Hints hint2; hint2.th = -1; hint2.nu = -1;      // the emulation runs the helper inline
#endif
  if (HELPER) {
    DEB_LANE0_BEGIN
      box->k = k; box->cmd = 0; box->seq = mode;
    DEB_LANE0_END
  }
  const double* tout = P.tau_out + (size_t)cosmo * P.nout;

  DEB_REGS(int, pcol, );           // head inverse: pivot column of this lane's row
  DEB_REGS(double, rscale, );      // head inverse: 1/pivot of this lane's row
  DEB_REGS(unsigned, pkey, );
  DEB_REGS(int, pivl, );
  DEB_REGS(double, fmul, );
  DEB_REGS(double, pval, );
  DEB_REGS(double, s1, ); DEB_REGS(double, s2, ); DEB_REGS(double, s3, ); DEB_REGS(double, s4, );
  DEB_REGS(int, nanflag, );

  // ---- prologue ----
  double t1 = DEB_LDG(tout);
  double tmin_out = t1;
  for (int j = 1; j < P.nout; ++j) { double tj = DEB_LDG(tout + j); t1 = fmax(t1, tj); tmin_out = fmin(tmin_out, tj); }
  double t, tnext;
  // tangents of t, tnext, t1 and of the output times (TAN only)
  double td = 0.0, tnextd = 0.0, t1d = 0.0, tmind = 0.0;
  const size_t item = (size_t)tan * ((size_t)P.ncosmo * P.nk) + mode;
  const double* toutd = nullptr;
  Dual kD = mk(k, 0.0);
  if (TAN) {
    if (P.d_kmodes) {
      const size_t nkm = P.k_per_cosmo ? (size_t)P.ncosmo * P.nk : (size_t)P.nk;
      kD.d = DEB_LDG(P.d_kmodes + (size_t)tan * nkm + (P.k_per_cosmo ? (size_t)cosmo * P.nk + kidx : (size_t)kidx));
    }
    DEB_LANE0_BEGIN
      load_cosmo_d(P, cosmo, tan, TW->cd);
    DEB_LANE0_END
    toutd = P.dtau_out + ((size_t)tan * P.ncosmo + cosmo) * P.nout;
    int jmax = 0, jmin = 0;
    for (int j = 1; j < P.nout; ++j) {
      if (DEB_LDG(tout + j) > DEB_LDG(tout + jmax)) jmax = j;
      if (DEB_LDG(tout + j) < DEB_LDG(tout + jmin)) jmin = j;
    }
    t1d = DEB_LDG(toutd + jmax); tmind = DEB_LDG(toutd + jmin);
  }
  if (P.mode == 1) {
    t = DEB_LDG(P.dbg_t0 + mode); tnext = DEB_LDG(P.dbg_t1 + mode); t1 = tnext;
    DEB_LANES_BEGIN
      for (int e = lane; e < n; e += 32) W.y()[e] = DEB_LDG(P.dbg_y0 + (size_t)mode * n + e);
    DEB_LANES_END
    if (TAN) {
      td = DEB_LDG(P.dbg_dt0 + item); tnextd = DEB_LDG(P.dbg_dt1 + item); t1d = tnextd;
      DEB_LANES_BEGIN
        for (int e = lane; e < n; e += 32) TW->yd[e] = DEB_LDG(P.dbg_dy0 + item * n + e);
      DEB_LANES_END
    }
  } else {
    const double st0 = start_time(c, k, DEB_LDG(P.lt_small + cosmo));
    double tau_start = 0.99 * fmin(tmin_out, st0);
    if (!(st0 == st0)) tau_start = st0;
    if (BATCH) {         // the batch starts at the earliest of its modes' start times (perturbations.py:786-805)
      double fdummy;
      batch_exchange(*BC, tau_start, 0.0, BATCH_MIN, &tau_start, &fdummy);
    }
    IcScalars ics = ic_scalars(c, tau_start, k);
    DEB_LANES_BEGIN
      for (int e = lane; e < n; e += 32) W.y()[e] = ic_value(P, c, nb, ics, elem_desc(P, e), k);
    DEB_LANES_END
    if (TAN) {
      const Dual stD = start_time_d(*TW->cd, kD);
      td = 0.99 * (tmin_out <= st0 ? tmind : stD.d);
      const IcScalarsD icd = ic_scalars_d(*TW->cd, mk(tau_start, td), kD);
      DEB_LANES_BEGIN
        for (int e = lane; e < n; e += 32) TW->yd[e] = ic_value_d(P, *TW->cd, nb, icd, elem_desc(P, e), kD).d;
      DEB_LANES_END
    }
    if (P.mode == 2) {
      DEB_LANES_BEGIN
        if (lane == 0) P.dbg_tau_start[item] = TAN ? td : tau_start;
        for (int e = lane; e < n; e += 32) P.dbg_ics[item * n + e] = TAN ? TW->yd[e] : W.y()[e];
      DEB_LANES_END
      return;
    }
    t = tau_start;
    tnext = t + fmin(t / 4.0, 0.5 * (t1 - t));          // dt0 (perturbations.py:756)
    if (TAN) tnextd = td + ((t / 4.0 <= 0.5 * (t1 - t)) ? td / 4.0 : 0.5 * (t1d - td));
    if (tnext > t1 - 1e-10) { tnext = t1; tnextd = t1d; }
  }

  double inv_prev = 1.0, inv_pprev = 1.0;
  int nsteps = 0, nacc = 0, save_idx = 0, status = 0;
  // a non-finite start (NaN in taumin or in the tables) must fail like the reference's NaN-propagating jnp.minimum,
  // not be swallowed by fmin
  if (!(t == t) || !(tnext == tnext) || !(t1 == t1)) status = 2;
  Hints hint; hint.th = -1; hint.nu = -1;

  while (t < t1 && nsteps < P.max_steps && status == 0) {
    if (P.mode == 3) {
      if (nsteps >= DEB_LDG(P.rp_n + mode)) break;
      tnext = DEB_LDG(P.rp_tnext + (size_t)mode * P.rp_stride + nsteps);
      if (TAN) tnextd = DEB_LDG(P.rp_dtnext + item * P.rp_stride + nsteps);
      // the prescribed end time comes from another evaluation of tau_of_a(aexp_out): snap it onto ours
      if (fabs(tnext - t1) <= 1e-12 * fabs(t1)) { tnext = t1; tnextd = t1d; }
    }
    DEB_REGS(double, ks, [7][NE]);    // stage vectors k_1..k_7: registers, scoped to one step (dead while W is factored)
    const double dt = tnext - t;
    const double ddt = tnextd - td;
    const double invdt = DEB_RCP(dt);
    const double idg = DEB_RCP(dt * RD_GAMMA);      // diagonal of W = I/(gamma dt) - J
    const double invt0 = DEB_RCP(t);

    // ================= Jacobian pieces at (t, y) =================
    double x0piv;      // W_00 = 1/(gamma dt) - d(H a)/da
    {
      Bg<Dual> bd;
      compute_bg<Dual>(c, nb, nq, mk(W.y()[0], 1.0), hint, W.ic(), bd);
      DEB_LANES_BEGIN
        if (lane < nch) chain_coeffs_lane<Dual>(c, nb, bd, k, lane, W.y(), P.iq0, W.kc(), W.kap(), W.nur(), W.nup(), W.sl());
        if (lane == 0) fill_slots<Dual>(c, bd, k, W.sl());
      DEB_LANES_END
      Metric<Dual> md;
      compute_metric<Dual>(P, c, nb, bd, W.y(), k, W.nur(), W.nup(), md);
      // f(t,y) -> r (stage-1 right-hand side incl. dt d1 dT), d f/d a -> ja
      DEB_LANES_BEGIN
        if (lane < nh) {                      // head rows: one lane per row
          const int e = C.hidx[lane];
          Dual f = head_row<Dual>(C, W.sl(), W.y(), lane, md);
          W.r()[e] = f.v; W.ja()[e] = f.d;
        }
        if (lane == 0) { Dual f = bd.H * bd.a; W.r()[0] = f.v; W.ja()[0] = f.d; }
        const double d1t = (dt * RD_D1) * invt0 * invt0;
#pragma unroll 2
        for (int tt = lane; tt < C.ntail; tt += 32) {
          int e; double tr;
          Dual f = tail_row<Dual>(C, W.kc(), W.kap(), W.y(), C.tail[tt], invt0, &e, &tr);
          W.r()[e] = f.v + d1t * tr * W.y()[e]; W.ja()[e] = f.d;
        }
      DEB_LANES_END
      x0piv = idg - W.ja()[0];

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
            case R_P0: { const double vb = W.kc()[3 + bin] / k; wr = bd.gnu.v * nb.w[bin] / vb; wp = bd.gnu.v * nb.w[bin] * vb / 3.0; } break;
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

      // ---- tails: pivot-free backward elimination l = L .. 3 (one lane per chain) ----
      DEB_LANES_BEGIN
        if (lane < nch) {
          const int base = C.ch_base[lane], s = C.ch_stride[lane], L = C.ch_lmax[lane];
          const double kc = W.kc()[lane], kp = W.kap()[lane];
          // row L (truncation): diag = idg + kap + (L+1)/tau, lower = -kc
          double e = idg + kp + (double)(L + 1) * invt0;
          double ie = DEB_RCP(e);
          int idx = base + L * s;
          W.ie()[idx] = ie;
          W.g()[idx] = kc * ie;                      // -W_{L,L-1}/e_L
          double lower_next = -kc;                 // W_{l+1,l}
          for (int l = L - 1; l >= 2; --l) {
            idx -= s;
            double up = kc * C.ch[l];              // W_{l,l+1}
            double mm = up * ie;                   // uses 1/e_{l+1}
            W.m()[idx] = mm;
            if (l >= 3) {
              e = idg + kp - mm * lower_next;
              ie = DEB_RCP(e);
              W.ie()[idx] = ie;
              lower_next = -kc * C.cl[l];
              W.g()[idx] = -lower_next * ie;
            } else {
              W.ie()[idx] = -mm * lower_next;        // Schur increment for head diagonal (l=2 row)
            }
          }
        }
      DEB_LANES_END

      // ---- head:  W_h = D - chv gh^T - cev ge^T  (+ the a h' row).  D = I/(gamma dt) - J_local is block diagonal
      //      (blocks of size 1, 2, 8, 3, 3 x nq, 2), the metric coupling is rank 2  ->  block inverses + Woodbury.
      DEB_LANES_BEGIN
        DEB_USE(pcol); DEB_USE(rscale);
        pcol = -1; rscale = 1.0;
        if (lane < nhb) {
          const int ty = C.htype[lane], lo = C.blo[lane], hi = C.bhi[lane];
          double* row = hrow(W, lane, lo);
#pragma unroll
          for (int i = 0; i < 8; ++i) row[lo + i] = 0.0;        // all 8 slots: the unused ones stay zero for the fixed-length loops
          double diag = idg;
          if (ty == R_F2 || ty == R_G2 || ty == R_N2 || ty == R_P2) {      // Schur complement of the chain tail
            const int chain = ty == R_F2 ? 0 : (ty == R_G2 ? 1 : (ty == R_N2 ? 2 : 3 + C.hbin[lane]));
            diag += W.ie()[C.ch_base[chain] + 2 * C.ch_stride[chain]];
          }
          row[lane] = diag;
#pragma unroll
          for (int q = 0; q < HOP_NT; ++q) {             // local couplings (all inside the block)
            const int m = C.hop_meta[q][lane], hc = (m >> 20) - 1;
            if (hc >= 0) row[hc] -= C.hop_c[q][lane] * W.sl()[m & 0xff];
          }
        }
      DEB_LANES_END
    }

    // ---- block inverses: in-place Gauss-Jordan with partial pivoting inside every block, all blocks at once
    //      (step t eliminates the t-th column of each block; lanes = rows; pivot search is a masked warp max).
    //      The pivot row is left unscaled until the end, so each step is one race-free phase; afterwards
    //      S[:, j] is column perm[j] of the accumulated row operations: x_j = (S b_perm)[perm[j]].
    // (the one-warp kernels keep the rows in shared memory: they run at the register limit, and the eight row registers of
    //  DEB_GJ_ALL cost more in spills than the round trips -- n = 111: +5 %, tangent kernel: +40 %; team and lane kernels use it)
    for (int bt = 0; bt < 8; ++bt) {
      DEB_LANES_BEGIN
        DEB_USE(pcol); DEB_USE(pkey);
        const int lo = C.blo[lane], hi = C.bhi[lane];
        pkey = (lane < nhb && lo + bt < hi && pcol < 0) ? hi32abs(hrow(W, lane, lo)[lo + bt]) + 1u : 0u;
      DEB_LANES_END
      DEB_LANES_BEGIN
        DEB_USE(pcol); DEB_USE(rscale); DEB_USE(pkey); DEB_USE(pivl); DEB_USE(fmul);
        const int lo = C.blo[lane], hi = C.bhi[lane], j = lo + bt;
        // pivot = first row of the block holding the largest key: a scan of the (at most 8) block rows by shuffles
        int piv = lane; unsigned best = 0u;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int src = (lo + i < hi) ? lo + i : lane;
          const unsigned kk = DEB_SHFL(pkey, src);
          if (lo + i < hi && kk > best) { best = kk; piv = lo + i; }
        }
        pivl = piv; fmul = 0.0;
        if (lane < nhb && j < hi) {
          const double ipv = DEB_RCP(hrow(W, piv, lo)[j]);
          if (lane == piv) { pcol = j; rscale = ipv; W.perm()[j] = piv; }
          else fmul = hrow(W, lane, lo)[j] * ipv;
        }
      DEB_LANES_END
      DEB_LANES_BEGIN
        DEB_USE(pivl); DEB_USE(fmul);
        const int lo = C.blo[lane], hi = C.bhi[lane], j = lo + bt;
        if (lane < nhb && j < hi) {
          double* row = hrow(W, lane, lo);
          if (lane == pivl) row[j] = 1.0;
          else {
            const double* rb = row + lo;
            const double* pb = hrow(W, pivl, lo) + lo;
            const double f = fmul;
            double ra[8], pa[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { ra[i] = rb[i]; pa[i] = (i == bt) ? 0.0 : pb[i]; }   // (the pivot lane rewrites slot bt; unused slots hold zeros)
#pragma unroll
            for (int i = 0; i < 8; ++i) row[lo + i] = (i == bt) ? -f : ra[i] - f * pa[i];
          }
        }
      DEB_LANES_END
    }
    // scale rows; Woodbury pieces: qh = D^-1 chv, qe = D^-1 cev, capacitance C = I - [gh ge]^T [qh qe]
    DEB_LANES_BEGIN
      DEB_USE(rscale);
      if (lane < nhb) {
        double* rb = hrow(W, lane, C.blo[lane]) + C.blo[lane];
        double ra[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) ra[i] = rb[i];
#pragma unroll
        for (int i = 0; i < 8; ++i) rb[i] = ra[i] * rscale;
      }
      // right-hand sides of the two Woodbury solves, gathered in pivot order
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
          ah += row[cc] * (C.hop_chc[pr] * W.sl()[C.hop_chs[pr]]);
          ae += row[cc] * C.hop_cec[pr];
        }
        W.qh()[pcol] = ah; W.qe()[pcol] = ae;
        const double ghc = W.gh()[pcol], gec = W.ge()[pcol], j1c = W.j1()[pcol];
        s1 = ghc * ah; s2 = ghc * ae; s3 = gec * ah; s4 = gec * ae;
        W.xb()[lane] = j1c;           // parked for the next phase (lane-private slot)
      }
    DEB_LANES_END
    double ci_hh, ci_he, ci_eh, ci_ee, jq_h, jq_e;     // inverse capacitance; a h' row applied to qh, qe
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
    const double gdt = dt * RD_GAMMA;      // 1 / idg

    // W x = v in place for a second vector with the factorisation above (the tangent right-hand sides); the same
    // sequence as the primal solve inside the stage loop: a column, backward tail sweep, block inverses + Woodbury,
    // a h' row, forward tail sweep
    auto solve_second = [&](double* v) {
      const double x0 = v[0] / x0piv;
      DEB_SYNC();
      DEB_LANES_BEGIN
        for (int e = lane; e < n; e += 32) v[e] = (e == 0) ? x0 : v[e] + W.ja()[e] * x0;
      DEB_LANES_END
      DEB_LANES_BEGIN
        if (lane < nch) {
          const int s = C.ch_stride[lane], L = C.ch_lmax[lane];
          double* rp = v + (C.ch_base[lane] + L * s);
          const double* mp = W.m() + (C.ch_base[lane] + L * s);
          double bp = *rp;
          for (int l = L - 1; l >= 2; --l) { rp -= s; mp -= s; bp = *rp - *mp * bp; *rp = bp; }
        }
      DEB_LANES_END
      DEB_LANES_BEGIN
        W.xb()[lane] = lane < nhb ? v[C.hidx[W.perm()[lane]]] : 0.0;
        if (lane < 8) W.xb()[NHMAX + lane] = 0.0;
      DEB_LANES_END
      DEB_LANES_BEGIN
        DEB_USE(pcol); DEB_USE(s1); DEB_USE(s2); DEB_USE(s3); DEB_USE(pval);
        s1 = s2 = s3 = 0.0; pval = 0.0;
        if (lane < nhb) {
          const int lo = C.blo[lane];
          const double* rb = hrow(W, lane, lo) + lo;
          const double* xb = W.xb() + lo;
          double acc = 0.0;
#pragma unroll
          for (int i = 0; i < 8; ++i) acc += rb[i] * xb[i];
          pval = acc;
          s1 = W.gh()[pcol] * acc; s2 = W.ge()[pcol] * acc; s3 = W.j1()[pcol] * acc;
        }
      DEB_LANES_END
      {
        const double th = DEB_WARP_SUM(s1), te = DEB_WARP_SUM(s2), ta = DEB_WARP_SUM(s3);
        const double sh = ci_hh * th + ci_he * te, se = ci_eh * th + ci_ee * te;
        DEB_LANES_BEGIN
          DEB_USE(pcol); DEB_USE(pval);
          if (lane < nhb) v[C.hidx[pcol]] = pval + W.qh()[pcol] * sh + W.qe()[pcol] * se;
          else if (lane == nhb) v[1] = (v[1] + ta + jq_h * sh + jq_e * se) * gdt;
        DEB_LANES_END
      }
      DEB_LANES_BEGIN
        if (lane < nch) {
          const int s = C.ch_stride[lane], L = C.ch_lmax[lane];
          const int i0 = C.ch_base[lane] + 2 * s;
          double* rp = v + i0;
          const double* ip = W.ie() + i0;
          const double* gp = W.g() + i0;
          double x = *rp;
          for (int l = 3; l <= L; ++l) { rp += s; ip += s; gp += s; x = *rp * *ip + *gp * x; *rp = x; }
        }
      DEB_LANES_END
    };

    // ================= 8 stages =================
    double errnorm2 = 0.0;
#pragma unroll 1
    for (int st = 1; st <= 8; ++st) {
      if (st > 1) {
        // ---- stage state u and evaluation time ----
        const double ts = st == 2 ? t + RD_CT2 * dt : st == 3 ? t + RD_CT3 * dt : st == 4 ? t + RD_CT4 * dt
                        : st == 5 ? t + RD_CT5 * dt : t + dt;
        const double dtd = st == 2 ? dt * RD_D2 : st == 3 ? dt * RD_D3 : st == 4 ? dt * RD_D4 : st == 5 ? dt * RD_D5 : 0.0;
        // (the switch sits outside the element loops so that every k_j keeps a fixed register)
#define DEB_FOR_OWN(body) _Pragma("unroll") for (int j = 0; j < NE; ++j) { const int e = lane + 32 * j; if (e < n) { body } }
        DEB_LANES_BEGIN
          DEB_USE(ks);
          switch (st) {
            case 2: DEB_FOR_OWN(W.u()[e] = W.y()[e] + RD_A21 * ks[0][j];) break;
            case 3: DEB_FOR_OWN(W.u()[e] = W.y()[e] + RD_A31 * ks[0][j] + RD_A32 * ks[1][j];) break;
            case 4: DEB_FOR_OWN(W.u()[e] = W.y()[e] + RD_A41 * ks[0][j] + RD_A42 * ks[1][j] + RD_A43 * ks[2][j];) break;
            case 5: DEB_FOR_OWN(W.u()[e] = W.y()[e] + RD_A51 * ks[0][j] + RD_A52 * ks[1][j] + RD_A53 * ks[2][j] + RD_A54 * ks[3][j];) break;
            case 6: DEB_FOR_OWN(W.u()[e] = W.y()[e] + RD_A61 * ks[0][j] + RD_A62 * ks[1][j] + RD_A63 * ks[2][j] + RD_A64 * ks[3][j] + RD_A65 * ks[4][j];) break;
            case 7: DEB_FOR_OWN(W.u()[e] = W.u()[e] + ks[5][j];) break;
            default: DEB_FOR_OWN(W.u()[e] = W.u()[e] + ks[6][j];) break;
          }
        DEB_LANES_END
        // ---- f(ts, u) + dt d_i dT + sum_j C_ij/dt k_j  -> r ----
        Bg<double> b;
        if (HELPER) {
          DEB_BAR_SYNC(BAR_RDY);                 // the helper finished the evaluation posted during the last solve
          b.a = box->a_req;
          b.H = box->bgs[0]; b.gc = box->bgs[1]; b.gb = box->bgs[2]; b.gg = box->bgs[3]; b.gr = box->bgs[4]; b.gnu = box->bgs[5];
          b.gq = box->bgs[6]; b.wq1 = box->bgs[7]; b.ca2 = box->bgs[8]; b.wq = box->bgs[9];
          b.opac = W.sl()[SL_OPAC]; b.pbo = W.sl()[SL_PBO]; b.cs2 = 0.0;
          DEB_LANES_BEGIN
            if (lane == 0) W.u()[0] = b.a;         // exactly the value the helper used
            if (lane >= 3 && lane < nch) nu_moments_lane(nb, box->vv, W.u(), P.iq0, lane - 3, W.nur(), W.nup());
          DEB_LANES_END
        } else {
          compute_bg<double>(c, nb, nq, W.u()[0], hint, W.ic(), b);
          DEB_LANES_BEGIN
            if (lane < nch) chain_coeffs_lane<double>(c, nb, b, k, lane, W.u(), P.iq0, W.kc(), W.kap(), W.nur(), W.nup(), W.sl());
            if (lane == 0) fill_slots<double>(c, b, k, W.sl());
          DEB_LANES_END
        }
        Metric<double> mt;
        compute_metric<double>(P, c, nb, b, W.u(), k, W.nur(), W.nup(), mt);
        const double invts = DEB_RCP(ts);
        DEB_LANES_BEGIN
          DEB_USE(ks);
          switch (st) {
            case 2: DEB_FOR_OWN(W.r()[e] = invdt * (RD_C21 * ks[0][j]);) break;
            case 3: DEB_FOR_OWN(W.r()[e] = invdt * (RD_C31 * ks[0][j] + RD_C32 * ks[1][j]);) break;
            case 4: DEB_FOR_OWN(W.r()[e] = invdt * (RD_C41 * ks[0][j] + RD_C42 * ks[1][j] + RD_C43 * ks[2][j]);) break;
            case 5: DEB_FOR_OWN(W.r()[e] = invdt * (RD_C51 * ks[0][j] + RD_C52 * ks[1][j] + RD_C53 * ks[2][j] + RD_C54 * ks[3][j]);) break;
            case 6: DEB_FOR_OWN(W.r()[e] = invdt * (RD_C61 * ks[0][j] + RD_C62 * ks[1][j] + RD_C63 * ks[2][j] + RD_C64 * ks[3][j] + RD_C65 * ks[4][j]);) break;
            case 7: DEB_FOR_OWN(W.r()[e] = invdt * (RD_C71 * ks[0][j] + RD_C72 * ks[1][j] + RD_C73 * ks[2][j] + RD_C74 * ks[3][j] + RD_C75 * ks[4][j] + RD_C76 * ks[5][j]);) break;
            default: DEB_FOR_OWN(W.r()[e] = invdt * (RD_C81 * ks[0][j] + RD_C82 * ks[1][j] + RD_C83 * ks[2][j] + RD_C84 * ks[3][j] + RD_C85 * ks[4][j] + RD_C86 * ks[5][j] + RD_C87 * ks[6][j]);) break;
          }
          if (TAN) { DEB_FOR_OWN(TW->cc[e] = W.r()[e];) }
        DEB_LANES_END
        const double dtt = dtd * invt0 * invt0;
#ifndef DEB_CPU_EMU
        // two-warp variant: the helper takes the upper part of the tail rows (it idles until the next request)
        const int split = HELPER ? 32 * (((C.ntail + 31) / 32) * 3 / 7) : C.ntail;
        if (HELPER) {
          DEB_LANE0_BEGIN
            box->invts = invts; box->dtt = dtt; box->split = split;
          DEB_LANE0_END
          DEB_BAR_ARRIVE(BAR_GO);
        }
#else
        const int split = C.ntail;
#endif
        DEB_LANES_BEGIN      // rows are distributed differently from the register-resident k's: new phase
          if (lane < nh) {
            const int e = C.hidx[lane];
            W.r()[e] += head_row<double>(C, W.sl(), W.u(), lane, mt);
          }
          if (lane == 0) W.r()[0] += b.H * b.a;
        DEB_LANES_END
        {
          stage_tail_rows(C, W, 0, split, invts, dtt DEB_LANE_ARG);
#ifndef DEB_CPU_EMU
          if (HELPER) DEB_BAR_SYNC(BAR_DONE);
          else
#endif
          (void)0;
        }
      }

      // ---- solve W x = r in place ----
      const double x0 = W.r()[0] / x0piv;
      DEB_SYNC();          // every lane has read r[0] before lane 0 overwrites it with x0
      if (HELPER && st < 8) {
        // scale factor of stage st+1: element 0 of y + sum_j a_{st+1,j} k_j with k_st[0] = x0 (lane 0 owns element 0)
        DEB_LANES_BEGIN
          DEB_USE(ks);
          if (lane == 0) {
            double an;
            switch (st) {
              case 1: an = W.y()[0] + RD_A21 * x0; break;
              case 2: an = W.y()[0] + RD_A31 * ks[0][0] + RD_A32 * x0; break;
              case 3: an = W.y()[0] + RD_A41 * ks[0][0] + RD_A42 * ks[1][0] + RD_A43 * x0; break;
              case 4: an = W.y()[0] + RD_A51 * ks[0][0] + RD_A52 * ks[1][0] + RD_A53 * ks[2][0] + RD_A54 * x0; break;
              case 5: an = W.y()[0] + RD_A61 * ks[0][0] + RD_A62 * ks[1][0] + RD_A63 * ks[2][0] + RD_A64 * ks[3][0] + RD_A65 * x0; break;
              default: an = W.u()[0] + x0; break;        // stages 7 and 8: u + k
            }
            box->a_req = an;
          }
        DEB_LANES_END
#ifdef DEB_CPU_EMU
        helper_compute(P, C, W, *box, hint2);
#else
        DEB_BAR_ARRIVE(BAR_REQ);
#endif
      }
      DEB_LANES_BEGIN
        DEB_FOR_OWN(W.r()[e] = (e == 0) ? x0 : W.r()[e] + W.ja()[e] * x0;)
      DEB_LANES_END
      DEB_LANES_BEGIN
        if (lane < nch) {          // backward sweep: b'_l = b_l - m_l b'_{l+1}, l = L-1 .. 2 (loads run one step ahead)
          const int s = C.ch_stride[lane], L = C.ch_lmax[lane];
          double* rp = W.r() + (C.ch_base[lane] + L * s);
          const double* mp = W.m() + (C.ch_base[lane] + L * s);
          double bp = *rp;
          rp -= s; mp -= s;
          double rn = *rp, mn = *mp;
#pragma unroll 4
          for (int l = L - 1; l > 2; --l) {
            const double rc = rn, mc = mn;
            rn = *(rp - s); mn = *(mp - s);
            bp = rc - mc * bp;
            *rp = bp;
            rp -= s; mp -= s;
          }
          *rp = rn - mn * bp;
        }
      DEB_LANES_END
      // head: p = D^-1 b (block inverses), then the rank-2 Woodbury correction and the a h' row
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
          for (int i = 0; i < 8; ++i) { ra[i] = rb[i]; xa[i] = xb[i]; }      // zero slots x finite padding
          const double acc = ((ra[0] * xa[0] + ra[1] * xa[1]) + (ra[2] * xa[2] + ra[3] * xa[3]))
                           + ((ra[4] * xa[4] + ra[5] * xa[5]) + (ra[6] * xa[6] + ra[7] * xa[7]));
          pval = acc;                       // p_c for the column c = pcol this lane's row became
          s1 = W.gh()[pcol] * acc; s2 = W.ge()[pcol] * acc; s3 = W.j1()[pcol] * acc;
        }
      DEB_LANES_END
      {
        const double th = DEB_WARP_SUM(s1), te = DEB_WARP_SUM(s2), ta = DEB_WARP_SUM(s3);
        const double sh = ci_hh * th + ci_he * te, se = ci_eh * th + ci_ee * te;
        DEB_LANES_BEGIN
          DEB_USE(pcol); DEB_USE(pval);
          if (lane < nhb) W.r()[C.hidx[pcol]] = pval + W.qh()[pcol] * sh + W.qe()[pcol] * se;
          else if (lane == nhb) W.r()[1] = (W.r()[1] + ta + jq_h * sh + jq_e * se) * gdt;     // a h' (state 1): closed row
        DEB_LANES_END
      }
      DEB_LANES_BEGIN
        if (lane < nch) {          // forward sweep: x_l = b'_l/e_l + g_l x_{l-1}, l = 3 .. L
          const int s = C.ch_stride[lane], L = C.ch_lmax[lane];
          const int i0 = C.ch_base[lane] + 2 * s;
          double* rp = W.r() + i0;
          const double* ip = W.ie() + i0;
          const double* gp = W.g() + i0;
          double x = *rp;
          rp += s; ip += s; gp += s;
          double cn = *rp * *ip, gn = *gp;
#pragma unroll 4
          for (int l = 3; l < L; ++l) {
            const double cc = cn, gc = gn;
            cn = *(rp + s) * *(ip + s); gn = *(gp + s);
            x = cc + gc * x;
            *rp = x;
            rp += s; ip += s; gp += s;
          }
          *rp = cn + gn * x;
        }
      DEB_LANES_END
      // ---- keep k_st in registers ----
      if (st < 8) {
        DEB_LANES_BEGIN
          DEB_USE(ks);
          switch (st) {
            case 1: DEB_FOR_OWN(ks[0][j] = W.r()[e];) break;
            case 2: DEB_FOR_OWN(ks[1][j] = W.r()[e];) break;
            case 3: DEB_FOR_OWN(ks[2][j] = W.r()[e];) break;
            case 4: DEB_FOR_OWN(ks[3][j] = W.r()[e];) break;
            case 5: DEB_FOR_OWN(ks[4][j] = W.r()[e];) break;
            case 6: DEB_FOR_OWN(ks[5][j] = W.r()[e];) break;
            default: DEB_FOR_OWN(ks[6][j] = W.r()[e];) break;
          }
        DEB_LANES_END
      }
      if (TAN) {
        // ---- tangent stage: rdot_i, solve with the same factorisation, keep kdot_i ----
        tan_stage_rhs(P, C, W, *TW, st, kD, t, td, dt, ddt, hint DEB_LANE_ARG);
        solve_second(TW->rd);
        if (st < 8) {
          DEB_LANES_BEGIN
            for (int e = lane; e < n; e += 32) TW->kd[(size_t)(st - 1) * P.np + e] = TW->rd[e];
          DEB_LANES_END
        }
      }
    }
    // y1 = u + k8 -> u ; err = k8 = r
    DEB_LANES_BEGIN
      DEB_USE(nanflag);
      nanflag = 0;
      DEB_FOR_OWN(const double y1v = W.u()[e] + W.r()[e]; W.u()[e] = y1v; nanflag |= (y1v != y1v);)
      if (TAN) { DEB_FOR_OWN(TW->ud[e] = TW->ud[e] + TW->rd[e];) }
    DEB_LANES_END

    if (P.mode == 1) {
      DEB_LANES_BEGIN
        for (int e = lane; e < n; e += 32) { P.dbg_y1[item * n + e] = W.u()[e]; P.dbg_err[item * n + e] = W.r()[e]; }
        if (TAN) { for (int e = lane; e < n; e += 32) P.dbg_dy1[item * n + e] = TW->ud[e]; }
      DEB_LANES_END
      return;
    }

    // ================= error norm, PID controller (diffrax semantics) =================
    {
      bool anynan = DEB_ANY(nanflag) != 0;
      const double ik2 = DEB_RCP(k2);
#define DEB_ERRC(e, w) { double y0v = W.y()[e], y1v = anynan ? y0v : W.u()[e], ev = W.r()[e]; if (ev != ev) ev = INFINITY; \
        double sc = ev / (P.atol + fmax(fabs(y0v), fabs(y1v)) * P.rtol) * (w); errnorm2 += sc * sc; }
      DEB_ERRC(0, 1.0) DEB_ERRC(2, k2) DEB_ERRC(3, 1.0) DEB_ERRC(5, 1.0) DEB_ERRC(6, ik2) DEB_ERRC(7, 1.0)
      if (BATCH && P.mode != 3) {
        // one norm for the batch: sum the 6 B squared components; a NaN candidate anywhere in the batch switches
        // EVERY mode's error scale to y0 (diffrax tests the whole (B, n) array), which needs a second round
        double tot, nanc;
        batch_exchange(*BC, errnorm2, anynan ? 1.0 : 0.0, BATCH_SUM, &tot, &nanc);
        if (nanc > 0.0) {
          if (!anynan) {
            anynan = true; errnorm2 = 0.0;
            DEB_ERRC(0, 1.0) DEB_ERRC(2, k2) DEB_ERRC(3, 1.0) DEB_ERRC(5, 1.0) DEB_ERRC(6, ik2) DEB_ERRC(7, 1.0)
          }
          batch_exchange(*BC, errnorm2, 0.0, BATCH_SUM, &tot, &nanc);
        }
        errnorm2 = tot / (double)BC->B;
      }
#undef DEB_ERRC
      DEB_SYNC();          // y, u, r are rewritten below (output sampling, accepted state)
    }
    const double E = sqrt(errnorm2 / 6.0);
    const bool keep = (P.mode == 3) ? (DEB_LDG(P.rp_keep + (size_t)mode * P.rp_stride + nsteps) != 0) : (E < 1.0);
    double inv = 1.0 / E;
    // inv^c1 * inv_prev^c2 * inv_pprev^c3 (PID law) through one exp; E = 0 or inf keep their limits
    double f1 = P.c1 != 0.0 ? ((inv > 0.0 && !isinf(inv)) ? DEB_EXP(P.c1 * DEB_LOG(inv)) : DEB_POW(inv, P.c1)) : 1.0;
    double f2 = P.c2 != 0.0 ? DEB_EXP(P.c2 * DEB_LOG(inv_prev)) : 1.0;
    double f3 = P.c3 != 0.0 ? DEB_EXP(P.c3 * DEB_LOG(inv_pprev)) : 1.0;
    double fac = fmin(fmax(P.safety * f1 * f2 * f3, keep ? 1.0 : P.factormin), P.factormax);
    if (!(fac == fac)) fac = NAN;
    const double dtn = dt * fac;
#ifdef DEB_CPU_EMU_TRACE
    fprintf(stderr, "TRACE %d %d %.17g %.17g %.17g %d\n", mode, nsteps, t, dt, E, (int)keep);
#endif
    if (inv == 0.0 || isinf(inv)) inv = 1.0;
    ++nsteps;
    if (keep) {
      ++nacc;
      // SaveAt(ts): linear interpolation inside the accepted step
      while (save_idx < P.nout && DEB_LDG(tout + save_idx) <= tnext) {
        const double tt = DEB_LDG(tout + save_idx);
        const double coeff = (tnext == t) ? 0.0 : (tt - t) / (tnext - t);
        const size_t obase = ((size_t)mode * P.nout + save_idx);
        const bool wr = !TAN || tan == 0;          // every direction repeats the primal solve; direction 0 writes it
        double coeffd = 0.0;
        if (TAN && tnext != t) coeffd = ((mk(tt, DEB_LDG(toutd + save_idx)) - mk(t, td)) / (mk(tnext, tnextd) - mk(t, td))).d;
        const size_t dbase = item * P.nout + save_idx;
        if (P.return_full) {
          DEB_LANES_BEGIN
            if (wr) { for (int e = lane; e < n; e += 32) P.y_out[obase * n + e] = W.y()[e] + coeff * (W.u()[e] - W.y()[e]); }
            if (TAN) {
              for (int e = lane; e < n; e += 32)
                P.dy_out[dbase * n + e] = TW->yd[e] + coeffd * (W.u()[e] - W.y()[e]) + coeff * (TW->ud[e] - TW->yd[e]);
            }
          DEB_LANES_END
        } else {
          DEB_LANES_BEGIN
            for (int e = lane; e < n; e += 32) W.r()[e] = W.y()[e] + coeff * (W.u()[e] - W.y()[e]);
            if (TAN) {
              for (int e = lane; e < n; e += 32) TW->rd[e] = TW->yd[e] + coeffd * (W.u()[e] - W.y()[e]) + coeff * (TW->ud[e] - TW->yd[e]);
            }
          DEB_LANES_END
          DEB_LANE0_BEGIN
            if (wr) {
              double o20[20];
              convert_outputs(P, c, nb, W.r(), k, o20);
              const bool has_pk = P.pk_out && P.power_idx >= 0;
              double pkv = 0.0;
              if (has_pk) { const double yv = o20[P.power_idx]; pkv = 2.0 * 9.869604401089358 * c.As * DEB_POW(k / c.kp, c.ns - 1.0) * DEB_POW(k, -3.0) * yv * yv; }
              store_fields(P, mode, save_idx, o20, has_pk, pkv);
            }
            if (TAN) {
              Dual od[20];
              const StateVD sy = {W.r(), TW->rd};
              convert_outputs_d(P, *TW->cd, nb, sy, kD, od);
              for (int q = 0; q < 20; ++q) P.dy_out[dbase * 20 + q] = od[q].d;
              if (P.dpk_out && P.power_idx >= 0) P.dpk_out[dbase] = power_d(*TW->cd, kD, od[P.power_idx]).d;
            }
          DEB_LANE0_END
        }
        ++save_idx;
      }
      DEB_LANES_BEGIN
        DEB_FOR_OWN(W.y()[e] = W.u()[e];)
        if (TAN) { DEB_FOR_OWN(TW->yd[e] = TW->ud[e];) }
      DEB_LANES_END
      inv_pprev = inv_prev; inv_prev = inv;
      td = (tnext <= t1) ? tnextd : t1d;
      t = fmin(tnext, t1);
    }
    // dt_next = dt * factor: the factor is non-differentiable (diffrax), the step length carries its tangent
    double tn = t + dtn, tnd = td + ddt * fac;
    if (tn > t1 - 1e-10) {
      if (keep) { tn = t1; tnd = t1d; } else { tn = t + 0.5 * (t1 - t); tnd = td + 0.5 * (t1d - td); }
    }
    tnext = tn; tnextd = tnd;
    if (!(tnext == tnext) || isinf(tnext)) status = 2;
  }
  if (status == 0 && t < t1) status = 1;
  DEB_LANE0_BEGIN
    if (!TAN || tan == 0) {
      store_status(P, mode, status, nsteps, nacc);
    }
  DEB_LANE0_END
}

}  // namespace deb
