// deb_team.cu -- sm_100a kernel of the CTA-per-mode ("team") variant (deb_team.cuh) and its launcher.
//
// Launch geometry: one CTA of TEAM warps integrates one (cosmology, k) mode; CTAs pull modes from the global
// ticket counter, largest k first.  Chosen by launch_evolve (deb_kernels.cu) for launches of at most
// DEB_TEAM_MAX_PER_SM modes per SM, where the time of the launch is the latency of its slowest mode.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../include/discoeb_b200.h"
// one out-of-line copy of each transcendental for this translation unit (see DEB_EXP in deb_core.cuh)
static __device__ __noinline__ double deb_ni_exp(double x) { return exp(x); }
static __device__ __noinline__ double deb_ni_log(double x) { return log(x); }
static __device__ __noinline__ double deb_ni_pow(double x, double y) { return pow(x, y); }
#define DEB_EXP(x) deb_ni_exp(x)
#define DEB_LOG(x) deb_ni_log(x)
#define DEB_POW(x, y) deb_ni_pow(x, y)
#define DEB_COLD static __device__ __noinline__
#include "deb_core.cuh"
#ifdef DEB_TEAM_TIMING
__device__ long long g_team_timing[32];
__device__ int g_team_skip;
__device__ long long g_mode_log[4096][4];      // per mode: start ns, end ns, SM id, CTA slot on the SM (measurement builds)
#endif
#include "deb_team.cuh"

using namespace deb;

#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "[discoeb_b200] CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return DEB_E_CUDA; } } while (0)

static __host__ __device__ size_t al16(size_t b) { return (b + 15) & ~(size_t)15; }
static __host__ __device__ size_t team_smem_bytes(int np, int segrows) {
  return al16(sizeof(CtaConst)) + 4 * al16((size_t)np * sizeof(int)) + al16((warp_ws_doubles(np) + team_ws_doubles(segrows)) * sizeof(double))
       + al16(sizeof(TeamBox)) + 16;
}

template <int NE, int TEAM, int MINB>
__global__ void __launch_bounds__(32 * TEAM, MINB) k_evolve_team(const __grid_constant__ Problem P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CtaConst* C = reinterpret_cast<CtaConst*>(smem_raw);
  size_t off = al16(sizeof(CtaConst));
  int* tail = reinterpret_cast<int*>(smem_raw + off);
  off += al16((size_t)P.np * sizeof(int));
  int* eslot = reinterpret_cast<int*>(smem_raw + off);
  off += al16((size_t)P.np * sizeof(int));
  int* epos = reinterpret_cast<int*>(smem_raw + off);
  off += al16((size_t)P.np * sizeof(int));
  int* tailpos = reinterpret_cast<int*>(smem_raw + off);
  off += al16((size_t)P.np * sizeof(int));
  const int segrows = team_segrows(P.lmaxg, P.lmaxgp, P.lmaxr, P.lmaxnu);
  double* wsb = reinterpret_cast<double*>(smem_raw + off);
  off += al16((warp_ws_doubles(P.np) + team_ws_doubles(segrows)) * sizeof(double));
  TeamBox* box = reinterpret_cast<TeamBox*>(smem_raw + off);
  off += al16(sizeof(TeamBox));
  unsigned int* s_tk = reinterpret_cast<unsigned int*>(smem_raw + off);
  const int tid = threadIdx.x;
  init_cta_const(P, *C, tail, tid, 32 * TEAM);
  __syncthreads();
  init_team_const(P, *C, eslot, epos, tailpos, tid, 32 * TEAM);
  __syncthreads();
  WarpWs W;
  carve(W, wsb, P.np);
  TeamWs X;
  carve_team(X, wsb + warp_ws_doubles(P.np), segrows, eslot, epos, tailpos);
  // Two CTAs share an SM and warp slot w issues from scheduler w % 4: with the same role order in both, the two
  // serial warps (and the two helpers) would compete for one scheduler while two others idle (+15 % per step,
  // measured).  The CTA on the odd group of warp slots rotates its roles by two warps.
  if (tid == 0) {
    unsigned int wslot;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wslot));
    s_tk[1] = ((wslot / TEAM) & 1u) * (TEAM / 2);
  }
  __syncthreads();
  const int ltid = ((((tid >> 5) + TEAM - (int)s_tk[1]) % TEAM) << 5) | (tid & 31);
  const int total = P.ncosmo * P.nk;
  // Work list.  Default: largest k first, cosmologies interleaved.  When the previous launch of this shape left a
  // cost-ordered list in the workspace (k_learn_order below), positions follow descending step count, and the FIRST
  // position of a CTA is static: the hardware places CTAs 2i and 2i+1 on one SM, so even CTAs take the G/2 longest
  // modes and odd CTAs the next G/2 -- every SM starts with one long mode instead of two neighbours in k that stay
  // paired (and out of step) for the whole launch.
  const int G = gridDim.x;
  // (the list is used only when this launch's pre-kernel verified it entry by entry: k_tau_out -> ticket[2])
  const bool ordered = P.order_hdr && P.ticket[2] == 1u;
  const int* order = P.order_hdr + 8;
  // hybrid launch with an accepted list: CTA b < split integrates the b-th longest mode and nothing else (one CTA per SM:
  // the second wave of CTAs leaves at once and makes room for the chain-lane CTAs that take the shorter modes)
  const bool hybrid = ordered && P.hybrid_split > 0;
  // (positions [0, nteam) of the list are the team's: the modes of more than 0.38 x the longest step count, see k_learn_order)
  const int nteam = hybrid ? P.order_hdr[3] : 0;
  bool first = true;
  for (;;) {
    if (tid == 0) {
      unsigned int t;
      if (hybrid) {
        if ((int)blockIdx.x >= P.hybrid_split) t = (unsigned int)total;
        else if (first) t = (unsigned int)blockIdx.x < (unsigned int)nteam ? (unsigned int)blockIdx.x : (unsigned int)total;
        else { const unsigned int q = (unsigned int)P.hybrid_split + atomicAdd(P.ticket, 1u); t = q < (unsigned int)nteam ? q : (unsigned int)total; }
      }
      else if (ordered && first) t = (G & 1) ? (unsigned int)blockIdx.x : (unsigned int)((blockIdx.x & 1) * (G / 2) + (blockIdx.x >> 1));
      else t = (ordered ? (unsigned int)G : 0u) + atomicAdd(P.ticket, 1u);
      *s_tk = t;
    }
    first = false;
    __syncthreads();
    const unsigned int tk = *s_tk;
    if (tk >= (unsigned int)total) break;
    const int kd = tk / P.ncosmo, cs = tk - kd * P.ncosmo;
    const int mode = ordered ? order[tk] : cs * P.nk + (P.nk - 1 - kd);
#ifdef DEB_TEAM_TIMING
    long long t_start_ = 0;
    if (tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start_));
#endif
    integrate_mode_team<NE, TEAM>(P, *C, W, X, *box, mode, ltid, 0, nullptr);
    __syncthreads();
#ifdef DEB_TEAM_TIMING
    if (tid == 0 && mode < 4096) {
      long long t_end_; unsigned int smid_;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end_));
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid_));
      g_mode_log[mode][0] = t_start_; g_mode_log[mode][1] = t_end_; g_mode_log[mode][2] = smid_; g_mode_log[mode][3] = s_tk[1];
    }
#endif
  }
}

// Two teams of four warps in ONE CTA per SM (instead of two CTAs of one team): same work per SM, one copy of the CTA
// constants, and the two serial warps advance through the step code together (DuoSync, deb_team.cuh).  Team q of the CTA
// synchronises on named barriers 1+5q .. 5+5q; barrier 0 is used by the whole CTA during set-up only.
static __host__ __device__ size_t duo_team_bytes(int np, int segrows) {
  return al16((warp_ws_doubles(np) + team_ws_doubles(segrows)) * sizeof(double)) + al16(sizeof(TeamBox));
}
static __host__ __device__ size_t duo_smem_bytes(int np, int segrows) {
  return al16(sizeof(CtaConst)) + 4 * al16((size_t)np * sizeof(int)) + 2 * duo_team_bytes(np, segrows) + 64;
}
template <int NE>
__global__ void __launch_bounds__(256, 1) k_evolve_duo(const __grid_constant__ Problem P) {
  constexpr int TEAM = 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CtaConst* C = reinterpret_cast<CtaConst*>(smem_raw);
  size_t off = al16(sizeof(CtaConst));
  int* tail = reinterpret_cast<int*>(smem_raw + off);
  off += al16((size_t)P.np * sizeof(int));
  int* eslot = reinterpret_cast<int*>(smem_raw + off);
  off += al16((size_t)P.np * sizeof(int));
  int* epos = reinterpret_cast<int*>(smem_raw + off);
  off += al16((size_t)P.np * sizeof(int));
  int* tailpos = reinterpret_cast<int*>(smem_raw + off);
  off += al16((size_t)P.np * sizeof(int));
  const int segrows = team_segrows(P.lmaxg, P.lmaxgp, P.lmaxr, P.lmaxnu);
  const int tq = threadIdx.x >> 7, tid = threadIdx.x & 127;          // team of the CTA, thread of the team
  const size_t per_team = duo_team_bytes(P.np, segrows);
  double* wsb = reinterpret_cast<double*>(smem_raw + off + tq * per_team);
  TeamBox* box = reinterpret_cast<TeamBox*>(smem_raw + off + tq * per_team + al16((warp_ws_doubles(P.np) + team_ws_doubles(segrows)) * sizeof(double)));
  // [0..1] ticket of team 0 / 1, [2..3] role rotation, [4..5] the rendezvous counters
  unsigned int* s_ctl = reinterpret_cast<unsigned int*>(smem_raw + off + 2 * per_team);
  init_cta_const(P, *C, tail, threadIdx.x, 256);
  if (threadIdx.x < 2) reinterpret_cast<int*>(s_ctl + 4)[threadIdx.x] = 0x7fffffff;
  __syncthreads();
  init_team_const(P, *C, eslot, epos, tailpos, threadIdx.x, 256);
  WarpWs W;
  carve(W, wsb, P.np);
  TeamWs X;
  carve_team(X, wsb + warp_ws_doubles(P.np), segrows, eslot, epos, tailpos);
  // warp slot w issues from scheduler w % 4: the second team rotates its roles by two warps so that the two serial warps
  // (and the two background warps) sit on different schedulers (see k_evolve_team)
  if (tid == 0) {
    unsigned int wslot;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wslot));
    s_ctl[2 + tq] = ((wslot / TEAM) & 1u) * (TEAM / 2);
  }
  __syncthreads();
  const int ltid = ((((tid >> 5) + TEAM - (int)s_ctl[2 + tq]) % TEAM) << 5) | (tid & 31);
  const int bb = 1 + 5 * tq;
  DuoSync duo;
  duo.mine = reinterpret_cast<int*>(s_ctl + 4) + tq;
  duo.other = reinterpret_cast<int*>(s_ctl + 4) + (tq ^ 1);
  duo.v = 0; duo.every = P.lockstep; duo.attached = false;
  duo.K = 1 + (P.lockstep ? (8 + P.lockstep - 1) / P.lockstep : 0);
  const int total = P.ncosmo * P.nk;
  // work list as in k_evolve_team, with G = 2 gridDim.x teams: the teams of CTA b start on positions b and G/2 + b of an
  // accepted cost-ordered list (one long and one medium mode per SM), then everybody pulls from the ticket counter
  const int G = 2 * gridDim.x;
  const bool ordered = P.order_hdr && P.ticket[2] == 1u;
  const int* order = P.order_hdr + 8;
  // The nsolo longest modes (CTAs 0 .. nsolo-1) keep their SM to themselves: team 1 of those CTAs leaves at once, and its
  // static positions G/2 .. G/2+nsolo-1 are served first by the ticket counter.  Any nsolo in [0, G/2] keeps the map
  // position <-> (CTA, team, ticket) one-to-one; the value only moves work around (results do not depend on it).
  int nsolo = 0;
  if (ordered) {
    nsolo = P.order_hdr[4];
    const int cap = (int)gridDim.x / 4;              // never give up more than a quarter of the second teams
    nsolo = nsolo < 0 ? 0 : (nsolo > cap ? cap : nsolo);
    if (total < G) nsolo = 0;                        // (not every team has a first mode: nothing to rearrange)
    if (total > 2 * G) nsolo = 0;                    // deeper launches are bound by the sum of the work, not by the longest mode:
                                                     // the slots given up cost more there (768 modes 21.7 -> 22.6 ms, 512 modes 20.1 -> 19.5)
  }
  if (P.duo_nosolo) nsolo = 0;
  bool first = true;
  for (;;) {
    if (tid == 0) {
      unsigned int t;
      if (ordered && first) t = (tq == 1 && (int)blockIdx.x < nsolo) ? (unsigned int)total : (unsigned int)(tq * (G / 2) + blockIdx.x);
      else if (ordered) { const unsigned int q = atomicAdd(P.ticket, 1u); t = q < (unsigned int)nsolo ? (unsigned int)(G / 2) + q : (unsigned int)G + (q - (unsigned int)nsolo); }
      else t = atomicAdd(P.ticket, 1u);
      s_ctl[tq] = t;
    }
    first = false;
    asm volatile("bar.sync %0, 128;" ::"r"(bb) : "memory");
    const unsigned int tk = s_ctl[tq];
    if (tk >= (unsigned int)total) break;
    const int kd = tk / P.ncosmo, cs = tk - kd * P.ncosmo;
    const int mode = ordered ? order[tk] : cs * P.nk + (P.nk - 1 - kd);
    integrate_mode_team<NE, TEAM, true>(P, *C, W, X, *box, mode, ltid, bb, &duo);
    if ((ltid >> 5) == 0) duo.detach(ltid & 31);
    asm volatile("bar.sync %0, 128;" ::"r"(bb) : "memory");
  }
}

// After a team launch: mode ids by descending attempted-step count -> workspace, for the next launch of the same shape
// (MCMC / emulator loops call with near-identical inputs).  One block, bitonic sort of <= 2048 keys in shared memory.
__global__ void __launch_bounds__(1024) k_learn_order(const int* __restrict__ nsteps, int total, int* hdr, int shape_hash) {
  __shared__ unsigned int key[2048];
  for (int i = threadIdx.x; i < 2048; i += 1024) {
    unsigned int v = 0u;
    if (i < total) { const int ns = nsteps[i]; v = ((unsigned int)(ns < 0 ? 0 : (ns > 0xfffff ? 0xfffff : ns)) << 11) | (unsigned int)(2047 - i); v += 1u << 31; }
    key[i] = v;                                  // real entries carry the top bit: padding sorts last
  }
  __syncthreads();
  for (int k2 = 2; k2 <= 2048; k2 <<= 1)
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < 2048; i += 1024) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned int a = key[i], b = key[l];
          const bool desc = (i & k2) == 0;       // descending overall
          if (desc ? a < b : a > b) { key[i] = b; key[l] = a; }
        }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < total; i += 1024) hdr[8 + i] = 2047 - (int)(key[i] & 2047u);
  // hybrid launches: how many modes (from the front of the list) take more than 0.38 x the longest step count.  Those go to
  // team CTAs (31 us per step with chain-lane warps beside them), the others to chain-lane warps (81 us per step measured
  // next to a team CTA): both halves then finish together.
  // two-team CTAs: how many modes are within 5.5 % of the longest step count.  A team beside a working partner needs
  // ~33 us per step, alone on its SM 31.3: these modes -- the ones the launch waits for -- get an SM without a partner
  // (k_evolve_duo), everything else has slack.
  __shared__ int nlong, nsolo;
  if (threadIdx.x == 0) { nlong = 0; nsolo = 0; }
  __syncthreads();
  {
    const unsigned int smax = (key[0] >> 11) & 0xfffffu;
    int c = 0, c2 = 0;
    for (int i = threadIdx.x; i < total; i += 1024) {
      const unsigned int ns = (key[i] >> 11) & 0xfffffu;
      c += (ns * 100u > smax * 38u) ? 1 : 0;
      c2 += (ns * 1000u > smax * 945u) ? 1 : 0;
    }
    atomicAdd(&nlong, c); atomicAdd(&nsolo, c2);
  }
  __syncthreads();
  if (threadIdx.x == 0) { hdr[1] = total; hdr[2] = shape_hash; hdr[3] = nlong; hdr[4] = nsolo; __threadfence(); hdr[0] = DEB_ORDER_MAGIC; }
}

typedef void (*evolve_kernel_t)(const Problem);
int deb_learn_order(const Problem& P, cudaStream_t st);

template <int TEAM, int MINB>
static evolve_kernel_t pick_team(int n) {
  const int ne = (n + 32 * TEAM - 1) / (32 * TEAM);
  if (ne <= 1) return k_evolve_team<1, TEAM, MINB>;
  if (ne <= 2) return k_evolve_team<2, TEAM, MINB>;
  if (ne <= 3) return k_evolve_team<3, TEAM, MINB>;
  return nullptr;
}

// returns DEB_OK after launching, DEB_E_UNSUPPORTED when no team kernel serves this shape (the caller falls back to
// the one-warp kernels, which are CUDA kernels too)
int deb_launch_team(const Problem& P, cudaStream_t st, int nsm) {
  int team = 4, minb = 2;
  if (const char* e = getenv("DEB_TEAM")) team = atoi(e);
  if (const char* e = getenv("DEB_TEAM_MINB")) minb = atoi(e);
  evolve_kernel_t kern = nullptr;
  if (team == 4) kern = minb >= 4 ? pick_team<4, 4>(P.n) : (minb == 3 ? pick_team<4, 3>(P.n) : pick_team<4, 2>(P.n));
  else if (team == 8) kern = pick_team<8, 2>(P.n);
  if (!kern) return DEB_E_UNSUPPORTED;
  // two teams per CTA, serial warps in lock-step (k_evolve_duo): the default geometry; DEB_DUO=0 keeps two CTAs of one team
  int duo = team == 4 && minb == 2 && P.hybrid_split == 0 && P.mode == 0;
  long duo_min = 2L * nsm;
  if (const char* e = getenv("DEB_DUO")) { duo = duo && atoi(e) != 0; if (atoi(e) == 2) duo_min = 2; }      // 2: at any size (tests)
  if (duo) {
    const int ne = (P.n + 127) / 128;
    evolve_kernel_t dk = ne <= 1 ? k_evolve_duo<1> : (ne <= 2 ? k_evolve_duo<2> : (ne <= 3 ? k_evolve_duo<3> : nullptr));
    const size_t dsmem = duo_smem_bytes(P.np, team_segrows(P.lmaxg, P.lmaxgp, P.lmaxr, P.lmaxnu));
    const long total = (long)P.ncosmo * P.nk;
    // (below two teams' worth of modes per SM the one-team CTAs spread over more SMs: measured equal or better there)
    if (dk && dsmem <= 227 * 1024 && total >= duo_min) {
      Problem Q = P;
      Q.lockstep = 1;
      if (const char* e = getenv("DEB_DUO_LOCKSTEP")) Q.lockstep = atoi(e);
      Q.duo_nosolo = 0;
      if (const char* e = getenv("DEB_DUO_SOLO")) Q.duo_nosolo = atoi(e) == 0;
      if (Q.lockstep != 0 && Q.lockstep != 1 && Q.lockstep != 2 && Q.lockstep != 4 && Q.lockstep != 8) Q.lockstep = 1;
      CUDA_TRY(cudaFuncSetAttribute((const void*)dk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem));
      long grid = nsm;
      if (grid > (total + 1) / 2) grid = (total + 1) / 2;
      CUDA_TRY(cudaMemsetAsync(P.ticket, 0, sizeof(unsigned int), st));
      dk<<<(unsigned)grid, 256, dsmem, st>>>(Q);
      CUDA_TRY(cudaGetLastError());
      return deb_learn_order(P, st);
    }
  }
  const size_t smem = team_smem_bytes(P.np, team_segrows(P.lmaxg, P.lmaxgp, P.lmaxr, P.lmaxnu));
  int occ = 0;
  CUDA_TRY(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // (same shared-memory carve-out as the chain-lane kernel, so that CTAs of both can share an SM in hybrid launches)
  CUDA_TRY(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)kern, 32 * team, smem));
  if (occ < 1) return DEB_E_UNSUPPORTED;
  const long total = (long)P.ncosmo * P.nk;
  long grid = (long)nsm * occ;
  if (grid > total) grid = total;
  CUDA_TRY(cudaMemsetAsync(P.ticket, 0, sizeof(unsigned int), st));
  kern<<<(unsigned)grid, 32 * team, smem, st>>>(P);
  CUDA_TRY(cudaGetLastError());
  if (P.hybrid_split > 0) return DEB_OK;       // hybrid launch: the list is refreshed after BOTH kernels (deb_learn_order)
  return deb_learn_order(P, st);
}

int deb_learn_order(const Problem& P, cudaStream_t st) {
  const long total = (long)P.ncosmo * P.nk;
  if (P.order_hdr && P.mode == 0 && total <= DEB_ORDER_MAX && !getenv("DEB_NO_ORDER")) {
    k_learn_order<<<1, 1024, 0, st>>>(P.nsteps, (int)total, P.order_hdr, P.shape_hash);
    CUDA_TRY(cudaGetLastError());
  }
  return DEB_OK;
}

#ifdef DEB_TEAM_TIMING
// measurement build only: cycles warp 0 of the largest-k mode spent per phase (see DEB_TICK), slot 15 = steps
extern "C" int deb_debug_team_mode_log(long long* out, int nmodes) {
  CUDA_TRY(cudaMemcpyFromSymbol(out, g_mode_log, sizeof(long long) * 4 * nmodes));
  return DEB_OK;
}
extern "C" int deb_debug_team_skip(int mask) { CUDA_TRY(cudaMemcpyToSymbol(g_team_skip, &mask, sizeof(int))); return DEB_OK; }
extern "C" int deb_debug_team_timing(long long* out16, int reset) {
  if (out16) CUDA_TRY(cudaMemcpyFromSymbol(out16, g_team_timing, sizeof(long long) * 32));
  if (reset) { long long z[32] = {0}; CUDA_TRY(cudaMemcpyToSymbol(g_team_timing, z, sizeof(z))); }
  return DEB_OK;
}
#endif
