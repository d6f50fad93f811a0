// deb_lane.cu -- sm_100a kernel of the register-resident chain-lane variant (deb_lane.cuh) and its launcher.
//
// Launch geometry: a persistent grid of CTAs of WARPS warps; every warp integrates one (cosmology, k) mode at a time
// and pulls the next from the global ticket counter, largest k first.  The CTA shares the launch-constant tables.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../include/discoeb_b200.h"
// one out-of-line copy of each transcendental for this translation unit: the step loop is executed once per attempted
// step by 8 warps per SM in different phases, so its instruction footprint is what the instruction cache sees
#ifdef DEB_LANE_NOINLINE
static __device__ __noinline__ double deb_ni_exp(double x) { return exp(x); }
static __device__ __noinline__ double deb_ni_log(double x) { return log(x); }
static __device__ __noinline__ double deb_ni_pow(double x, double y) { return pow(x, y); }
#define DEB_EXP(x) deb_ni_exp(x)
#define DEB_LOG(x) deb_ni_log(x)
#define DEB_POW(x, y) deb_ni_pow(x, y)
#define DEB_COLD static __device__ __noinline__
#endif
// tables may sit in shared memory in this translation unit (staged variant): plain generic loads instead of ld.global.nc
#define DEB_LDG(p) (*(p))
#include "deb_core.cuh"
#include "deb_lane.cuh"

using namespace deb;

#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "[discoeb_b200] CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return DEB_E_CUDA; } } while (0)

static __host__ __device__ size_t al16(size_t b) { return (b + 15) & ~(size_t)15; }
template <int NT> static __host__ __device__ size_t lane_smem_bytes(int np, int warps) {
  return al16(sizeof(CtaConst)) + al16((size_t)np * sizeof(int)) + al16(sizeof(LaneTab<NT>)) + warps * al16(sizeof(LaneWs<NT>));
}

template <int NT, int WARPS, int MINB>
__global__ void __launch_bounds__(32 * WARPS, MINB) k_evolve_lane(const __grid_constant__ Problem P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: CtaConst | LaneTab | WARPS x LaneWs | tail table (the only run-time-sized part, last).  The per-warp block
  // offset is made opaque to the compiler: otherwise it re-derives (threadIdx.x >> 5) * sizeof(LaneWs) in front of shared
  // accesses all over the step loop (2.3 k instructions per step, profiles/r2_lane_v3_*) instead of keeping one register.
  CtaConst* C = reinterpret_cast<CtaConst*>(smem_raw);
  constexpr unsigned OFF_T = (unsigned)((sizeof(CtaConst) + 15) & ~(size_t)15);
  constexpr unsigned OFF_W = OFF_T + (unsigned)((sizeof(LaneTab<NT>) + 15) & ~(size_t)15);
  constexpr unsigned SZ_W = (unsigned)((sizeof(LaneWs<NT>) + 15) & ~(size_t)15);
  LaneTab<NT>* T = reinterpret_cast<LaneTab<NT>*>(smem_raw + OFF_T);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned woff = OFF_W + (unsigned)warp * SZ_W;
  asm volatile("" : "+r"(woff));
  LaneWs<NT>& W = *reinterpret_cast<LaneWs<NT>*>(smem_raw + woff);
  int* tail = reinterpret_cast<int*>(smem_raw + OFF_W + WARPS * SZ_W);
  // north_star "tables staged into shared memory by TMA / bulk async copy": for single-cosmology launches the three RHS
  // splines (cs2a, xe, log rho_nu: the first 6 nth + 3 nnu doubles of the table set, 24 KB) are brought in ONCE per CTA by
  // one cp.async.bulk tracked by an mbarrier; A/B against the read-only-cache path: profiles/r2_table_staging_ab.txt
  const double* stab = nullptr;
  if (P.stage_tables) {
    const unsigned off_tab = ((OFF_W + WARPS * SZ_W + (((unsigned)P.np * 4u + 15u) & ~15u) + 127u) & ~127u) + 128u;
    double* dst = reinterpret_cast<double*>(smem_raw + off_tab);
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(smem_raw + off_tab - 16);
    const unsigned bytes = (unsigned)(6 * P.nth + 3 * P.nnu) * 8u;
    const unsigned mb = (unsigned)__cvta_generic_to_shared(mbar), ds = (unsigned)__cvta_generic_to_shared(dst);
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(ds), "l"(P.tables), "r"(bytes), "r"(mb) : "memory");
    }
    unsigned done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mb) : "memory");
    }
    stab = dst;
  }
  init_cta_const(P, *C, tail, threadIdx.x, 32 * WARPS);
  __syncthreads();
  init_lane_tab<NT>(P, *C, *T, threadIdx.x, 32 * WARPS);
  __syncthreads();
  int total = P.ncosmo * P.nk;
  // hybrid launch: this kernel takes positions hybrid_split.. of the learned work list (the shorter modes), the team kernel
  // the first hybrid_split; with a rejected list the team kernel does everything and this one has nothing to do
  const bool hybrid = P.hybrid_split > 0;
  const int* order = hybrid ? P.order_hdr + 8 : nullptr;
  unsigned int* queue = hybrid ? P.ticket2 : P.ticket;
  int nteam = 0;
  if (hybrid) { if (P.ticket[2] != 1u) return; nteam = P.order_hdr[3]; total -= nteam; }
  LaneSync SY;
  SY.cnt = 32 * WARPS; SY.on = (P.mode == 0 && P.lockstep) ? 1 : 0; SY.every = P.lockstep > 1 ? P.lockstep : 1;
  for (;;) {
    unsigned int tk = 0;
    if (lane == 0) tk = atomicAdd(queue, 1u);
    tk = __shfl_sync(0xffffffffu, tk, 0);
    // largest k first, cosmologies interleaved; mode -1 = out of work: with lock-step on, the warp goes through the SAME
    // call (one copy of the code, the same barrier instructions) to attend the barriers until every warp is done
    int mode = -1;
    if (tk < (unsigned int)total) {
      if (hybrid) mode = order[nteam + tk];
      else { const int kd = tk / P.ncosmo, cs = tk - kd * P.ncosmo; mode = cs * P.nk + (P.nk - 1 - kd); }
    }
    if (mode < 0 && !SY.on) break;
    integrate_mode_lane<NT>(P, *C, *T, W, SY, mode, lane, stab);
    __syncwarp();
    if (mode < 0) break;
  }
}

typedef void (*evolve_kernel_t)(const Problem);

struct LaneKernel { evolve_kernel_t fn; size_t smem; };
template <int WARPS, int MINB>
static LaneKernel pick_lane(int nt, int np) {
  switch (nt) {
    case 1: return {k_evolve_lane<1, WARPS, MINB>, lane_smem_bytes<1>(np, WARPS)};
    case 2: return {k_evolve_lane<2, WARPS, MINB>, lane_smem_bytes<2>(np, WARPS)};
    case 3: return {k_evolve_lane<3, WARPS, MINB>, lane_smem_bytes<3>(np, WARPS)};
    case 4: return {k_evolve_lane<4, WARPS, MINB>, lane_smem_bytes<4>(np, WARPS)};
    case 6: return {k_evolve_lane<6, WARPS, MINB>, lane_smem_bytes<6>(np, WARPS)};
    case 8: return {k_evolve_lane<8, WARPS, MINB>, lane_smem_bytes<8>(np, WARPS)};
    default: return {nullptr, 0};
  }
}

static LaneKernel pick_lane12(int nt, int np) {
  switch (nt) {
    case 1: return {k_evolve_lane<1, 12, 1>, lane_smem_bytes<1>(np, 12)};
    case 2: return {k_evolve_lane<2, 12, 1>, lane_smem_bytes<2>(np, 12)};
    case 3: return {k_evolve_lane<3, 12, 1>, lane_smem_bytes<3>(np, 12)};
    case 4: return {k_evolve_lane<4, 12, 1>, lane_smem_bytes<4>(np, 12)};
    default: return {nullptr, 0};
  }
}

// can the chain-lane kernel serve this shape (used by the hybrid launch before it commits the team kernel to a subset)
int deb_lane_supported(const Problem& P) {
  const int nt = lane_nt(P.lmaxg, P.lmaxgp, P.lmaxr, P.lmaxnu);
  return nt != 0 && LN_NSEG * P.nch <= 32 && P.nh <= 32 && P.n <= 34 + 32 * nt && pick_lane<4, 2>(nt, P.np).smem <= 160 * 1024;
}

// returns DEB_OK after launching, DEB_E_UNSUPPORTED when this variant does not serve the shape.  hybrid_ctas > 0: the
// lane half of a hybrid launch (4-warp CTAs, as many as fit beside one team CTA per SM), queue counter P.ticket2.
int deb_launch_lane(const Problem& P, cudaStream_t st, int nsm, int hybrid_ctas) {
  const int nt = lane_nt(P.lmaxg, P.lmaxgp, P.lmaxr, P.lmaxnu);
  if (nt == 0 || LN_NSEG * P.nch > 32 || P.nh > 32 || P.n > 34 + 32 * nt) return DEB_E_UNSUPPORTED;
  // 8 modes in flight per SM (255 registers each) as ONE CTA of 8 warps advancing in lock-step (LN_BAR): 225.7 -> 199.0 ms
  // on 16384 modes against free-running warps; two CTAs of 4 warps in lock-step: 206.2 ms (profiles/r2_lane_lockstep.txt)
  // small hierarchies (NT <= 4: the stage vectors take <= 56 registers per lane) run 12 warps per SM under 168 registers
  // 12 warps per CTA pay for NT <= 3 (n = 72) once the launch is deep (>= 48 modes per SM: 8192 modes 58.0 vs 64.8 ms,
  // 4096 modes 39.2 vs 37.5 ms); at NT = 4 (n = 111) eight warps stay ahead at every size (profiles/r2_lane_small_n.txt)
  int warps = (nt <= 3 && (long)P.ncosmo * P.nk >= 48L * nsm) ? 12 : 8;
  if (hybrid_ctas > 0) warps = 4;
  else if (const char* e = getenv("DEB_LANE_WARPS")) warps = atoi(e);
  if (warps == 12 && nt > 4) warps = 8;
  const LaneKernel lk = warps == 4 ? pick_lane<4, 2>(nt, P.np) : (warps == 12 ? pick_lane12(nt, P.np) : pick_lane<8, 1>(nt, P.np));
  if (warps != 4 && warps != 12) warps = 8;
  if (!lk.fn) return DEB_E_UNSUPPORTED;
  int occ = 0;
  const size_t smem_max = ((lk.smem + 127) & ~(size_t)127) + 128 + (size_t)(6 * P.nth + 3 * P.nnu) * 8;
  const bool can_stage = smem_max <= 227 * 1024;
  CUDA_TRY(cudaFuncSetAttribute((const void*)lk.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(can_stage ? smem_max : lk.smem)));
  CUDA_TRY(cudaFuncSetAttribute((const void*)lk.fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)lk.fn, 32 * warps, lk.smem));
  if (occ < 1) return DEB_E_UNSUPPORTED;
  const long total = (long)P.ncosmo * P.nk;
  long grid = (long)nsm * occ;
  if (grid * warps > total) grid = (total + warps - 1) / warps;
  if (hybrid_ctas > 0) grid = hybrid_ctas;
  else CUDA_TRY(cudaMemsetAsync(P.ticket, 0, sizeof(unsigned int), st));
  Problem Q = P;
  // lock-step barriers at the top of the Jacobian phase and of stages 1 and 5 (DEB_LANE_LOCKSTEP = 0 off, 1 / 2 / 4 / 8 = every
  // n-th stage: 177.4 / 175.0 / 171.5 / 172.6 ms on 16384 modes, 225.7 ms without)
  Q.lockstep = getenv("DEB_LANE_LOCKSTEP") ? atoi(getenv("DEB_LANE_LOCKSTEP")) : 4;
  // staged tables: on whenever one cosmology serves the whole launch and the copy fits beside the modes (measured -0.4 %
  // at n = 72 / 111; at n = 265 eight modes leave no room); DEB_STAGE_TABLES=0/1 overrides for the A/B
  Q.stage_tables = P.ncosmo == 1 ? 1 : 0;
  if (const char* e = getenv("DEB_STAGE_TABLES")) Q.stage_tables = (atoi(e) && P.ncosmo == 1) ? 1 : 0;
  size_t smem = lk.smem;
  if (Q.stage_tables) smem = ((smem + 127) & ~(size_t)127) + 128 + (size_t)(6 * P.nth + 3 * P.nnu) * 8;
  if (Q.stage_tables && !can_stage) Q.stage_tables = 0;         // no room beside 8 modes at this hierarchy size
  if (!Q.stage_tables) smem = lk.smem;
  lk.fn<<<(unsigned)grid, 32 * warps, smem, st>>>(Q);
  CUDA_TRY(cudaGetLastError());
  return DEB_OK;
}
