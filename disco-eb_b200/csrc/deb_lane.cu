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
  init_cta_const(P, *C, tail, threadIdx.x, 32 * WARPS);
  __syncthreads();
  init_lane_tab<NT>(P, *C, *T, threadIdx.x, 32 * WARPS);
  __syncthreads();
  const int total = P.ncosmo * P.nk;
  LaneSync SY;
  SY.cnt = 32 * WARPS; SY.on = (P.mode == 0 && P.lockstep) ? 1 : 0;
  for (;;) {
    unsigned int tk = 0;
    if (lane == 0) tk = atomicAdd(P.ticket, 1u);
    tk = __shfl_sync(0xffffffffu, tk, 0);
    // largest k first, cosmologies interleaved; mode -1 = out of work: with lock-step on, the warp goes through the SAME
    // call (one copy of the code, the same barrier instructions) to attend the barriers until every warp is done
    int mode = -1;
    if (tk < (unsigned int)total) { const int kd = tk / P.ncosmo, cs = tk - kd * P.ncosmo; mode = cs * P.nk + (P.nk - 1 - kd); }
    if (mode < 0 && !SY.on) break;
    integrate_mode_lane<NT>(P, *C, *T, W, SY, mode, lane);
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

// returns DEB_OK after launching, DEB_E_UNSUPPORTED when this variant does not serve the shape
int deb_launch_lane(const Problem& P, cudaStream_t st, int nsm) {
  const int nt = lane_nt(P.lmaxg, P.lmaxgp, P.lmaxr, P.lmaxnu);
  if (nt == 0 || LN_NSEG * P.nch > 32 || P.nh > 32 || P.n > 34 + 32 * nt) return DEB_E_UNSUPPORTED;
  // 8 modes in flight per SM (255 registers each) as ONE CTA of 8 warps advancing in lock-step (LN_BAR): 225.7 -> 199.0 ms
  // on 16384 modes against free-running warps; two CTAs of 4 warps in lock-step: 206.2 ms (profiles/r2_lane_lockstep.txt)
  int warps = 8;
  if (const char* e = getenv("DEB_LANE_WARPS")) warps = atoi(e);
  const LaneKernel lk = warps == 4 ? pick_lane<4, 2>(nt, P.np) : pick_lane<8, 1>(nt, P.np);
  if (warps != 4) warps = 8;
  if (!lk.fn) return DEB_E_UNSUPPORTED;
  int occ = 0;
  CUDA_TRY(cudaFuncSetAttribute((const void*)lk.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lk.smem));
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)lk.fn, 32 * warps, lk.smem));
  if (occ < 1) return DEB_E_UNSUPPORTED;
  const long total = (long)P.ncosmo * P.nk;
  long grid = (long)nsm * occ;
  if (grid * warps > total) grid = (total + warps - 1) / warps;
  CUDA_TRY(cudaMemsetAsync(P.ticket, 0, sizeof(unsigned int), st));
  Problem Q = P;
  Q.lockstep = getenv("DEB_LANE_LOCKSTEP") ? atoi(getenv("DEB_LANE_LOCKSTEP")) : 1;
  lk.fn<<<(unsigned)grid, 32 * warps, lk.smem, st>>>(Q);
  CUDA_TRY(cudaGetLastError());
  return DEB_OK;
}
