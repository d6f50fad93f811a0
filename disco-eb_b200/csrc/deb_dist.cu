// deb_dist.cu -- multi-GPU entry of the hot path: one k grid dealt round-robin over the ranks of
// one box (mode i -> rank i mod W, because cost rises steeply with k), every rank integrates its share, and every rank
// ends with the full-size result.  Two ways to get it there, both behind deb_evolve_sharded_f64:
//
//   gather = 0  one ncclAllGather of a packed row per mode (20 nout fields, nout P(k), status, step count) on the caller's
//               stream, then a scatter kernel that undoes the round-robin deal;
//   gather = 1  no collective at all: the kernels' epilogue stores every mode's row straight into the full-size buffers
//               of ALL ranks through NVLink peer mappings (Problem::y_peer, store_fields in deb_core.cuh).  The caller
//               supplies the peer pointers (symmetric / IPC-mapped allocations) and a cross-rank barrier afterwards.
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 the process already has, e.g. PyTorch's): the library keeps no
// link-time dependency on it and single-GPU users never load it.  The communicator is the library's own
// (deb_comm_create: ncclCommInitRank from an id the caller broadcasts by whatever means it has).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/discoeb_b200.h"

#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "[discoeb_b200] CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return DEB_E_CUDA; } } while (0)

// ---- the four NCCL entry points used, resolved lazily ----
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8 };
static struct {
  void* h;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(ncclResult_t);
} g_nccl;

static int nccl_load() {
  if (g_nccl.h) return DEB_OK;
  const char* names[] = {getenv("DEB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (int i = 0; i < 3 && !h; ++i) if (names[i]) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!h) { fprintf(stderr, "[discoeb_b200] cannot load NCCL (%s); set DEB_NCCL_LIB\n", dlerror()); return DEB_E_UNSUPPORTED; }
  *(void**)&g_nccl.GetUniqueId = dlsym(h, "ncclGetUniqueId");
  *(void**)&g_nccl.CommInitRank = dlsym(h, "ncclCommInitRank");
  *(void**)&g_nccl.CommDestroy = dlsym(h, "ncclCommDestroy");
  *(void**)&g_nccl.AllGather = dlsym(h, "ncclAllGather");
  *(void**)&g_nccl.GetErrorString = dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllGather) return DEB_E_UNSUPPORTED;
  g_nccl.h = h;
  return DEB_OK;
}
#define NCCL_TRY(x) do { ncclResult_t r_ = (x); if (r_ != 0) { \
  fprintf(stderr, "[discoeb_b200] NCCL error %s at %s:%d\n", g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?", __FILE__, __LINE__); return DEB_E_CUDA; } } while (0)

// (the host entry keeps its device arena, stream and events on the communicator: nothing is allocated per call, and the
//  work list learned by one call -- it lives in the workspace part of the arena -- serves the next one)
struct deb_comm {
  ncclComm_t nccl; int world, rank;
  char* arena = nullptr; size_t arena_bytes = 0; cudaStream_t st = nullptr; cudaEvent_t e0 = nullptr, e1 = nullptr;
};

// local share of rank r: modes r, r + W, ...
static inline int share(int nk, int world, int rank) { return nk > rank ? (nk - rank + world - 1) / world : 0; }

// k_local[j] = k_all[j W + r]
__global__ void k_take_modes(const double* __restrict__ kall, double* __restrict__ kloc, int nloc, int world, int rank) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < nloc) kloc[j] = kall[(size_t)j * world + rank];
}
// one packed row per (cosmology, local mode): [20 nout fields | nout P(k) | status | steps]
__global__ void k_pack_rows(const double* __restrict__ y, const double* __restrict__ pk, const int* __restrict__ st, const int* __restrict__ ns,
                            double* __restrict__ rows, int ncosmo, int nloc, int per, int nout, int has_pk) {
  const int RL = 21 * nout + 2;
  const long total = (long)ncosmo * per * RL;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int q = (int)(i % RL);
    const long row = i / RL;
    const int j = (int)(row % per), c = (int)(row / per);
    double v = 0.0;
    if (j < nloc) {
      const size_t m = (size_t)c * nloc + j;
      if (q < 20 * nout) v = y[m * 20 * nout + q];
      else if (q < 21 * nout) v = has_pk ? pk[m * nout + (q - 20 * nout)] : 0.0;
      else if (q == 21 * nout) v = (double)st[m];
      else v = (double)ns[m];
    }
    rows[i] = v;
  }
}
// undo the deal: gathered[r][c][j][RL] -> y_all[c][j W + r][...]
__global__ void k_unpack_rows(const double* __restrict__ g, double* __restrict__ y, double* __restrict__ pk, int* __restrict__ st, int* __restrict__ ns,
                              int ncosmo, int nk, int per, int nout, int world, int has_pk) {
  const int RL = 21 * nout + 2;
  const long total = (long)world * ncosmo * per * RL;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int q = (int)(i % RL);
    long row = i / RL;
    const int j = (int)(row % per); row /= per;
    const int c = (int)(row % ncosmo), r = (int)(row / ncosmo);
    const long gk = (long)j * world + r;
    if (gk >= nk) continue;
    const size_t m = (size_t)c * nk + gk;
    const double v = g[i];
    if (q < 20 * nout) y[m * 20 * nout + q] = v;
    else if (q < 21 * nout) { if (has_pk) pk[m * nout + (q - 20 * nout)] = v; }
    else if (q == 21 * nout) st[m] = (int)v;
    else ns[m] = (int)v;
  }
}

static inline size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

extern "C" {

int deb_nccl_unique_id(void* id128) {
  if (!id128) return DEB_E_ARG;
  int rc = nccl_load();
  if (rc) return rc;
  ncclUniqueId id;
  NCCL_TRY(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, 128);
  return DEB_OK;
}

int deb_comm_create(int32_t world, int32_t rank, const void* id128, deb_comm** out) {
  if (!out || !id128 || world < 1 || rank < 0 || rank >= world) return DEB_E_ARG;
  *out = nullptr;
  int rc = nccl_load();
  if (rc) return rc;
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  deb_comm* c = new deb_comm();
  c->world = world; c->rank = rank; c->nccl = nullptr;
  ncclResult_t r = g_nccl.CommInitRank(&c->nccl, world, id, rank);
  if (r != 0) { fprintf(stderr, "[discoeb_b200] ncclCommInitRank failed: %s\n", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); delete c; return DEB_E_CUDA; }
  *out = c;
  return DEB_OK;
}

int deb_comm_create_on(int32_t device, int32_t world, int32_t rank, const void* id128, deb_comm** out) {
  CUDA_TRY(cudaSetDevice(device));
  return deb_comm_create(world, rank, id128, out);
}

void deb_comm_destroy(deb_comm* c) {
  if (!c) return;
  if (c->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl);
  if (c->e0) cudaEventDestroy(c->e0);
  if (c->e1) cudaEventDestroy(c->e1);
  if (c->st) cudaStreamDestroy(c->st);
  if (c->arena) cudaFree(c->arena);
  delete c;
}

size_t deb_sharded_workspace_bytes(const deb_dims* d, int32_t world) {
  if (!d || world < 1) return 0;
  const size_t per = ((size_t)d->nk + world - 1) / world, nc = d->ncosmo, nout = d->nout, RL = 21 * nout + 2;
  deb_dims loc = *d;
  loc.nk = (int32_t)per;
  return al256(per * 8) + al256(nc * per * nout * 20 * 8) + al256(nc * per * nout * 8) + 3 * al256(nc * per * 4)
       + al256(nc * per * RL * 8) + al256((size_t)world * nc * per * RL * 8) + al256(deb_workspace_bytes(&loc)) + 256;
}

int deb_evolve_sharded_f64(deb_comm* comm, int32_t world, int32_t rank, const deb_dims* dims, const deb_ctrl* ctrl,
                           const double* scalars, const double* tables, const double* kmodes, const double* aexp_out,
                           double* y_all, double* pk_all, double* tau_out, int32_t* status_all, int32_t* nsteps_all,
                           void* workspace, size_t workspace_bytes, int32_t gather,
                           double* const* y_peer, double* const* pk_peer, int32_t* const* st_peer, int32_t* const* ns_peer,
                           void* stream) {
  if (!dims || !ctrl || world < 1 || world > 8 || rank < 0 || rank >= world) return DEB_E_ARG;
  if (dims->return_full || dims->k_per_cosmo || dims->ntan || dims->batch_size) return DEB_E_UNSUPPORTED;
  if (!y_all || !status_all || !nsteps_all || !workspace || !kmodes) return DEB_E_ARG;
  if (workspace_bytes < deb_sharded_workspace_bytes(dims, world)) return DEB_E_WORKSPACE;
  if (gather == 0 && world > 1 && (!comm || comm->world != world || comm->rank != rank)) return DEB_E_ARG;
  if (gather == 1 && (!y_peer || !st_peer || !ns_peer || (dims->power_idx >= 0 && !pk_peer))) return DEB_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int nk = dims->nk, nloc = share(nk, world, rank), per = (nk + world - 1) / world, nc = dims->ncosmo, nout = dims->nout;
  const int has_pk = dims->power_idx >= 0 && pk_all;
  if (world == 1) {       // nothing to deal or gather: the plain entry writes the full-size buffers directly
    deb_dims one = *dims;
    if (!has_pk) one.power_idx = -1;
    char* w1 = (char*)workspace;
    int32_t* na1 = (int32_t*)w1;
    void* ews1 = w1 + al256((size_t)nc * nk * 4);
    return deb_evolve_f64(&one, ctrl, scalars, tables, kmodes, aexp_out, y_all, pk_all, tau_out, status_all, nsteps_all, na1, ews1,
                          workspace_bytes - al256((size_t)nc * nk * 4), stream);
  }
  const size_t RL = 21 * (size_t)nout + 2;
  char* w = (char*)workspace;
  double* kloc = (double*)w; w += al256((size_t)per * 8);
  double* yloc = (double*)w; w += al256((size_t)nc * per * nout * 20 * 8);
  double* pkloc = (double*)w; w += al256((size_t)nc * per * nout * 8);
  int32_t* stloc = (int32_t*)w; w += al256((size_t)nc * per * 4);
  int32_t* nsloc = (int32_t*)w; w += al256((size_t)nc * per * 4);
  int32_t* naloc = (int32_t*)w; w += al256((size_t)nc * per * 4);
  double* rows = (double*)w; w += al256((size_t)nc * per * RL * 8);
  double* gath = (double*)w; w += al256((size_t)world * nc * per * RL * 8);
  void* ews = (void*)w;
  deb_dims loc = *dims;
  loc.nk = nloc;
  if (!has_pk) loc.power_idx = -1;
  const size_t ews_bytes = al256(deb_workspace_bytes(&loc));
  if (nloc > 0) {
    k_take_modes<<<(nloc + 127) / 128, 128, 0, st>>>(kmodes, kloc, nloc, world, rank);
    CUDA_TRY(cudaGetLastError());
  }
  if (gather == 1) {
    // epilogue = gather: the kernels write rows kidx W + rank of every rank's full-size buffers
    if (nloc > 0) {
      int rc = deb_evolve_peer_f64(&loc, ctrl, scalars, tables, kloc, aexp_out, yloc, pkloc, tau_out, stloc, nsloc, naloc, ews, ews_bytes, stream,
                                   world, world, rank, nk, y_peer, pk_peer, st_peer, ns_peer);
      if (rc) return rc;
    }
    return DEB_OK;
  }
  if (nloc > 0) {
    int rc = deb_evolve_f64(&loc, ctrl, scalars, tables, kloc, aexp_out, yloc, pkloc, tau_out, stloc, nsloc, naloc, ews, ews_bytes, stream);
    if (rc) return rc;
  }
  const long total = (long)nc * per * (long)RL;
  int blocks = (int)((total + 255) / 256); if (blocks > 1184) blocks = 1184; if (blocks < 1) blocks = 1;
  k_pack_rows<<<blocks, 256, 0, st>>>(yloc, pkloc, stloc, nsloc, rows, nc, nloc, per, nout, has_pk);
  CUDA_TRY(cudaGetLastError());
  if (world > 1) NCCL_TRY(g_nccl.AllGather(rows, gath, (size_t)total, ncclFloat64, comm->nccl, st));
  else CUDA_TRY(cudaMemcpyAsync(gath, rows, (size_t)total * 8, cudaMemcpyDeviceToDevice, st));
  const long gtotal = total * world;
  blocks = (int)((gtotal + 255) / 256); if (blocks > 1184) blocks = 1184;
  k_unpack_rows<<<blocks, 256, 0, st>>>(gath, y_all, pk_all, status_all, nsteps_all, nc, nk, per, nout, world, has_pk);
  CUDA_TRY(cudaGetLastError());
  return DEB_OK;
}

// host-buffer convenience (gather = 0): H2D of the inputs, the sharded solve + NCCL all-gather, D2H of the full-size results
int deb_evolve_sharded_host_f64(deb_comm* comm, int32_t device, const deb_dims* dims, const deb_ctrl* ctrl, const double* scalars,
                                const double* tables, const double* kmodes, const double* aexp_out, double* y_all, double* pk_all,
                                double* tau_out, int32_t* status_all, int32_t* nsteps_all, float* elapsed_ms) {
  if (!comm || !dims || !ctrl || !scalars || !tables || !kmodes || !aexp_out || !y_all || !tau_out || !status_all || !nsteps_all) return DEB_E_ARG;
  CUDA_TRY(cudaSetDevice(device));
  const size_t nc = dims->ncosmo, nk = dims->nk, nout = dims->nout, tl = deb_table_len(dims);
  const bool pk = dims->power_idx >= 0 && pk_all;
  const size_t b_sc = al256(nc * DEB_NSCAL * 8), b_tb = al256(nc * tl * 8), b_k = al256(nk * 8), b_a = al256(nout * 8),
               b_y = al256(nc * nk * nout * 20 * 8), b_pk = al256(nc * nk * nout * 8), b_tau = al256(nc * nout * 8), b_i = al256(nc * nk * 4),
               b_ws = al256(deb_sharded_workspace_bytes(dims, comm->world));
  int rc = DEB_OK;
  const size_t need = b_sc + b_tb + b_k + b_a + b_y + b_pk + b_tau + 2 * b_i + b_ws;
  if (!comm->st && (cudaStreamCreate(&comm->st) != cudaSuccess || cudaEventCreate(&comm->e0) != cudaSuccess ||
                    cudaEventCreate(&comm->e1) != cudaSuccess)) return DEB_E_CUDA;
  if (need > comm->arena_bytes) {
    if (comm->arena) { cudaStreamSynchronize(comm->st); cudaFree(comm->arena); comm->arena = nullptr; comm->arena_bytes = 0; }
    if (cudaMalloc((void**)&comm->arena, need) != cudaSuccess) return DEB_E_CUDA;
    comm->arena_bytes = need;
    // (a fresh arena starts zeroed: the workspace must never present a stale work-list header)
    if (cudaMemsetAsync(comm->arena, 0, need, comm->st) != cudaSuccess) return DEB_E_CUDA;
  }
  // the workspace sits at the FRONT of the arena (fixed offset whatever the call shape, like ctx_evolve's)
  char* d = comm->arena;
  cudaStream_t st = comm->st;
  cudaEvent_t e0 = comm->e0, e1 = comm->e1;
  if (rc == DEB_OK) {
    char* p_ws = d; char* p_sc = p_ws + b_ws; char* p_tb = p_sc + b_sc; char* p_k = p_tb + b_tb; char* p_a = p_k + b_k; char* p_y = p_a + b_a;
    char* p_pk = p_y + b_y; char* p_tau = p_pk + b_pk; char* p_st = p_tau + b_tau; char* p_ns = p_st + b_i;
    cudaMemcpyAsync(p_sc, scalars, nc * DEB_NSCAL * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(p_tb, tables, nc * tl * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(p_k, kmodes, nk * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(p_a, aexp_out, nout * 8, cudaMemcpyHostToDevice, st);
    cudaEventRecord(e0, st);
    rc = deb_evolve_sharded_f64(comm, comm->world, comm->rank, dims, ctrl, (double*)p_sc, (double*)p_tb, (double*)p_k, (double*)p_a, (double*)p_y,
                                pk ? (double*)p_pk : nullptr, (double*)p_tau, (int32_t*)p_st, (int32_t*)p_ns, p_ws, b_ws, 0, nullptr, nullptr, nullptr,
                                nullptr, (void*)st);
    cudaEventRecord(e1, st);
    if (rc == DEB_OK) {
      cudaMemcpyAsync(y_all, p_y, nc * nk * nout * 20 * 8, cudaMemcpyDeviceToHost, st);
      if (pk) cudaMemcpyAsync(pk_all, p_pk, nc * nk * nout * 8, cudaMemcpyDeviceToHost, st);
      cudaMemcpyAsync(tau_out, p_tau, nc * nout * 8, cudaMemcpyDeviceToHost, st);
      cudaMemcpyAsync(status_all, p_st, nc * nk * 4, cudaMemcpyDeviceToHost, st);
      cudaMemcpyAsync(nsteps_all, p_ns, nc * nk * 4, cudaMemcpyDeviceToHost, st);
    }
    if (cudaStreamSynchronize(st) != cudaSuccess && rc == DEB_OK) rc = DEB_E_CUDA;
    if (rc == DEB_OK && elapsed_ms) cudaEventElapsedTime(elapsed_ms, e0, e1);
  }
  return rc;
}

}  // extern "C"
