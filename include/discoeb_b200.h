/* discoeb_b200.h -- C-ABI of the B200-native Einstein-Boltzmann hot path.
 *
 * One entry point replaces the reference's per-mode solve
 *   evolve_perturbations -> jax.vmap(evolve_one_mode) -> diffrax.diffeqsolve(Rodas5Transformed)
 *   (/root/reference/src/discoeb/perturbations.py:926-997, :728-781;
 *    ode_integrators_stiff.py:694-843)
 * for every (cosmology, k) pair of a batch, including the prologue (start time :630-681,
 * adiabatic initial conditions :526-627) and the epilogue (output conversion :374-523,
 * get_power :1101-1123).  The reference has no FFI of its own: its seam is the Python
 * keyword API on the `param` dict, mirrored by disco-eb_b200/discoeb_b200/perturbations.py,
 * which binds this library with ctypes (INTEGRATION.md shows the jax.ffi binding).
 *
 * Conventions: plain pointers and sizes, float64 everywhere, row-major, no allocation and no
 * synchronisation inside the *_f64 device entry (asynchronous on `stream`), re-entrant, no
 * global state.  Return value 0 or a negative deb_status code; per-mode solver outcomes go to
 * status[] (0 ok, 1 max_steps exhausted, 2 non-finite, 3 not processed: the value every entry holds until
 * a warp has integrated the mode), mirroring diffrax's RESULTS.
 */
#ifndef DISCOEB_B200_H
#define DISCOEB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEB_ABI_VERSION 2
#define DEB_NSCAL 24        /* doubles per cosmology in `scalars` */
#define DEB_NSPLINE 7       /* splines per cosmology in `tables`  */
#define DEB_NFIELD 20       /* output fields (perturbations.py:511-521) */

/* index of each scalar inside one cosmology's DEB_NSCAL block */
enum deb_scalar {
  DEB_S_OMEGAM = 0, DEB_S_OMEGAB, DEB_S_OMEGADE, DEB_S_OMEGAK,
  DEB_S_GRHOM, DEB_S_GRHOG, DEB_S_GRHOR, DEB_S_NEFF, DEB_S_NMNU, DEB_S_AMNU,
  DEB_S_W0, DEB_S_WA, DEB_S_CS2DE, DEB_S_YHE, DEB_S_H0, DEB_S_TAUMIN,
  DEB_S_AS, DEB_S_NS, DEB_S_KP
};

/* order of the splines inside one cosmology's table block.  Every spline is stored as
 * x[n], y[n], S[n] (knots, values, second derivatives -- spline_interpolation.py:111-113).
 * Splines 0,1 and 4,5,6 have nth knots, splines 2,3 have nnu knots; 0 and 1 must share
 * their knots (they do in the reference, background.py:254-255).
 * Block length = 3*(5*nth + 2*nnu) doubles. */
enum deb_spline {
  DEB_T_CS2A_OF_LOGA = 0, DEB_T_XE_OF_LOGA, DEB_T_LOGRHONU_OF_LOGA, DEB_T_LOGPNU_OF_LOGA,
  DEB_T_A_OF_TAU, DEB_T_XE_OF_TAU, DEB_T_TAU_OF_A
};

typedef struct deb_dims {
  int32_t ncosmo;        /* cosmologies in the batch                                  */
  int32_t nk;            /* k-modes per cosmology                                     */
  int32_t nout;          /* output times (ascending aexp_out)                         */
  int32_t ntan;          /* forward-tangent directions (0 for the primal entries)     */
  int32_t lmaxg, lmaxgp, lmaxr, lmaxnu;   /* hierarchy cut-offs, each >= 3            */
  int32_t nqmax;         /* massive-neutrino momentum bins (3,4,5)                    */
  int32_t nth, nnu;      /* knots of the thermo / neutrino splines                    */
  int32_t max_steps;     /* attempted (accepted+rejected) step budget per mode        */
  int32_t return_full;   /* 0: 20 output fields, 1: raw state vector (return_full)    */
  int32_t k_per_cosmo;   /* 0: kmodes[nk] shared, 1: kmodes[ncosmo*nk]                */
  int32_t power_idx;     /* >=0: also write P(k) of this field to pk_out (get_power)  */
  int32_t batch_size;    /* 0: every mode has its own adaptive step sequence (evolve_perturbations);
                            B > 0: consecutive groups of B k-modes share ONE step size and start time, the
                            numerics of evolve_perturbations_batched / Rodas5Batched (perturbations.py:786-922,
                            ode_integrators_stiff.py:846-1010); needs nk % B == 0, B <= 64, ntan == 0 */
} deb_dims;

typedef struct deb_ctrl {
  double rtol, atol;
  double pcoeff, icoeff, dcoeff;
  double factormax, factormin;
  double safety;         /* diffrax default 0.9 */
} deb_ctrl;

enum deb_status {
  DEB_OK = 0, DEB_E_ARG = -1, DEB_E_UNSUPPORTED = -2, DEB_E_WORKSPACE = -3,
  DEB_E_CUDA = -4, DEB_E_NODEVICE = -5
};

/* number of state variables n = 7+(lg+1)+(lp+1)+(lr+1)+nq*(lnu+1)+2 (perturbations.py:739) */
int32_t deb_nvar(const deb_dims* dims);
/* doubles per cosmology in `tables` */
size_t deb_table_len(const deb_dims* dims);
/* bytes of device scratch deb_evolve_f64 wants: a work-queue counter, one start-time root per cosmology and -- optional,
 * used when the buffer is at least this large -- a list of the modes ordered by the step counts of the PREVIOUS call
 * with the same buffer and shape.  Callers that keep the buffer between calls (an MCMC or emulator loop evaluating
 * near-identical cosmologies) get longest-first scheduling from the second call on; results do not depend on it.
 * The contents need no initialisation; 256 + 8 ncosmo bytes are the hard minimum. */
size_t deb_workspace_bytes(const deb_dims* dims);
const char* deb_strerror(int code);
int32_t deb_abi_version(void);
int32_t deb_device_count(void);

/* Device-pointer entry.  All array arguments are DEVICE pointers.
 *   scalars   [ncosmo, DEB_NSCAL]
 *   tables    [ncosmo, deb_table_len]
 *   kmodes    [nk] or [ncosmo, nk]
 *   aexp_out  [nout]                      ascending
 *   y_out     [ncosmo, nk, nout, 20|n]
 *   pk_out    [ncosmo, nk, nout] or NULL  (requires power_idx >= 0)
 *   tau_out   [ncosmo, nout]              tau_of_a(aexp_out), returned like param['tau_out']
 *   status    [ncosmo, nk]   nsteps [ncosmo, nk] attempted steps   naccept [ncosmo,nk] or NULL
 *   stream    a cudaStream_t passed as void* (NULL = default stream)
 */
int deb_evolve_f64(const deb_dims* dims, const deb_ctrl* ctrl,
                   const double* scalars, const double* tables, const double* kmodes,
                   const double* aexp_out,
                   double* y_out, double* pk_out, double* tau_out,
                   int32_t* status, int32_t* nsteps, int32_t* naccept,
                   void* workspace, size_t workspace_bytes, void* stream);

/* Forward tangents (jax.jvp / jax.jacfwd of evolve_perturbations in the reference: the minimal
 * notebook cell 14, the Fisher notebook cell 7).  For each of dims->ntan
 * directions the caller supplies the tangent of every input, laid out like the primal input
 * with a leading [ntan] axis:
 *   d_scalars [ntan, ncosmo, DEB_NSCAL]      d_tables [ntan, ncosmo, deb_table_len]
 * (x, y and S tangents of every spline; zero where an array does not depend on the parameter),
 * and receives the exact derivative of the discrete solve -- tangents flow through the start
 * time, the initial conditions, every Rosenbrock stage including W = I/(gamma dt) - J, the step
 * end points, the output interpolation and conversion, while accept/reject, the PID factor,
 * spline intervals and bisection branches follow the primal run (diffrax semantics):
 *   dy_out   [ntan, ncosmo, nk, nout, 20|n]   dpk_out [ntan, ncosmo, nk, nout] or NULL
 *   dtau_out [ntan, ncosmo, nout]             tangent of tau_of_a(aexp_out)
 * d_kmodes [ntan, nk] (or [ntan, ncosmo, nk] with k_per_cosmo) is the tangent of the wavenumbers
 * themselves, for callers whose k grid depends on a parameter (kmin, kmax scaled by h); NULL = 0.
 * The primal outputs are those of deb_evolve_f64 for the same step sequence.  aexp_out carries
 * no tangent.  With ntan == 0 this is deb_evolve_f64. */
int deb_evolve_tangent_f64(const deb_dims* dims, const deb_ctrl* ctrl,
                           const double* scalars, const double* tables, const double* kmodes,
                           const double* aexp_out, const double* d_scalars, const double* d_tables,
                           const double* d_kmodes,
                           double* y_out, double* dy_out, double* pk_out, double* dpk_out,
                           double* tau_out, double* dtau_out,
                           int32_t* status, int32_t* nsteps, int32_t* naccept,
                           void* workspace, size_t workspace_bytes, void* stream);

/* Host-pointer entry: same arguments with HOST buffers; copies inputs to device `device`,
 * runs deb_evolve_f64, copies results back and synchronises.  If kernel_ms is non-NULL it
 * receives the device time of the solve kernel alone (CUDA events). */
int deb_evolve_host_f64(const deb_dims* dims, const deb_ctrl* ctrl,
                        const double* scalars, const double* tables, const double* kmodes,
                        const double* aexp_out,
                        double* y_out, double* pk_out, double* tau_out,
                        int32_t* status, int32_t* nsteps, int32_t* naccept,
                        int32_t device, float* kernel_ms);

/* Host-pointer entry with an explicit context.  A deb_ctx owns a stream, two events, one device
 * arena and one pinned host arena on `device` (grow-only), so that repeated calls allocate
 * nothing: inputs are packed into the pinned arena and cross PCIe/C2C in ONE copy, all results
 * come back in ONE copy.  deb_evolve_host_f64 above uses a context cached per (host thread,
 * device); deb_host_cache_release() frees the calling thread's cached contexts.  A context must
 * not be used by two threads at once. */
typedef struct deb_ctx deb_ctx;
int deb_ctx_create(int32_t device, deb_ctx** out);
void deb_ctx_destroy(deb_ctx* ctx);
int deb_ctx_evolve_host_f64(deb_ctx* ctx, const deb_dims* dims, const deb_ctrl* ctrl,
                            const double* scalars, const double* tables, const double* kmodes,
                            const double* aexp_out,
                            double* y_out, double* pk_out, double* tau_out,
                            int32_t* status, int32_t* nsteps, int32_t* naccept, float* kernel_ms);
void deb_host_cache_release(void);
/* host-buffer forms of deb_evolve_tangent_f64 (explicit context / per-thread cached context) */
int deb_ctx_evolve_tangent_host_f64(deb_ctx* ctx, const deb_dims* dims, const deb_ctrl* ctrl,
                                    const double* scalars, const double* tables, const double* kmodes,
                                    const double* aexp_out, const double* d_scalars, const double* d_tables,
                                    const double* d_kmodes,
                                    double* y_out, double* dy_out, double* pk_out, double* dpk_out,
                                    double* tau_out, double* dtau_out,
                                    int32_t* status, int32_t* nsteps, int32_t* naccept, float* kernel_ms);
int deb_evolve_tangent_host_f64(const deb_dims* dims, const deb_ctrl* ctrl,
                                const double* scalars, const double* tables, const double* kmodes,
                                const double* aexp_out, const double* d_scalars, const double* d_tables,
                                const double* d_kmodes,
                                double* y_out, double* dy_out, double* pk_out, double* dpk_out,
                                double* tau_out, double* dtau_out,
                                int32_t* status, int32_t* nsteps, int32_t* naccept,
                                int32_t device, float* kernel_ms);

/* Debug/validation entry (host pointers): ONE attempted Rodas5 step per mode from a given
 * state.  t0[nmodes], t1[nmodes], y0[nmodes, n] -> y1[nmodes, n], yerr[nmodes, n].
 * Modes index (cosmology, k) like deb_evolve_f64.  Used by the parity tests to compare the
 * structured solve with the oracle's dense LU stage by stage. */
int deb_debug_step_host_f64(const deb_dims* dims, const double* scalars, const double* tables,
                            const double* kmodes, const double* t0, const double* t1,
                            const double* y0, double* y1, double* yerr, int32_t device);

/* Debug/validation entry (host pointers): prologue only.  tau_start[ncosmo,nk] and the
 * adiabatic initial state y0[ncosmo,nk,n]. */
int deb_debug_ics_host_f64(const deb_dims* dims, const double* scalars, const double* tables,
                           const double* kmodes, const double* aexp_out,
                           double* tau_start, double* y0, int32_t device);

/* Debug/validation entry (host pointers): integrate every mode along a PRESCRIBED step sequence
 * (rp_tnext[mode, s] = end of attempted step s, rp_keep[mode, s] = accepted?, rp_n[mode] steps;
 * rp_stride = row length) instead of the PID controller.  The adaptive step sequence of the
 * reference algorithm is chaotic under round-off (DESIGN.md "Parity"), so the parity tests replay
 * the oracle's own sequence to compare whole trajectories at round-off level. */
int deb_debug_replay_host_f64(const deb_dims* dims, const deb_ctrl* ctrl, const double* scalars,
                              const double* tables, const double* kmodes, const double* aexp_out,
                              const double* rp_tnext, const int32_t* rp_keep, const int32_t* rp_n,
                              int32_t rp_stride, double* y_out, int32_t* nsteps, int32_t device);

/* Debug/validation entry (host pointers): deb_debug_replay_host_f64 with tangents.  rp_dtnext
 * [ntan, ncosmo*nk, rp_stride] holds the tangent of every prescribed step end. */
int deb_debug_replay_tangent_host_f64(const deb_dims* dims, const deb_ctrl* ctrl, const double* scalars,
                                      const double* tables, const double* kmodes, const double* aexp_out,
                                      const double* d_scalars, const double* d_tables, const double* d_kmodes,
                                      const double* rp_tnext, const double* rp_dtnext,
                                      const int32_t* rp_keep, const int32_t* rp_n, int32_t rp_stride,
                                      double* y_out, double* dy_out, double* dtau_out, int32_t* nsteps,
                                      int32_t device);

/* ---- table production --------------------------------------------------------------------
 * Replaces, for a batch of cosmologies, what the reference's
 *   evolve_background(param, thermo_module='RECFAST', num_thermo)      /root/reference/src/discoeb/background.py:191-255
 *     -> setup_background_evolution                                     background.py:140-188
 *     -> evaluate_thermo / compute_thermal_history / solve_ionization   thermodynamics_recfast.py:124-500
 *        (GRKT4.step ode_integrators_stiff.py:77-215 under diffrax's PID loop)
 *     -> spline_interpolation(...) constructors                         spline_interpolation.py:8-113
 * leaves in `param` for evolve_perturbations: one CTA per cosmology writes scalars[DEB_NSCAL] (deb_scalar order; slots
 * 19..22 = taumax, Omegamnu, Tcmb, mnu) and tables[deb_table_len] (deb_spline order, nnu = 512) in exactly the layout
 * deb_evolve_f64 consumes, so a cosmology batch goes from parameters to P(k) without leaving the device.
 *   bg_in [ncosmo, 16]: Omegam, Omegab, Omegak, w_DE_0, w_DE_a, cs2_DE, H0, Tcmb, YHe, Neff, Nmnu, mnu, A_s, n_s, k_p, (pad)
 * All pointers of deb_background_f64 are DEVICE pointers; no allocation, no synchronisation, asynchronous on `stream`.
 * nth = num_thermo (16..1024).  workspace: deb_background_workspace_bytes(ncosmo, nth) bytes, contents undefined. */
size_t deb_background_workspace_bytes(int32_t ncosmo, int32_t nth);
size_t deb_background_nin(void);
int deb_background_f64(int32_t ncosmo, int32_t nth, const double* bg_in, double* scalars, double* tables,
                       void* workspace, size_t workspace_bytes, void* stream);
/* host-buffer form: H2D of bg_in, the kernel, D2H of scalars and tables; kernel_ms (optional) = device time */
int deb_background_host_f64(int32_t device, int32_t ncosmo, int32_t nth, const double* bg_in, double* scalars,
                            double* tables, float* kernel_ms);

/* The rest of what evolve_background leaves in `param` (not read by evolve_perturbations; for callers of the reference
 * that use the thermal history or the visibility function): background.py:238-246 (x_HII, x_HeII, x_HeIII fractions, c_s^2,
 * T_m), :250-251 (second derivatives of the a c_s^2 and a T_m splines over tau), :167 (log pseudo-pressure table of the
 * massive neutrinos), :306-342 (opacity, optical depth from today, visibility g and its first two derivatives, x_e' from
 * the spline and from RECFAST's own right-hand side, thermodynamics_recfast.py:477).
 *   extras [ncosmo, deb_background_extras_len(nth)], rows of nth doubles in this order:
 *     xeHI, xeHeI, xeHeII, cs2, Tm, xeprime_recf, xeprime, opac, optical_depth, gvis, gvisprime, gvispprime,
 *     S of cs2a_of_tau_spline, S of tempba_of_tau_spline (x = tau, y = aexp cs2 / aexp Tm), then
 *     log ppseudo_nu [512] and its second derivatives [512] on the knots of the logrhonu table.
 * extras = NULL makes the _ex entries identical to the plain ones. */
size_t deb_background_extras_len(int32_t nth);
int deb_background_ex_f64(int32_t ncosmo, int32_t nth, const double* bg_in, double* scalars, double* tables, double* extras,
                          void* workspace, size_t workspace_bytes, void* stream);
int deb_background_host_ex_f64(int32_t device, int32_t ncosmo, int32_t nth, const double* bg_in, double* scalars,
                               double* tables, double* extras, float* kernel_ms);

/* ---- multi-GPU: one k grid dealt round-robin over the ranks of one box ---------------------------
 * The reference has no multi-device code; its vmap over k (perturbations.py:980-987) is what gets sharded: rank r integrates
 * modes r, r + W, ... (cost rises steeply with k) and EVERY rank ends with the full-size y_all[ncosmo, nk, nout, 20],
 * pk_all[ncosmo, nk, nout], status_all / nsteps_all[ncosmo, nk].  dims describes the WHOLE problem (nk = all modes;
 * return_full = 0, k_per_cosmo = 0, ntan = 0, batch_size = 0).  All array arguments are device pointers, the call is
 * asynchronous on `stream`, nothing is allocated.
 *   gather = 0: one ncclAllGather (communicator of deb_comm_create) + a kernel that undoes the deal;
 *   gather = 1: no collective -- the kernels' epilogue stores each mode's row into the full-size buffers of ALL ranks:
 *               y_peer[r] / pk_peer[r] / st_peer[r] / ns_peer[r] (host arrays of W device pointers, r = rank) must be
 *               addresses valid on THIS device for rank r's y_all / pk_all / status_all / nsteps_all (CUDA IPC or
 *               symmetric-memory mappings over NVLink; entry `rank` is the local buffer).  The caller runs a cross-rank
 *               barrier after the stream work before anybody reads. */
typedef struct deb_comm deb_comm;
int deb_nccl_unique_id(void* id128);                       /* rank 0: 128 bytes to hand to every rank */
int deb_comm_create(int32_t world, int32_t rank, const void* id128, deb_comm** out);   /* on the current device; collective */
int deb_comm_create_on(int32_t device, int32_t world, int32_t rank, const void* id128, deb_comm** out);
void deb_comm_destroy(deb_comm* comm);
size_t deb_sharded_workspace_bytes(const deb_dims* dims, int32_t world);
int deb_evolve_sharded_f64(deb_comm* comm, int32_t world, int32_t rank, const deb_dims* dims, const deb_ctrl* ctrl,
                           const double* scalars, const double* tables, const double* kmodes, const double* aexp_out,
                           double* y_all, double* pk_all, double* tau_out, int32_t* status_all, int32_t* nsteps_all,
                           void* workspace, size_t workspace_bytes, int32_t gather,
                           double* const* y_peer, double* const* pk_peer, int32_t* const* st_peer, int32_t* const* ns_peer,
                           void* stream);
/* host-buffer form of deb_evolve_sharded_f64 with gather = 0 (every rank passes the same host inputs) */
int deb_evolve_sharded_host_f64(deb_comm* comm, int32_t device, const deb_dims* dims, const deb_ctrl* ctrl, const double* scalars,
                                const double* tables, const double* kmodes, const double* aexp_out, double* y_all, double* pk_all,
                                double* tau_out, int32_t* status_all, int32_t* nsteps_all, float* elapsed_ms);
/* deb_evolve_f64 whose epilogue also stores every mode's row (20 nout fields, P(k), status, step count) into `npeer`
 * full-size peer buffers: local mode kidx is row kidx * out_mul + out_add of out_nk. */
int deb_evolve_peer_f64(const deb_dims* dims, const deb_ctrl* ctrl, const double* scalars, const double* tables,
                        const double* kmodes, const double* aexp_out, double* y_out, double* pk_out, double* tau_out,
                        int32_t* status, int32_t* nsteps, int32_t* naccept, void* workspace, size_t workspace_bytes, void* stream,
                        int32_t npeer, int32_t out_mul, int32_t out_add, int32_t out_nk,
                        double* const* y_peer, double* const* pk_peer, int32_t* const* st_peer, int32_t* const* ns_peer);

/* ---- spectra epilogues -----------------------------------------------------------------
 * What the reference's host API derives from y right after the solve (/root/reference/src/discoeb/perturbations.py):
 * power_multipoles :1202-1224 (P0, P2, P4), power_Kaiser :1162-1199 (Pkmu[nk, nmu] on the given mu grid),
 * get_power_smoothed :1126-1160 (Ps_delta, Ps_theta: Savitzky-Golay in log-log, sg_coef[sg_window] = the centre weights of
 * util.savgol_filter :407-444; sg_window = 0: none, Kaiser then uses the raw amplitudes) and get_xi_from_P :1065-1098
 * (FFTlog of P_delta -- smoothed if sg_window > 0 -- multipole `ell`; xi, r ascending in r; NULL to skip).
 * y[nk, 20] is the solver output of ONE cosmology and output time; all pointers of deb_spectra_f64 are DEVICE pointers
 * (asynchronous on `stream`, so the call chains behind deb_evolve_f64 without y leaving the GPU). */
size_t deb_spectra_workspace_bytes(int32_t nk);
int deb_spectra_f64(int32_t nk, int32_t nmu, const double* y, const double* kmodes, double As, double ns, double kp, double bias,
                    const double* sg_coef, int32_t sg_window, const double* mu, int32_t ell,
                    double* P0, double* P2, double* P4, double* Pkmu, double* Ps_delta, double* Ps_theta, double* xi, double* r,
                    void* workspace, size_t workspace_bytes, void* stream);
int deb_spectra_host_f64(int32_t device, int32_t nk, int32_t nmu, const double* y, const double* kmodes, double As, double ns, double kp, double bias,
                         const double* sg_coef, int32_t sg_window, const double* mu, int32_t ell,
                         double* P0, double* P2, double* P4, double* Pkmu, double* Ps_delta, double* Ps_theta, double* xi, double* r);

/* Measures the FP64 FMA peak of `device` (dependent-free DFMA streams on every SM) and
 * returns it in TFLOP/s; the roofline denominator bench.py reports against. */
int deb_fp64_peak_tflops(int32_t device, double* tflops, float* sm_clock_mhz);

#ifdef __cplusplus
}
#endif
#endif /* DISCOEB_B200_H */
