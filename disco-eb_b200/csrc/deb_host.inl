// deb_host.inl -- argument validation and launch-problem set-up shared by the CUDA library
// (deb_kernels.cu) and the CPU emulation harness of the test-suite (tests/emu/deb_emu.cpp).
#pragma once
#include <string.h>

static inline int deb_nvar_impl(const deb_dims* d) {
  return 7 + (d->lmaxg + 1) + (d->lmaxgp + 1) + (d->lmaxr + 1) + d->nqmax * (d->lmaxnu + 1) + 2;
}

// returns DEB_OK or a negative code; fills everything of P that does not depend on buffers
static inline int fill_problem(const deb_dims* d, const deb_ctrl* c, deb::Problem* P) {
  if (!d || !c || !P) return DEB_E_ARG;
  if (d->ncosmo < 1 || d->nk < 1 || d->nout < 1 || d->max_steps < 1) return DEB_E_ARG;
  if (d->nth < 2 || d->nnu < 2) return DEB_E_ARG;
  if (d->ntan < 0 || d->batch_size < 0) return DEB_E_ARG;
  if (d->batch_size > 0 && (d->nk % d->batch_size != 0 || d->batch_size > 64 || d->ntan != 0)) return DEB_E_ARG;
  if (d->lmaxg < 3 || d->lmaxgp < 3 || d->lmaxr < 3 || d->lmaxnu < 3) return DEB_E_UNSUPPORTED;
  if (d->lmaxg >= deb::LMAXCAP || d->lmaxgp >= deb::LMAXCAP || d->lmaxr >= deb::LMAXCAP || d->lmaxnu >= deb::LMAXCAP)
    return DEB_E_UNSUPPORTED;
  if (d->nqmax < 3 || d->nqmax > deb::NQMAX) return DEB_E_UNSUPPORTED;
  if (d->power_idx >= DEB_NFIELD) return DEB_E_ARG;
  if (!(c->rtol > 0.0) || !(c->atol >= 0.0)) return DEB_E_ARG;
  memset(P, 0, sizeof(*P));
  P->ncosmo = d->ncosmo; P->nk = d->nk; P->nout = d->nout;
  P->lmaxg = d->lmaxg; P->lmaxgp = d->lmaxgp; P->lmaxr = d->lmaxr; P->lmaxnu = d->lmaxnu; P->nq = d->nqmax;
  P->nth = d->nth; P->nnu = d->nnu;
  P->max_steps = d->max_steps; P->return_full = d->return_full; P->k_per_cosmo = d->k_per_cosmo;
  P->power_idx = d->power_idx;
  P->ntan = d->ntan;
  P->batch_size = d->batch_size;
  P->n = deb_nvar_impl(d);
  P->np = (P->n + 1) & ~1;
  P->nh = 17 + 3 * d->nqmax;
  P->nch = 3 + d->nqmax;
  P->ig = 7; P->igp = 7 + (d->lmaxg + 1); P->ir = 9 + d->lmaxg + d->lmaxgp; P->iq0 = 10 + d->lmaxg + d->lmaxgp + d->lmaxr;
  if (P->n > 32 * 12) return DEB_E_UNSUPPORTED;
  const double order = 5.0;    // Rodas5Transformed.order (ode_integrators_stiff.py:702-703)
  P->rtol = c->rtol; P->atol = c->atol;
  P->c1 = (c->icoeff + c->pcoeff + c->dcoeff) / order;
  P->c2 = -(c->pcoeff + 2.0 * c->dcoeff) / order;
  P->c3 = c->dcoeff / order;
  P->factormax = c->factormax; P->factormin = c->factormin; P->safety = c->safety;
  return DEB_OK;
}

static inline const char* deb_strerror_impl(int code) {
  switch (code) {
    case DEB_OK: return "ok";
    case DEB_E_ARG: return "invalid argument";
    case DEB_E_UNSUPPORTED: return "unsupported configuration (need lmax* in [3,95], nqmax in [3,5], n <= 384)";
    case DEB_E_WORKSPACE: return "workspace too small";
    case DEB_E_CUDA: return "CUDA runtime error";
    case DEB_E_NODEVICE: return "no CUDA device";
    default: return "unknown error";
  }
}
