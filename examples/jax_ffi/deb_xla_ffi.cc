// deb_xla_ffi.cc -- XLA FFI handlers around deb_evolve_f64 and deb_evolve_tangent_f64 (jax.ffi custom calls).
//
// NOT built by this repository's Makefile: the XLA FFI headers ship with jaxlib, which is not
// installable in the build environment (no wheels, no network).  Where jaxlib exists:
//
//   g++ -O2 -shared -fPIC -std=c++17 -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \
//       -I../../include -I/usr/local/cuda/include deb_xla_ffi.cc -L../discoeb_b200 -ldiscoeb_b200 \
//       -o ../discoeb_b200/libdiscoeb_b200_ffi.so
//
// and register it as shown in INTEGRATION.md section 2.  The handler allocates nothing and does not
// synchronise: it forwards XLA's device buffers and stream to the C-ABI, which is why the ABI has
// that shape.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define DEB_HAVE_XLA_FFI 1
#endif
#endif

#ifdef DEB_HAVE_XLA_FFI
#include <cuda_runtime_api.h>
#include "xla/ffi/api/ffi.h"
#include "discoeb_b200.h"

namespace ffi = xla::ffi;

static ffi::Error DebEvolveImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> scalars, ffi::Buffer<ffi::F64> tables,
                                ffi::Buffer<ffi::F64> kmodes, ffi::Buffer<ffi::F64> aexp_out,
                                ffi::ResultBuffer<ffi::F64> y, ffi::ResultBuffer<ffi::F64> pk,
                                ffi::ResultBuffer<ffi::F64> tau_out, ffi::ResultBuffer<ffi::S32> status,
                                ffi::ResultBuffer<ffi::S32> nsteps, ffi::ResultBuffer<ffi::S32> naccept,
                                ffi::ResultBuffer<ffi::U8> workspace, int32_t lmaxg, int32_t lmaxgp, int32_t lmaxr,
                                int32_t lmaxnu, int32_t nqmax, int32_t nth, int32_t nnu, int32_t max_steps,
                                int32_t return_full, int32_t power_idx, double rtol, double atol, double pcoeff,
                                double icoeff, double dcoeff, double factormax, double factormin) {
  deb_dims d = {};
  auto sd = scalars.dimensions();
  auto kd = kmodes.dimensions();
  d.ncosmo = static_cast<int32_t>(sd[0]);
  d.nk = static_cast<int32_t>(kd[kd.size() - 1]);
  d.k_per_cosmo = kd.size() == 2;
  d.nout = static_cast<int32_t>(aexp_out.dimensions()[0]);
  d.lmaxg = lmaxg; d.lmaxgp = lmaxgp; d.lmaxr = lmaxr; d.lmaxnu = lmaxnu; d.nqmax = nqmax;
  d.nth = nth; d.nnu = nnu; d.max_steps = max_steps; d.return_full = return_full; d.power_idx = power_idx;
  deb_ctrl c = {rtol, atol, pcoeff, icoeff, dcoeff, factormax, factormin, 0.9};
  if (workspace->size_bytes() < deb_workspace_bytes(&d)) return ffi::Error::InvalidArgument("workspace too small");
  int rc = deb_evolve_f64(&d, &c, scalars.typed_data(), tables.typed_data(), kmodes.typed_data(), aexp_out.typed_data(),
                          y->typed_data(), pk->typed_data(), tau_out->typed_data(), status->typed_data(),
                          nsteps->typed_data(), naccept->typed_data(), workspace->untyped_data(), workspace->size_bytes(),
                          stream);
  return rc == 0 ? ffi::Error::Success() : ffi::Error::Internal(deb_strerror(rc));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    DebEvolve, DebEvolveImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
        .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>()
        .Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::U8>>()
        .Attr<int32_t>("lmaxg").Attr<int32_t>("lmaxgp").Attr<int32_t>("lmaxr").Attr<int32_t>("lmaxnu").Attr<int32_t>("nqmax")
        .Attr<int32_t>("nth").Attr<int32_t>("nnu").Attr<int32_t>("max_steps").Attr<int32_t>("return_full")
        .Attr<int32_t>("power_idx")
        .Attr<double>("rtol").Attr<double>("atol").Attr<double>("pcoeff").Attr<double>("icoeff").Attr<double>("dcoeff")
        .Attr<double>("factormax").Attr<double>("factormin"));
// Tangent handler: the jvp rule of the custom_jvp around DebEvolve ("Derivatives").  jax.jacfwd
// batches the rule over directions; with vmap_method="expand_dims" that batch axis is the leading [ntan] axis here.
static ffi::Error DebEvolveTangentImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> scalars, ffi::Buffer<ffi::F64> tables,
                                       ffi::Buffer<ffi::F64> kmodes, ffi::Buffer<ffi::F64> aexp_out,
                                       ffi::Buffer<ffi::F64> d_scalars, ffi::Buffer<ffi::F64> d_tables,
                                       ffi::Buffer<ffi::F64> d_kmodes, ffi::ResultBuffer<ffi::F64> y,
                                       ffi::ResultBuffer<ffi::F64> dy, ffi::ResultBuffer<ffi::F64> pk,
                                       ffi::ResultBuffer<ffi::F64> dpk, ffi::ResultBuffer<ffi::F64> tau_out,
                                       ffi::ResultBuffer<ffi::F64> dtau_out, ffi::ResultBuffer<ffi::S32> status,
                                       ffi::ResultBuffer<ffi::S32> nsteps, ffi::ResultBuffer<ffi::S32> naccept,
                                       ffi::ResultBuffer<ffi::U8> workspace, int32_t lmaxg, int32_t lmaxgp, int32_t lmaxr,
                                       int32_t lmaxnu, int32_t nqmax, int32_t nth, int32_t nnu, int32_t max_steps,
                                       int32_t return_full, int32_t power_idx, double rtol, double atol, double pcoeff,
                                       double icoeff, double dcoeff, double factormax, double factormin) {
  deb_dims d = {};
  auto sd = scalars.dimensions();
  auto kd = kmodes.dimensions();
  d.ncosmo = static_cast<int32_t>(sd[0]);
  d.nk = static_cast<int32_t>(kd[kd.size() - 1]);
  d.k_per_cosmo = kd.size() == 2;
  d.nout = static_cast<int32_t>(aexp_out.dimensions()[0]);
  d.ntan = static_cast<int32_t>(d_scalars.dimensions()[0]);
  d.lmaxg = lmaxg; d.lmaxgp = lmaxgp; d.lmaxr = lmaxr; d.lmaxnu = lmaxnu; d.nqmax = nqmax;
  d.nth = nth; d.nnu = nnu; d.max_steps = max_steps; d.return_full = return_full; d.power_idx = power_idx;
  deb_ctrl c = {rtol, atol, pcoeff, icoeff, dcoeff, factormax, factormin, 0.9};
  if (workspace->size_bytes() < deb_workspace_bytes(&d)) return ffi::Error::InvalidArgument("workspace too small");
  int rc = deb_evolve_tangent_f64(&d, &c, scalars.typed_data(), tables.typed_data(), kmodes.typed_data(), aexp_out.typed_data(),
                                  d_scalars.typed_data(), d_tables.typed_data(), d_kmodes.typed_data(), y->typed_data(),
                                  dy->typed_data(), pk->typed_data(), dpk->typed_data(), tau_out->typed_data(),
                                  dtau_out->typed_data(), status->typed_data(), nsteps->typed_data(), naccept->typed_data(),
                                  workspace->untyped_data(), workspace->size_bytes(), stream);
  return rc == 0 ? ffi::Error::Success() : ffi::Error::Internal(deb_strerror(rc));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    DebEvolveTangent, DebEvolveTangentImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
        .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>()
        .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>()
        .Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::U8>>()
        .Attr<int32_t>("lmaxg").Attr<int32_t>("lmaxgp").Attr<int32_t>("lmaxr").Attr<int32_t>("lmaxnu").Attr<int32_t>("nqmax")
        .Attr<int32_t>("nth").Attr<int32_t>("nnu").Attr<int32_t>("max_steps").Attr<int32_t>("return_full")
        .Attr<int32_t>("power_idx")
        .Attr<double>("rtol").Attr<double>("atol").Attr<double>("pcoeff").Attr<double>("icoeff").Attr<double>("dcoeff")
        .Attr<double>("factormax").Attr<double>("factormin"));
#else
// Built without jaxlib: keep the translation unit non-empty and say why.
extern "C" const char* deb_xla_ffi_unavailable(void) {
  return "deb_xla_ffi.cc was compiled without xla/ffi/api/ffi.h (jaxlib headers not found)";
}
#endif
