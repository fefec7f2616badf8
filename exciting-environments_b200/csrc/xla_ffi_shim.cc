// xla_ffi_shim.cc -- XLA typed-FFI handlers over the C ABI, so that the reference's host code can stay JAX and call the
// sm_100a kernels with jax.ffi.ffi_call (BASELINE.json north_star).  Compiled only where the XLA FFI headers exist
// (`python -c "import jax; print(jax.ffi.include_dir())"`); this image ships no jax, so the file is a guarded no-op
// here and cannot be exercised -- it is kept small and mechanical on purpose.  See INTEGRATION.md.
//
// Buffers arrive in the reference's pytree-leaf order: D physical-state leaves (B,), then the action
// (B, A); results: obs (B, O'), D new state leaves (+ NODE solver_state leaves for Tsit5).  The plan handle created with
// exenv_plan_create() travels as an int64 attribute.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define EXENV_HAVE_XLA_FFI 1
#endif
#endif

#ifdef EXENV_HAVE_XLA_FFI
#include <cuda_runtime.h>

#include <string>

#include "../../include/exenv_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error fail(int rc) {
  char msg[512];
  exenv_last_error(msg, sizeof(msg));
  return ffi::Error(rc == EXENV_ERR_INVALID || rc == EXENV_ERR_NULL ? ffi::ErrorCode::kInvalidArgument
                                                                    : ffi::ErrorCode::kInternal,
                    std::string(msg));
}

// vmap_step: args = D state leaves + action (+ D reference leaves when control_state is set);
//            rets = obs + D state leaves (+ NODE k leaves for Tsit5)
static ffi::Error StepImpl(cudaStream_t stream, ffi::RemainingArgs args, ffi::RemainingRets rets, int64_t plan,
                           int64_t n_phys, int64_t n_ode, int64_t n_control) {
  ExenvStepArgs a = {};
  auto first = args.get<ffi::AnyBuffer>(0);
  if (!first.has_value()) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "missing state leaf");
  a.batch = first->dimensions()[0];
  for (int64_t d = 0; d < n_phys; ++d) a.state_in[d] = args.get<ffi::AnyBuffer>(d)->untyped_data();
  a.action = args.get<ffi::AnyBuffer>(n_phys)->untyped_data();
  if (n_control > 0)
    for (int64_t d = 0; d < n_phys; ++d) a.ref[d] = args.get<ffi::AnyBuffer>(n_phys + 1 + d)->untyped_data();
  a.obs = (*rets.get<ffi::AnyBuffer>(0))->untyped_data();
  for (int64_t d = 0; d < n_phys; ++d) a.state_out[d] = (*rets.get<ffi::AnyBuffer>(1 + d))->untyped_data();
  for (int64_t d = 0; d < n_ode && 1 + n_phys + d < (int64_t)rets.size(); ++d)
    a.k_out[d] = (*rets.get<ffi::AnyBuffer>(1 + n_phys + d))->untyped_data();
  const int rc = exenv_step(reinterpret_cast<exenv_plan_t>(plan), &a, stream);
  return rc == EXENV_OK ? ffi::Error::Success() : fail(rc);
}

// vmap_sim_ahead: args = D state leaves + actions (B, T, A) (+ D reference leaves);
//                 rets = observations (B, N+1, O') + D trajectory leaves (B, N+1) + D last-state leaves (+ NODE k_last)
static ffi::Error SimAheadImpl(cudaStream_t stream, ffi::RemainingArgs args, ffi::RemainingRets rets, int64_t plan,
                               int64_t n_phys, int64_t n_ode, int64_t n_control, double obs_stepsize,
                               double action_stepsize, int64_t max_steps, int64_t layout) {
  // `layout` = ExenvLayout.  XLA always reports the LOGICAL dimensions (B, T, A); with EXENV_LAYOUT_TIME_MAJOR the
  // caller has declared the physical layouts on the custom call (jax_ffi.py: input_layouts / output_layouts), so the
  // buffers behind these pointers are (T, B, A), (N+1, B, O'), (N+1, B).
  ExenvSimArgs a = {};
  a.layout = (int32_t)layout;
  auto acts = args.get<ffi::AnyBuffer>(n_phys);
  if (!acts.has_value()) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "missing actions");
  a.batch = acts->dimensions()[0];
  a.n_actions = acts->dimensions()[1];
  a.obs_stepsize = obs_stepsize;
  a.action_stepsize = action_stepsize;
  a.max_steps = max_steps;
  for (int64_t d = 0; d < n_phys; ++d) a.state_in[d] = args.get<ffi::AnyBuffer>(d)->untyped_data();
  a.actions = acts->untyped_data();
  if (n_control > 0)
    for (int64_t d = 0; d < n_phys; ++d) a.ref[d] = args.get<ffi::AnyBuffer>(n_phys + 1 + d)->untyped_data();
  a.obs = (*rets.get<ffi::AnyBuffer>(0))->untyped_data();
  for (int64_t d = 0; d < n_phys; ++d) {
    a.traj[d] = (*rets.get<ffi::AnyBuffer>(1 + d))->untyped_data();
    a.last[d] = (*rets.get<ffi::AnyBuffer>(1 + n_phys + d))->untyped_data();
  }
  for (int64_t d = 0; d < n_ode && 1 + 2 * n_phys + d < (int64_t)rets.size(); ++d)
    a.k_last[d] = (*rets.get<ffi::AnyBuffer>(1 + 2 * n_phys + d))->untyped_data();
  const int rc = exenv_sim_ahead(reinterpret_cast<exenv_plan_t>(plan), &a, stream);
  return rc == EXENV_OK ? ffi::Error::Success() : fail(rc);
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(ExenvStepFfi, StepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .RemainingArgs()
                                  .RemainingRets()
                                  .Attr<int64_t>("plan")
                                  .Attr<int64_t>("n_phys")
                                  .Attr<int64_t>("n_ode")
                                  .Attr<int64_t>("n_control"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(ExenvSimAheadFfi, SimAheadImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .RemainingArgs()
                                  .RemainingRets()
                                  .Attr<int64_t>("plan")
                                  .Attr<int64_t>("n_phys")
                                  .Attr<int64_t>("n_ode")
                                  .Attr<int64_t>("n_control")
                                  .Attr<double>("obs_stepsize")
                                  .Attr<double>("action_stepsize")
                                  .Attr<int64_t>("max_steps")
                                  .Attr<int64_t>("layout"));
#endif  // EXENV_HAVE_XLA_FFI
