// xla_ffi_shim.cc -- XLA typed-FFI handlers over the C ABI, so that the reference's host code can stay JAX and call the
// sm_100a kernels with jax.ffi.ffi_call (BASELINE.json north_star).  Compiled only where the XLA FFI headers exist
// (`python -c "import jax; print(jax.ffi.include_dir())"`).  This image ships no jax: the file cannot be exercised
// here; tests/test_abi.py compiles it against tests/ffi_stub/ (a stub of the part of the public xla::ffi API used
// below, written from its documentation) so that it at least stays consistent with include/exenv_b200.h.
// See INTEGRATION.md for the Python side of every handler.
//
// Conventions shared by all handlers
//   * buffers arrive in the reference's pytree-leaf order: the D physical-state leaves (B,) first;
//   * the plan handle created with exenv_plan_create() travels as the int64 attribute `plan`;
//   * `epilogue` != 0 appends three results -- reward, truncated (uint8 / PRED), terminated -- to a step / sim_ahead
//     call: vmap(generate_reward / truncated / terminated) of gym_wrapper.py:117-126 and
//     vmap_generate_rew_trunc_term_ahead of core_env.py:618-647 fused into the same kernel;
//   * exenv_sim_ahead inside a custom call must not allocate: the Python wrapper calls exenv_sim_ahead_prepare for
//     the call shape at trace time (jax_ffi.py), so the handler only launches.
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
static ffi::Error done(int rc) { return rc == EXENV_OK ? ffi::Error::Success() : fail(rc); }
static void *arg_ptr(ffi::RemainingArgs &args, int64_t i) { return args.get<ffi::AnyBuffer>(i)->untyped_data(); }
static void *ret_ptr(ffi::RemainingRets &rets, int64_t i) { return (*rets.get<ffi::AnyBuffer>(i))->untyped_data(); }

// vmap_step (core_env.py:533-569): args = D state leaves + action (+ D reference leaves when control_state is set);
//            rets = obs + D state leaves (+ NODE k leaves for Tsit5) (+ reward, truncated, terminated)
static ffi::Error StepImpl(cudaStream_t stream, ffi::RemainingArgs args, ffi::RemainingRets rets, int64_t plan,
                           int64_t n_phys, int64_t n_ode, int64_t n_control, int64_t epilogue) {
  ExenvStepArgs a = {};
  auto first = args.get<ffi::AnyBuffer>(0);
  if (!first.has_value()) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "missing state leaf");
  a.batch = first->dimensions()[0];
  for (int64_t d = 0; d < n_phys; ++d) a.state_in[d] = arg_ptr(args, d);
  a.action = arg_ptr(args, n_phys);
  if (n_control > 0)
    for (int64_t d = 0; d < n_phys; ++d) a.ref[d] = arg_ptr(args, n_phys + 1 + d);
  int64_t r = 0;
  a.obs = ret_ptr(rets, r++);
  for (int64_t d = 0; d < n_phys; ++d) a.state_out[d] = ret_ptr(rets, r++);
  for (int64_t d = 0; d < n_ode; ++d) a.k_out[d] = ret_ptr(rets, r++);
  if (epilogue) {
    a.reward = ret_ptr(rets, r++);
    a.truncated = static_cast<uint8_t *>(ret_ptr(rets, r++));
    a.terminated = static_cast<uint8_t *>(ret_ptr(rets, r++));
  }
  return done(exenv_step(reinterpret_cast<exenv_plan_t>(plan), &a, stream));
}

// vmap_sim_ahead (core_env.py:571-616): args = D state leaves + actions (B, T, A) (+ D reference leaves);
//   rets = observations (B, N+1, O') + D trajectory leaves (B, N+1) + D last-state leaves (+ NODE k_last)
//          (+ reward (B, N, 1), truncated (B, N+1, n_trunc), terminated (B, N, 1))
static ffi::Error SimAheadImpl(cudaStream_t stream, ffi::RemainingArgs args, ffi::RemainingRets rets, int64_t plan,
                               int64_t n_phys, int64_t n_ode, int64_t n_control, double obs_stepsize,
                               double action_stepsize, int64_t max_steps, int64_t layout, int64_t epilogue) {
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
  for (int64_t d = 0; d < n_phys; ++d) a.state_in[d] = arg_ptr(args, d);
  a.actions = acts->untyped_data();
  if (n_control > 0)
    for (int64_t d = 0; d < n_phys; ++d) a.ref[d] = arg_ptr(args, n_phys + 1 + d);
  int64_t r = 0;
  a.obs = ret_ptr(rets, r++);
  for (int64_t d = 0; d < n_phys; ++d) a.traj[d] = ret_ptr(rets, r++);
  for (int64_t d = 0; d < n_phys; ++d) a.last[d] = ret_ptr(rets, r++);
  for (int64_t d = 0; d < n_ode; ++d) a.k_last[d] = ret_ptr(rets, r++);
  if (epilogue) {
    a.reward = ret_ptr(rets, r++);
    a.truncated = static_cast<uint8_t *>(ret_ptr(rets, r++));
    a.terminated = static_cast<uint8_t *>(ret_ptr(rets, r++));
  }
  return done(exenv_sim_ahead(reinterpret_cast<exenv_plan_t>(plan), &a, stream));
}

// vmap_reset state packing (core_env.py:664-687) / vmap_generate_state_from_observation (:689-705):
//   mode EXENV_RESET_DEFAULT: no args;  FROM_NORM / FROM_OBS: args = values (B, D) / (B, O');
//   FROM_STATE: args = D state leaves (+ D reference leaves), aliased to the state results by the caller
//   rets = D state leaves + D reference leaves + obs (B, O')
static ffi::Error ResetPackImpl(cudaStream_t stream, ffi::RemainingArgs args, ffi::RemainingRets rets, int64_t plan,
                                int64_t n_phys, int64_t mode, int64_t batch) {
  ExenvResetArgs a = {};
  a.batch = batch;
  a.mode = (int32_t)mode;
  if (mode == EXENV_RESET_FROM_NORM || mode == EXENV_RESET_FROM_OBS) a.values = arg_ptr(args, 0);
  for (int64_t d = 0; d < n_phys; ++d) {
    a.state_out[d] = ret_ptr(rets, d);
    a.ref_out[d] = ret_ptr(rets, n_phys + d);
  }
  a.obs = ret_ptr(rets, 2 * n_phys);
  return done(exenv_reset_pack(reinterpret_cast<exenv_plan_t>(plan), &a, stream));
}

// vmap(generate_reward / truncated / terminated) on materialised states (gym_wrapper.py:117-126) and
// vmap_generate_rew_trunc_term_ahead on flattened trajectory leaves (core_env.py:618-647):
//   args = D state leaves (n,) (+ D reference leaves); rets = reward (n,), truncated (n, n_trunc), terminated (n,)
static ffi::Error EpilogueImpl(cudaStream_t stream, ffi::RemainingArgs args, ffi::RemainingRets rets, int64_t plan,
                               int64_t n_phys, int64_t n_control) {
  ExenvEpilogueArgs a = {};
  auto first = args.get<ffi::AnyBuffer>(0);
  if (!first.has_value()) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "missing state leaf");
  a.batch = first->element_count();
  for (int64_t d = 0; d < n_phys; ++d) a.state[d] = arg_ptr(args, d);
  if (n_control > 0)
    for (int64_t d = 0; d < n_phys; ++d) a.ref[d] = arg_ptr(args, n_phys + d);
  a.reward = ret_ptr(rets, 0);
  a.truncated = static_cast<uint8_t *>(ret_ptr(rets, 1));
  a.terminated = static_cast<uint8_t *>(ret_ptr(rets, 2));
  return done(exenv_epilogue(reinterpret_cast<exenv_plan_t>(plan), &a, stream));
}

// GymWrapper.gym_step after vmap_step (gym_wrapper.py:111-126, 170-192): reference generator + epilogue.
//   args = D state leaves, D reference leaves, hold_steps (B,) int32, rng (B, 2) uint64
//   rets = D reference leaves, hold_steps, rng (each aliased to its argument: input_output_aliases), reward, truncated,
//          terminated
static ffi::Error GymEpilogueImpl(cudaStream_t stream, ffi::RemainingArgs args, ffi::RemainingRets rets, int64_t plan,
                                  int64_t n_phys, int64_t hold_steps_min, int64_t hold_steps_max, int64_t force_new) {
  ExenvGymArgs a = {};
  auto first = args.get<ffi::AnyBuffer>(0);
  if (!first.has_value()) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "missing state leaf");
  a.batch = first->dimensions()[0];
  for (int64_t d = 0; d < n_phys; ++d) {
    a.state[d] = arg_ptr(args, d);
    a.ref[d] = ret_ptr(rets, d);  // == the reference argument (aliased)
  }
  a.hold_steps = static_cast<int32_t *>(ret_ptr(rets, n_phys));
  a.rng = static_cast<uint64_t *>(ret_ptr(rets, n_phys + 1));
  a.hold_steps_min = (int32_t)hold_steps_min;
  a.hold_steps_max = (int32_t)hold_steps_max;
  a.force_new = (int32_t)force_new;
  a.reward = ret_ptr(rets, n_phys + 2);
  a.truncated = static_cast<uint8_t *>(ret_ptr(rets, n_phys + 3));
  a.terminated = static_cast<uint8_t *>(ret_ptr(rets, n_phys + 4));
  return done(exenv_gym_epilogue(reinterpret_cast<exenv_plan_t>(plan), &a, stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(ExenvStepFfi, StepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .RemainingArgs()
                                  .RemainingRets()
                                  .Attr<int64_t>("plan")
                                  .Attr<int64_t>("n_phys")
                                  .Attr<int64_t>("n_ode")
                                  .Attr<int64_t>("n_control")
                                  .Attr<int64_t>("epilogue"));

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
                                  .Attr<int64_t>("layout")
                                  .Attr<int64_t>("epilogue"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(ExenvResetPackFfi, ResetPackImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .RemainingArgs()
                                  .RemainingRets()
                                  .Attr<int64_t>("plan")
                                  .Attr<int64_t>("n_phys")
                                  .Attr<int64_t>("mode")
                                  .Attr<int64_t>("batch"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(ExenvEpilogueFfi, EpilogueImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .RemainingArgs()
                                  .RemainingRets()
                                  .Attr<int64_t>("plan")
                                  .Attr<int64_t>("n_phys")
                                  .Attr<int64_t>("n_control"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(ExenvGymEpilogueFfi, GymEpilogueImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .RemainingArgs()
                                  .RemainingRets()
                                  .Attr<int64_t>("plan")
                                  .Attr<int64_t>("n_phys")
                                  .Attr<int64_t>("hold_steps_min")
                                  .Attr<int64_t>("hold_steps_max")
                                  .Attr<int64_t>("force_new"));
#endif  // EXENV_HAVE_XLA_FFI
