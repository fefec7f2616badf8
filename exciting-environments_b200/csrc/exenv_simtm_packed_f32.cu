// exenv_simtm_packed_f32.cu -- translation unit of the packed-arithmetic (FFMA2) time-major trajectory kernels:
// environments marked env_packable, fp32, two or four environments per thread (exenv_simtm_packed.inl).
#include "exenv_simtm_common.cuh"

// Built as two objects (Makefile: -DEXENV_PACKED_PART=1 / 2) so that the two halves compile in parallel.
#ifndef EXENV_PACKED_PART
#define EXENV_PACKED_PART 0  // 0: everything in one object
#endif

namespace exenv {

int launch_simtm_packed_f32_part2(const ExenvPlan &plan, const ExenvSimArgs &x, const SimSchedule &sch,
                                  cudaStream_t stream, int v);

#include "exenv_simtm_packed.inl"

template <class Env, int SOLVER, int V, bool EPI>
static int launch_simtm_packed(const ExenvPlan &plan, const ExenvSimArgs &x, const SimSchedule &sch, cudaStream_t stream) {
  const SimTmK<float> a = simtm_args<Env, float>(plan, x, sch);
  const KParams<float> kp = make_kparams<float>(plan);
  return launch_simtm_packed_v<Env, SOLVER, V, EPI>(a, kp, x.batch, stream);
}

template <class Env>
static int packed_dispatch(const ExenvPlan &plan, const ExenvSimArgs &x, const SimSchedule &sch, cudaStream_t st, int v) {
  const bool tsit = plan.desc.solver == EXENV_TSIT5, epi = x.reward != nullptr;
#define EXENV_PK(S, V, E) return launch_simtm_packed<Env, S, V, E>(plan, x, sch, st)
  if (tsit) {
    if (v == 4) { if (epi) EXENV_PK(EXENV_TSIT5, 4, true); EXENV_PK(EXENV_TSIT5, 4, false); }
    if (epi) EXENV_PK(EXENV_TSIT5, 2, true);
    EXENV_PK(EXENV_TSIT5, 2, false);
  }
  if (v == 4) { if (epi) EXENV_PK(EXENV_EULER, 4, true); EXENV_PK(EXENV_EULER, 4, false); }
  if (epi) EXENV_PK(EXENV_EULER, 2, true);
  EXENV_PK(EXENV_EULER, 2, false);
#undef EXENV_PK
}

// v = 2 or 4 environments per thread; the caller (exenv_simtm.inl: launch_simtm_fast) has checked alignment, that the
// batch splits into groups of v, shared parameters and control_state == [].
#if EXENV_PACKED_PART != 2
int launch_simtm_packed_f32_part1(const ExenvPlan &plan, const ExenvSimArgs &x, const SimSchedule &sch,
                                  cudaStream_t stream, int v) {
  switch (plan.desc.env) {
    case EXENV_PENDULUM: return packed_dispatch<Pendulum<float>>(plan, x, sch, stream, v);
    case EXENV_MASS_SPRING_DAMPER: return packed_dispatch<MassSpringDamper<float>>(plan, x, sch, stream, v);
    case EXENV_FLUID_TANK: return packed_dispatch<FluidTank<float>>(plan, x, sch, stream, v);
  }
  set_error("no packed-arithmetic trajectory kernel for env %d", plan.desc.env);
  return EXENV_ERR_UNSUPPORTED;
}
#endif
#if EXENV_PACKED_PART != 1
int launch_simtm_packed_f32_part2(const ExenvPlan &plan, const ExenvSimArgs &x, const SimSchedule &sch,
                                  cudaStream_t stream, int v) {
  switch (plan.desc.env) {
    case EXENV_CART_POLE: return packed_dispatch<CartPole<float>>(plan, x, sch, stream, v);
    case EXENV_ACROBOT: return packed_dispatch<Acrobot<float>>(plan, x, sch, stream, v);
  }
  set_error("no packed-arithmetic trajectory kernel for env %d", plan.desc.env);
  return EXENV_ERR_UNSUPPORTED;
}
#endif
#if EXENV_PACKED_PART != 2
int launch_simtm_packed_f32(const ExenvPlan &plan, const ExenvSimArgs &x, const SimSchedule &sch, cudaStream_t stream,
                            int v) {
  if (plan.desc.env == EXENV_CART_POLE || plan.desc.env == EXENV_ACROBOT)
    return launch_simtm_packed_f32_part2(plan, x, sch, stream, v);
  return launch_simtm_packed_f32_part1(plan, x, sch, stream, v);
}
#endif

}  // namespace exenv
