// exenv_step.cu -- vmap_step and vmap_reset state packing as sm_100a kernels.
//
// step_kernel: one thread owns VEC consecutive environments (VEC = 16 bytes / sizeof(T)), so every state leaf
// (B,), the action rows (B, A) and the observation rows (B, O) are moved with 128-bit coalesced accesses; the
// state lives in registers; shared parameters sit in the constant bank (kernel arguments), per-environment
// parameters are extra coalesced loads.  Fused in one pass: action de-normalisation (PMSM: dq->alpha/beta,
// voltage-hexagon clip, dead-time buffer), the solver step (Euler or 7-stage Tsit5), wrap/clip, torque,
// observation and the optional reward / truncated / terminated epilogue (reference gym_wrapper.py:117-126).
// Algorithmic traffic per env-step: S * (2D + A + O) bytes (+ S * NODE for the Tsit5 k output) -> HBM bound.
//
// Compiled with -fmad=false: the Euler path then reproduces the reference's golden vectors like un-contracted IEEE
// arithmetic does (the kernel is bandwidth bound, contraction would buy nothing).
#include <cstdio>

#include "exenv_kernel_common.cuh"

namespace exenv {

template <typename T>
struct StepK {
  long first, count;  // environments [first, first + count)
  const T *sin[EXENV_MAX_PHYS];
  T *sout[EXENV_MAX_PHYS];
  const T *action;
  T *obs;
  T *kout[EXENV_MAX_ODE];
  const T *ref[EXENV_MAX_PHYS];
  T *reward;
  uint8_t *trunc;
  uint8_t *term;
  int n_control;
  int control[EXENV_MAX_CONTROL];
  int deadtime;
  T tau, pmsm_adv;
};

template <class Env, int SOLVER, typename T, bool PER_ENV, int VEC>
__global__ void __launch_bounds__(256) step_kernel(const StepK<T> a, const KParams<T> kp, const LutView<T> lut) {
  constexpr int D = Env::D, A = Env::A, NODE = Env::NODE, O = Env::O, NP = num_params<Env>();
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t * VEC >= a.count) return;
  const long b0 = a.first + t * VEC;

  // ---- coalesced 128-bit loads: D state leaves, A*VEC action scalars (+ references when tracked)
  T s[VEC][D], act[VEC * A], ref[VEC][D];
  {
    T tmp[VEC];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      vload<T, VEC>(a.sin[d] + b0, tmp);
#pragma unroll
      for (int v = 0; v < VEC; ++v) s[v][d] = tmp[v];
    }
#pragma unroll
    for (int i = 0; i < A; ++i) vload<T, VEC>(a.action + b0 * A + i * VEC, act + i * VEC);
    if (a.n_control > 0) {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        vload<T, VEC>(a.ref[d] + b0, tmp);
#pragma unroll
        for (int v = 0; v < VEC; ++v) ref[v][d] = tmp[v];
      }
    }
  }

  const int n_control = a.n_control;
  const int OW = O + n_control;  // observation row width
  T ob[VEC][O + EXENV_MAX_CONTROL], fl[VEC][NODE], rew[VEC];
  uint8_t trc[VEC][O + EXENV_MAX_CONTROL], trm[VEC];
  const bool epilogue = a.reward || a.trunc || a.term;

#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    T pl[NP];
    const T *p = kp.v;
    if constexpr (PER_ENV) {
      load_env_params<Env, T, true>(kp, b0 + v, pl);
      p = pl;
    }
    typename Env::Aux aux{};
    if constexpr (Env::IS_PMSM) aux.lut = lut;
    env_step_one<Env, SOLVER, T>(s[v], act + v * A, p, aux, a.tau, a.pmsm_adv, a.deadtime, fl[v]);
    observe<Env, T>(s[v], p, ob[v]);
#pragma unroll
    for (int j = 0; j < EXENV_MAX_CONTROL; ++j)
      if (j < n_control) ob[v][O + j] = norm_field<Env, T>(a.control[j], ref[v], p);
    if (epilogue) {
      rew[v] = n_control > 0 ? reward_of<Env, T>(s[v], ref[v], p, n_control, a.control) : (T)0;
      const int nt = trunc_dim<Env>(n_control);
#pragma unroll
      for (int i = 0; i < O + EXENV_MAX_CONTROL; ++i)
        if (i < nt) trc[v][i] = trunc_entry<Env, T>(i, s[v], ob[v], p);
      trm[v] = terminated_of<Env, T>(rew[v], trc[v][0] != 0);
    }
  }

  // ---- coalesced stores
  {
    T tmp[VEC];
#pragma unroll
    for (int d = 0; d < D; ++d) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) tmp[v] = s[v][d];
      vstore<T, VEC>(a.sout[d] + b0, tmp);
    }
    if (SOLVER == EXENV_TSIT5 && a.kout[0]) {
#pragma unroll
      for (int d = 0; d < NODE; ++d) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) tmp[v] = fl[v][d];
        vstore<T, VEC>(a.kout[d] + b0, tmp);
      }
    }
    if (n_control == 0) {  // VEC * O contiguous scalars = O vectors
      T flat[VEC * O];
#pragma unroll
      for (int v = 0; v < VEC; ++v)
#pragma unroll
        for (int o = 0; o < O; ++o) flat[v * O + o] = ob[v][o];
#pragma unroll
      for (int i = 0; i < O; ++i) vstore<T, VEC>(a.obs + b0 * O + i * VEC, flat + i * VEC);
    } else {
#pragma unroll
      for (int v = 0; v < VEC; ++v)
#pragma unroll
        for (int o = 0; o < O + EXENV_MAX_CONTROL; ++o)
          if (o < OW) a.obs[(b0 + v) * OW + o] = ob[v][o];
    }
    if (epilogue) {
      if (a.reward) vstore<T, VEC>(a.reward + b0, rew);
      const int nt = trunc_dim<Env>(n_control);
      if (a.trunc) {
#pragma unroll
        for (int v = 0; v < VEC; ++v)
#pragma unroll
          for (int i = 0; i < O + EXENV_MAX_CONTROL; ++i)
            if (i < nt) a.trunc[(b0 + v) * nt + i] = trc[v][i];
      }
      if (a.term) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) a.term[b0 + v] = trm[v];
      }
    }
  }
}

// vmap_reset packing / generate_state_from_observation: one thread per environment.
// (reference core_env.py:664-687, :361-391; init_state e.g. pendulum_env.py:261-295, pmsm_env.py:383-485;
//  generate_state_from_observation e.g. pendulum_env.py:331-364, pmsm_env.py:921-970)
template <typename T>
struct ResetK {
  long B;
  int mode;
  const T *values;
  T *sout[EXENV_MAX_PHYS];
  T *rout[EXENV_MAX_PHYS];
  T *obs;
  int n_control;
  int control[EXENV_MAX_CONTROL];
};

template <class Env, typename T, bool PER_ENV>
__global__ void __launch_bounds__(256) reset_kernel(const ResetK<T> a, const KParams<T> kp) {
  constexpr int D = Env::D, O = Env::O, NP = num_params<Env>();
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  T pl[NP];
  const T *p = kp.v;
  if constexpr (PER_ENV) {
    load_env_params<Env, T, true>(kp, b, pl);
    p = pl;
  }
  const T *pn = p + Env::NSTATIC;
  const int OW = O + a.n_control;
  T z[D], s[D], r[D];
#pragma unroll
  for (int i = 0; i < D; ++i) r[i] = nan_of<T>();
  bool physical = false;
  if (a.mode == EXENV_RESET_DEFAULT) {
    if constexpr (Env::IS_PMSM) {  // pmsm_env.py:386-399: already physical values
#pragma unroll
      for (int i = 0; i < D; ++i) z[i] = (T)0;
      z[3] = (pn[6] + pn[7]) / (T)2;
      z[6] = (pn[12] + pn[13]) / (T)2;
      physical = true;
    } else {
      Env::default_norm_state(z);
    }
  } else if (a.mode == EXENV_RESET_FROM_NORM) {
#pragma unroll
    for (int i = 0; i < D; ++i) z[i] = a.values[b * D + i];
    physical = Env::IS_PMSM;
  } else if (a.mode == EXENV_RESET_FROM_OBS) {
    const T *ob = a.values + b * OW;
    if constexpr (Env::IS_PMSM) {  // pmsm_env.py:928-936
      z[0] = ob[6];
      z[1] = ob[7];
      z[2] = Mth<T>::atan2(ob[5], ob[4]) / Cst<T>::pi();
      z[3] = ob[0];
      z[4] = ob[1];
      z[5] = ob[3];
      z[6] = ob[2];
    } else {
#pragma unroll
      for (int i = 0; i < D; ++i) z[i] = ob[i];
    }
    for (int j = 0; j < a.n_control; ++j) {
      const int f = a.control[j];
      const T rv = denorm_mm<T>(ob[O + j], pn[2 * f], pn[2 * f + 1]);
#pragma unroll
      for (int i = 0; i < D; ++i)
        if (i == f) r[i] = rv;
    }
  } else {  // EXENV_RESET_FROM_STATE: initial_state passed through, only the observation is generated
#pragma unroll
    for (int i = 0; i < D; ++i) z[i] = a.sout[i][b];
    if (a.n_control > 0 && a.rout[0]) {
#pragma unroll
      for (int i = 0; i < D; ++i) r[i] = a.rout[i][b];
    }
    physical = true;
  }
#pragma unroll
  for (int i = 0; i < D; ++i) s[i] = physical ? z[i] : denorm_mm<T>(z[i], pn[2 * i], pn[2 * i + 1]);
  T ob[O + EXENV_MAX_CONTROL];
  observe<Env, T>(s, p, ob);
#pragma unroll
  for (int j = 0; j < EXENV_MAX_CONTROL; ++j)
    if (j < a.n_control) ob[O + j] = norm_field<Env, T>(a.control[j], r, p);
  if (a.mode != EXENV_RESET_FROM_STATE) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      a.sout[i][b] = s[i];
      if (a.rout[i]) a.rout[i][b] = r[i];
    }
  }
#pragma unroll
  for (int o = 0; o < O + EXENV_MAX_CONTROL; ++o)
    if (o < OW) a.obs[b * OW + o] = ob[o];
}

// ---- launchers ---------------------------------------------------------------------------------------------------
static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <class Env, int SOLVER, typename T, bool PER_ENV>
static int launch_step_t(const ExenvPlan &plan, const ExenvStepArgs &x, cudaStream_t stream) {
  constexpr int VEC = 16 / sizeof(T);
  StepK<T> a{};
  bool vec_ok = true;
  for (int d = 0; d < Env::D; ++d) {
    a.sin[d] = (const T *)x.state_in[d];
    a.sout[d] = (T *)x.state_out[d];
    a.ref[d] = (const T *)x.ref[d];
    vec_ok &= aligned16(a.sin[d]) && aligned16(a.sout[d]) && aligned16(a.ref[d]);
  }
  for (int d = 0; d < Env::NODE; ++d) {
    a.kout[d] = (T *)x.k_out[d];
    vec_ok &= aligned16(a.kout[d]);
  }
  a.action = (const T *)x.action;
  a.obs = (T *)x.obs;
  a.reward = (T *)x.reward;
  a.trunc = x.truncated;
  a.term = x.terminated;
  vec_ok &= aligned16(a.action) && aligned16(a.obs) && aligned16(a.reward);
  a.n_control = plan.desc.n_control;
  for (int j = 0; j < EXENV_MAX_CONTROL; ++j) a.control[j] = plan.desc.control[j];
  a.deadtime = plan.deadtime;
  a.tau = (T)plan.desc.tau;
  a.pmsm_adv = (T)plan.pmsm_adv;
  const KParams<T> kp = make_kparams<T>(plan);
  const LutView<T> lut = make_lut_view<T>(plan);
  const long B = x.batch;
  const long n_vec = vec_ok ? (B / VEC) * VEC : 0;
  if (n_vec > 0) {
    a.first = 0;
    a.count = n_vec;
    const long threads = n_vec / VEC;
    const unsigned grid = (unsigned)((threads + 255) / 256);
    step_kernel<Env, SOLVER, T, PER_ENV, VEC><<<grid, 256, 0, stream>>>(a, kp, lut);
    g_launch_count++;
  }
  if (n_vec < B) {  // ragged tail or unaligned views: scalar accesses
    a.first = n_vec;
    a.count = B - n_vec;
    const unsigned grid = (unsigned)((a.count + 255) / 256);
    step_kernel<Env, SOLVER, T, PER_ENV, 1><<<grid, 256, 0, stream>>>(a, kp, lut);
    g_launch_count++;
  }
  const cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("exenv_step launch failed: %s", cudaGetErrorString(e));
    return EXENV_ERR_CUDA;
  }
  return EXENV_OK;
}

template <class Env, typename T, bool PER_ENV>
static int launch_reset_t(const ExenvPlan &plan, const ExenvResetArgs &x, cudaStream_t stream) {
  ResetK<T> a{};
  a.B = x.batch;
  a.mode = x.mode;
  a.values = (const T *)x.values;
  for (int d = 0; d < Env::D; ++d) {
    a.sout[d] = (T *)x.state_out[d];
    a.rout[d] = (T *)x.ref_out[d];
  }
  a.obs = (T *)x.obs;
  a.n_control = plan.desc.n_control;
  for (int j = 0; j < EXENV_MAX_CONTROL; ++j) a.control[j] = plan.desc.control[j];
  const KParams<T> kp = make_kparams<T>(plan);
  const unsigned grid = (unsigned)((x.batch + 255) / 256);
  reset_kernel<Env, T, PER_ENV><<<grid, 256, 0, stream>>>(a, kp);
  g_launch_count++;
  const cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("exenv_reset_pack launch failed: %s", cudaGetErrorString(e));
    return EXENV_ERR_CUDA;
  }
  return EXENV_OK;
}

template <class Env, typename T>
static int step_dispatch2(const ExenvPlan &plan, const ExenvStepArgs &x, cudaStream_t st) {
  const bool tsit = plan.desc.solver == EXENV_TSIT5;
  if constexpr (Env::IS_PMSM) {  // PMSM properties are static jit arguments in the reference: shared only
    return tsit ? launch_step_t<Env, EXENV_TSIT5, T, false>(plan, x, st)
                : launch_step_t<Env, EXENV_EULER, T, false>(plan, x, st);
  } else {
    if (plan.per_env)
      return tsit ? launch_step_t<Env, EXENV_TSIT5, T, true>(plan, x, st)
                  : launch_step_t<Env, EXENV_EULER, T, true>(plan, x, st);
    return tsit ? launch_step_t<Env, EXENV_TSIT5, T, false>(plan, x, st)
                : launch_step_t<Env, EXENV_EULER, T, false>(plan, x, st);
  }
}

template <typename T>
static int step_dispatch(const ExenvPlan &plan, const ExenvStepArgs &x, cudaStream_t st) {
  switch (plan.desc.env) {
    case EXENV_PENDULUM: return step_dispatch2<Pendulum<T>, T>(plan, x, st);
    case EXENV_MASS_SPRING_DAMPER: return step_dispatch2<MassSpringDamper<T>, T>(plan, x, st);
    case EXENV_FLUID_TANK: return step_dispatch2<FluidTank<T>, T>(plan, x, st);
    case EXENV_CART_POLE: return step_dispatch2<CartPole<T>, T>(plan, x, st);
    case EXENV_ACROBOT: return step_dispatch2<Acrobot<T>, T>(plan, x, st);
    case EXENV_PMSM:
      if (plan.desc.flags & EXENV_FLAG_SATURATED) return step_dispatch2<Pmsm<T, true>, T>(plan, x, st);
      return step_dispatch2<Pmsm<T, false>, T>(plan, x, st);
  }
  set_error("unknown env %d", plan.desc.env);
  return EXENV_ERR_INVALID;
}

int launch_step(const ExenvPlan &plan, const ExenvStepArgs &a, cudaStream_t stream) {
  return plan.desc.dtype == EXENV_F64 ? step_dispatch<double>(plan, a, stream) : step_dispatch<float>(plan, a, stream);
}

template <class Env, typename T>
static int reset_dispatch2(const ExenvPlan &plan, const ExenvResetArgs &x, cudaStream_t st) {
  if constexpr (!Env::IS_PMSM) {
    if (plan.per_env) return launch_reset_t<Env, T, true>(plan, x, st);
  }
  return launch_reset_t<Env, T, false>(plan, x, st);
}

template <typename T>
static int reset_dispatch(const ExenvPlan &plan, const ExenvResetArgs &x, cudaStream_t st) {
  switch (plan.desc.env) {
    case EXENV_PENDULUM: return reset_dispatch2<Pendulum<T>, T>(plan, x, st);
    case EXENV_MASS_SPRING_DAMPER: return reset_dispatch2<MassSpringDamper<T>, T>(plan, x, st);
    case EXENV_FLUID_TANK: return reset_dispatch2<FluidTank<T>, T>(plan, x, st);
    case EXENV_CART_POLE: return reset_dispatch2<CartPole<T>, T>(plan, x, st);
    case EXENV_ACROBOT: return reset_dispatch2<Acrobot<T>, T>(plan, x, st);
    case EXENV_PMSM: return reset_dispatch2<Pmsm<T, false>, T>(plan, x, st);
  }
  set_error("unknown env %d", plan.desc.env);
  return EXENV_ERR_INVALID;
}

int launch_reset(const ExenvPlan &plan, const ExenvResetArgs &a, cudaStream_t stream) {
  return plan.desc.dtype == EXENV_F64 ? reset_dispatch<double>(plan, a, stream)
                                      : reset_dispatch<float>(plan, a, stream);
}

}  // namespace exenv
