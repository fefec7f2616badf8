// exenv_step.inl (included by exenv_step_f32.cu / exenv_step_f64.cu with EXENV_REAL / EXENV_SFX) -- vmap_step and vmap_reset state packing as sm_100a kernels.
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
#include <cstring>

#include "exenv_kernel_common.cuh"

#define EXENV_CAT_(a, b) a##b
#define EXENV_CAT(a, b) EXENV_CAT_(a, b)

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

// Contiguous byte run with the widest naturally aligned stores (N is a compile-time count, dst is N-aligned when N is
// a power of two <= 16).
template <int N>
__device__ __forceinline__ void store_bytes(uint8_t *dst, const uint8_t *v) {
  if constexpr (N % 16 == 0) {
#pragma unroll
    for (int i = 0; i < N / 16; ++i) {
      uint4 w;
      memcpy(&w, v + 16 * i, 16);
      reinterpret_cast<uint4 *>(dst)[i] = w;
    }
  } else if constexpr (N % 8 == 0) {
#pragma unroll
    for (int i = 0; i < N / 8; ++i) {
      uint2 w;
      memcpy(&w, v + 8 * i, 8);
      reinterpret_cast<uint2 *>(dst)[i] = w;
    }
  } else if constexpr (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N / 4; ++i) {
      uint32_t w;
      memcpy(&w, v + 4 * i, 4);
      reinterpret_cast<uint32_t *>(dst)[i] = w;
    }
  } else if constexpr (N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      uint16_t w;
      memcpy(&w, v + 2 * i, 2);
      reinterpret_cast<uint16_t *>(dst)[i] = w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) dst[i] = v[i];
  }
}

// MODE 0: control_state == [] and no epilogue (plain vmap_step)            -> leanest register footprint
// MODE 1: control_state == [] with the fused reward/truncated/terminated   -> GymWrapper default
// MODE 2: generic (runtime n_control, runtime epilogue pointers, references; rows written with scalar stores)
template <class Env, int SOLVER, typename T>
constexpr int step_min_blocks() {
  if (Env::IS_PMSM) return sizeof(T) == 4 ? 3 : 2;  // 7 state + 8 obs scalars per env, two (one) envs per thread
  return SOLVER == EXENV_EULER ? (sizeof(T) == 4 ? 4 : 3) : (sizeof(T) == 4 ? 3 : 2);
}

template <class Env, int SOLVER, typename T, bool PER_ENV, int VEC, int MODE>
__global__ void __launch_bounds__(256, step_min_blocks<Env, SOLVER, T>())
    step_kernel(const StepK<T> a, const KParams<T> kp, const LutView<T> lut) {
  constexpr int D = Env::D, A = Env::A, NODE = Env::NODE, O = Env::O, NP = num_params<Env>();
  constexpr int NT0 = trunc_dim<Env>(0);  // truncated width without reference entries
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t * VEC >= a.count) return;
  const long b0 = a.first + t * VEC;

  // ---- coalesced 128-bit loads: D state leaves, A*VEC action scalars
  T s[VEC][D], act[VEC * A];
  {
    T tmp[VEC];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      vload<T, VEC>(a.sin[d] + b0, tmp);
#pragma unroll
      for (int v = 0; v < VEC; ++v) s[v][d] = tmp[v];
    }
#pragma unroll
    {  // VEC * A contiguous scalars with the widest (<= 16-byte) vectors
      constexpr int VW = (VEC * A) % (16 / (int)sizeof(T)) == 0 ? 16 / (int)sizeof(T) : VEC;
#pragma unroll
      for (int i = 0; i < VEC * A / VW; ++i) vload<T, VW>(a.action + b0 * A + i * VW, act + i * VW);
    }
  }

  T ob[VEC * O], fl[VEC][NODE];
  uint8_t trc[VEC * NT0], trm[VEC];
  T rew[VEC];

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
    observe<Env, T>(s[v], p, ob + v * O);
    if constexpr (MODE == 1) {
      // control_state == []: reward is the constant 0 (reference: the Python int 0)
      rew[v] = (T)0;
#pragma unroll
      for (int i = 0; i < NT0; ++i) trc[v * NT0 + i] = trunc_entry<Env, T>(i, s[v], ob + v * O, p);
      trm[v] = terminated_of<Env, T>((T)0, trc[v * NT0] != 0);
    }
    if constexpr (MODE == 2) {
      const int n_control = a.n_control;
      const int OW = O + n_control;
      const long b = b0 + v;
      T full[O + EXENV_MAX_CONTROL], ref[D];
#pragma unroll
      for (int o = 0; o < O; ++o) full[o] = ob[v * O + o];
      if (n_control > 0) {
#pragma unroll
        for (int d = 0; d < D; ++d) ref[d] = __ldg(a.ref[d] + b);
#pragma unroll
        for (int j = 0; j < EXENV_MAX_CONTROL; ++j)
          if (j < n_control) full[O + j] = norm_field<Env, T>(a.control[j], ref, p);
      }
#pragma unroll
      for (int o = 0; o < O + EXENV_MAX_CONTROL; ++o)
        if (o < OW) a.obs[b * OW + o] = full[o];
      if (a.reward || a.trunc || a.term) {
        const T rw = n_control > 0 ? reward_of<Env, T>(s[v], ref, p, n_control, a.control) : (T)0;
        const int nt = trunc_dim<Env>(n_control);
        bool t0 = false;
#pragma unroll
        for (int i = 0; i < O + EXENV_MAX_CONTROL; ++i)
          if (i < nt) {
            const bool tv = trunc_entry<Env, T>(i, s[v], full, p);
            if (i == 0) t0 = tv;
            if (a.trunc) a.trunc[b * nt + i] = tv;
          }
        if (a.reward) a.reward[b] = rw;
        if (a.term) a.term[b] = terminated_of<Env, T>(rw, t0);
      }
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
    if constexpr (MODE != 2) {  // VEC * O contiguous scalars
      {
        constexpr int VW = (VEC * O) % (16 / (int)sizeof(T)) == 0 ? 16 / (int)sizeof(T) : VEC;
#pragma unroll
        for (int i = 0; i < VEC * O / VW; ++i) vstore<T, VW>(a.obs + b0 * O + i * VW, ob + i * VW);
      }
    }
    if constexpr (MODE == 1) {
      if (a.reward) vstore<T, VEC>(a.reward + b0, rew);
      if (a.trunc) store_bytes<VEC * NT0>(a.trunc + b0 * NT0, trc);
      if (a.term) store_bytes<VEC>(a.term + b0, trm);
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

// reward / truncated / terminated of given states: one thread per state.
#if EXENV_STEP_SOLVER == 0
template <typename T>
struct EpiK {
  long B;
  const T *s[EXENV_MAX_PHYS];
  const T *ref[EXENV_MAX_PHYS];
  T *reward;
  uint8_t *trunc;
  uint8_t *term;
  int n_control;
  int control[EXENV_MAX_CONTROL];
};

template <class Env, typename T>
__global__ void __launch_bounds__(256) epilogue_kernel(const EpiK<T> a, const KParams<T> kp) {
  constexpr int D = Env::D, O = Env::O;
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  const T *p = kp.v;
  T s[D], r[D], ob[O + EXENV_MAX_CONTROL];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    s[d] = __ldg(a.s[d] + b);
    r[d] = a.n_control > 0 ? __ldg(a.ref[d] + b) : (T)0;
  }
  observe<Env, T>(s, p, ob);
#pragma unroll
  for (int j = 0; j < EXENV_MAX_CONTROL; ++j)
    if (j < a.n_control) ob[O + j] = norm_field<Env, T>(a.control[j], r, p);
  const int nt = trunc_dim<Env>(a.n_control);
  bool t0 = false;
#pragma unroll
  for (int i = 0; i < O + EXENV_MAX_CONTROL; ++i)
    if (i < nt) {
      const bool tv = trunc_entry<Env, T>(i, s, ob, p);
      if (i == 0) t0 = tv;
      if (a.trunc) a.trunc[b * nt + i] = tv;
    }
  const T rw = a.n_control > 0 ? reward_of<Env, T>(s, r, p, a.n_control, a.control) : (T)0;
  if (a.reward) a.reward[b] = rw;
  if (a.term) a.term[b] = terminated_of<Env, T>(rw, t0);
}

// GymWrapper.gym_step after vmap_step (reference gym_wrapper.py:111-126): update_ref -- where an environment's hold
// counter has run out, the tracked reference fields take the physical state of a freshly drawn random initial state and
// a new hold time is drawn (:170-192) -- then reward / terminated / truncated against the UPDATED reference.  One thread
// per environment; random numbers from Philox4x32-10 keyed per environment (rng[2 b] = key, rng[2 b + 1] = draw counter).
template <typename T>
struct GymK {
  long B;
  const T *s[EXENV_MAX_PHYS];
  T *ref[EXENV_MAX_PHYS];
  int32_t *hold;
  uint64_t *rng;
  int hold_min, hold_max, force_new;
  T *reward;
  uint8_t *trunc;
  uint8_t *term;
  int n_control;
  int control[EXENV_MAX_CONTROL];
};

template <class Env, typename T>
__global__ void __launch_bounds__(256) gym_epilogue_kernel(const GymK<T> a, const KParams<T> kp, const LutView<T> lut) {
  constexpr int D = Env::D, O = Env::O;
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  const T *p = kp.v;
  T s[D], r[D], ob[O + EXENV_MAX_CONTROL];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    s[d] = __ldg(a.s[d] + b);
    r[d] = a.n_control > 0 ? a.ref[d][b] : (T)0;
  }
  if (a.hold && a.n_control > 0) {
    int h = a.hold[b];
    if (a.force_new || h == 0) {
      const uint64_t key = a.rng[2 * b], ctr = a.rng[2 * b + 1];
      uint32_t w[8];
      philox4x32_10((uint32_t)key, (uint32_t)(key >> 32), (uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u, w);
      philox4x32_10((uint32_t)key, (uint32_t)(key >> 32), (uint32_t)ctr, (uint32_t)(ctr >> 32), 1u, 0u, w + 4);
      T u[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = u01_from_bits<T>(w[i]);
      typename Env::Aux aux{};
      if constexpr (Env::IS_PMSM) aux.lut = lut;
      T fresh[D];
      random_physical_state<Env, T>(u, p, aux, fresh);
      for (int j = 0; j < a.n_control; ++j) {
        const int f = a.control[j];
#pragma unroll
        for (int d = 0; d < D; ++d)
          if (d == f) {
            r[d] = fresh[d];
            a.ref[d][b] = fresh[d];
          }
      }
      const uint32_t span = (uint32_t)(a.hold_max > a.hold_min ? a.hold_max - a.hold_min : 1);
      h = a.hold_min + (int)(w[7] % span);  // jax.random.randint(minval, maxval): maxval exclusive
      a.rng[2 * b + 1] = ctr + 1;
    }
    a.hold[b] = a.force_new ? h : h - 1;
  }
  observe<Env, T>(s, p, ob);
#pragma unroll
  for (int j = 0; j < EXENV_MAX_CONTROL; ++j)
    if (j < a.n_control) ob[O + j] = norm_field<Env, T>(a.control[j], r, p);
  const int nt = trunc_dim<Env>(a.n_control);
  bool t0 = false;
#pragma unroll
  for (int i = 0; i < O + EXENV_MAX_CONTROL; ++i)
    if (i < nt) {
      const bool tv = trunc_entry<Env, T>(i, s, ob, p);
      if (i == 0) t0 = tv;
      if (a.trunc) a.trunc[b * nt + i] = tv;
    }
  const T rw = a.n_control > 0 ? reward_of<Env, T>(s, r, p, a.n_control, a.control) : (T)0;
  if (a.reward) a.reward[b] = rw;
  if (a.term) a.term[b] = terminated_of<Env, T>(rw, t0);
}

template <class Env, typename T>
static int launch_gym_epilogue_t(const ExenvPlan &plan, const ExenvGymArgs &x, cudaStream_t stream) {
  GymK<T> a{};
  a.B = x.batch;
  for (int d = 0; d < Env::D; ++d) {
    a.s[d] = (const T *)x.state[d];
    a.ref[d] = (T *)x.ref[d];
  }
  a.hold = x.hold_steps;
  a.rng = x.rng;
  a.hold_min = x.hold_steps_min;
  a.hold_max = x.hold_steps_max;
  a.force_new = x.force_new;
  a.reward = (T *)x.reward;
  a.trunc = x.truncated;
  a.term = x.terminated;
  a.n_control = plan.desc.n_control;
  for (int j = 0; j < EXENV_MAX_CONTROL; ++j) a.control[j] = plan.desc.control[j];
  const unsigned grid = (unsigned)((x.batch + 255) / 256);
  gym_epilogue_kernel<Env, T><<<grid, 256, 0, stream>>>(a, make_kparams<T>(plan), make_lut_view<T>(plan));
  g_launch_count++;
  const cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("exenv_gym_epilogue launch failed: %s", cudaGetErrorString(e));
    return EXENV_ERR_CUDA;
  }
  return EXENV_OK;
}

template <class Env, typename T>
static int launch_epilogue_t(const ExenvPlan &plan, const ExenvEpilogueArgs &x, cudaStream_t stream) {
  EpiK<T> a{};
  a.B = x.batch;
  for (int d = 0; d < Env::D; ++d) {
    a.s[d] = (const T *)x.state[d];
    a.ref[d] = (const T *)x.ref[d];
  }
  a.reward = (T *)x.reward;
  a.trunc = x.truncated;
  a.term = x.terminated;
  a.n_control = plan.desc.n_control;
  for (int j = 0; j < EXENV_MAX_CONTROL; ++j) a.control[j] = plan.desc.control[j];
  const KParams<T> kp = make_kparams<T>(plan);
  const unsigned grid = (unsigned)((x.batch + 255) / 256);
  epilogue_kernel<Env, T><<<grid, 256, 0, stream>>>(a, kp);
  g_launch_count++;
  const cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("exenv_epilogue launch failed: %s", cudaGetErrorString(e));
    return EXENV_ERR_CUDA;
  }
  return EXENV_OK;
}

int EXENV_CAT(launch_epilogue_, EXENV_SFX)(const ExenvPlan &plan, const ExenvEpilogueArgs &x, cudaStream_t st) {
  using T = EXENV_REAL;
  switch (plan.desc.env) {
    case EXENV_PENDULUM: return launch_epilogue_t<Pendulum<T>, T>(plan, x, st);
    case EXENV_MASS_SPRING_DAMPER: return launch_epilogue_t<MassSpringDamper<T>, T>(plan, x, st);
    case EXENV_FLUID_TANK: return launch_epilogue_t<FluidTank<T>, T>(plan, x, st);
    case EXENV_CART_POLE: return launch_epilogue_t<CartPole<T>, T>(plan, x, st);
    case EXENV_ACROBOT: return launch_epilogue_t<Acrobot<T>, T>(plan, x, st);
    case EXENV_PMSM: return launch_epilogue_t<Pmsm<T, false>, T>(plan, x, st);
  }
  set_error("unknown env %d", plan.desc.env);
  return EXENV_ERR_INVALID;
}

int EXENV_CAT(launch_gym_epilogue_, EXENV_SFX)(const ExenvPlan &plan, const ExenvGymArgs &x, cudaStream_t st) {
  using T = EXENV_REAL;
  switch (plan.desc.env) {
    case EXENV_PENDULUM: return launch_gym_epilogue_t<Pendulum<T>, T>(plan, x, st);
    case EXENV_MASS_SPRING_DAMPER: return launch_gym_epilogue_t<MassSpringDamper<T>, T>(plan, x, st);
    case EXENV_FLUID_TANK: return launch_gym_epilogue_t<FluidTank<T>, T>(plan, x, st);
    case EXENV_CART_POLE: return launch_gym_epilogue_t<CartPole<T>, T>(plan, x, st);
    case EXENV_ACROBOT: return launch_gym_epilogue_t<Acrobot<T>, T>(plan, x, st);
    case EXENV_PMSM:
      if (plan.desc.flags & EXENV_FLAG_SATURATED) return launch_gym_epilogue_t<Pmsm<T, true>, T>(plan, x, st);
      return launch_gym_epilogue_t<Pmsm<T, false>, T>(plan, x, st);
  }
  set_error("unknown env %d", plan.desc.env);
  return EXENV_ERR_INVALID;
}

#endif  // epilogue: Euler translation unit only

// ---- launchers ---------------------------------------------------------------------------------------------------
static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <class Env, int SOLVER, typename T, bool PER_ENV>
static int launch_step_t(const ExenvPlan &plan, const ExenvStepArgs &x, cudaStream_t stream) {
  // environments per thread: 16-byte vectors; 8-byte ones where the per-environment state is large (PMSM: 7 leaves + 8
  // observation entries; cart-pole / acrobot Tsit5 fp32 and Euler fp64), so that twice as many warps fit the register
  // file and cover the gap between the load and store phases (measured on B200, profiles/README.md)
  constexpr bool NARROW = Env::IS_PMSM || (Env::ID >= 3 && ((SOLVER == EXENV_TSIT5 && sizeof(T) == 4) ||
                                                            (SOLVER == EXENV_EULER && sizeof(T) == 8)));
  constexpr int VEC = (NARROW ? 8 : 16) / sizeof(T);
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
  vec_ok &= aligned16(a.action) && aligned16(a.obs) && aligned16(a.reward) && aligned16(a.trunc) && aligned16(a.term);
  a.n_control = plan.desc.n_control;
  for (int j = 0; j < EXENV_MAX_CONTROL; ++j) a.control[j] = plan.desc.control[j];
  a.deadtime = plan.deadtime;
  a.tau = (T)plan.desc.tau;
  a.pmsm_adv = (T)plan.pmsm_adv;
  const KParams<T> kp = make_kparams<T>(plan);
  const LutView<T> lut = make_lut_view<T>(plan);
  const long B = x.batch;
  const long n_vec = vec_ok ? (B / VEC) * VEC : 0;
  const bool epi = a.reward || a.trunc || a.term;
  if (n_vec > 0) {
    a.first = 0;
    a.count = n_vec;
    const long threads = n_vec / VEC;
    const unsigned grid = (unsigned)((threads + 255) / 256);
    if (PER_ENV || a.n_control > 0) {
      step_kernel<Env, SOLVER, T, PER_ENV, VEC, 2><<<grid, 256, 0, stream>>>(a, kp, lut);
    } else if constexpr (!PER_ENV) {
      if (epi)
        step_kernel<Env, SOLVER, T, false, VEC, 1><<<grid, 256, 0, stream>>>(a, kp, lut);
      else
        step_kernel<Env, SOLVER, T, false, VEC, 0><<<grid, 256, 0, stream>>>(a, kp, lut);
    }
    g_launch_count++;
  }
  if (n_vec < B) {  // ragged tail or unaligned views: scalar accesses
    a.first = n_vec;
    a.count = B - n_vec;
    const unsigned grid = (unsigned)((a.count + 255) / 256);
    step_kernel<Env, SOLVER, T, PER_ENV, 1, 2><<<grid, 256, 0, stream>>>(a, kp, lut);
    g_launch_count++;
  }
  const cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("exenv_step launch failed: %s", cudaGetErrorString(e));
    return EXENV_ERR_CUDA;
  }
  return EXENV_OK;
}

#if EXENV_STEP_SOLVER == 0
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

#endif

// One translation unit per solver (EXENV_STEP_SOLVER): the Euler instantiations are compiled with -fmad=false (HBM
// bound; golden-vector replay), the Tsit5 ones with FMA contraction (7 right-hand sides per step: instruction bound).
template <class Env, typename T>
static int step_dispatch2(const ExenvPlan &plan, const ExenvStepArgs &x, cudaStream_t st) {
  if constexpr (Env::IS_PMSM) {  // PMSM properties are static jit arguments in the reference: shared only
    return launch_step_t<Env, EXENV_STEP_SOLVER, T, false>(plan, x, st);
  } else {
    if (plan.per_env) return launch_step_t<Env, EXENV_STEP_SOLVER, T, true>(plan, x, st);
    return launch_step_t<Env, EXENV_STEP_SOLVER, T, false>(plan, x, st);
  }
}

// PMSM is ill-conditioned enough (cancellation in the voltage equations, 2x2 LU of the flux-map Jacobian) for FMA
// contraction to move a single Tsit5 step by more than the per-step parity tolerance, so its Tsit5 instantiations
// live in the un-contracted translation unit as well.
int EXENV_CAT(launch_step_tsit5_pmsm_, EXENV_SFX)(const ExenvPlan &plan, const ExenvStepArgs &x, cudaStream_t st);
#if EXENV_STEP_SOLVER == 0
int EXENV_CAT(launch_step_tsit5_pmsm_, EXENV_SFX)(const ExenvPlan &plan, const ExenvStepArgs &x, cudaStream_t st) {
  using T = EXENV_REAL;
  if (plan.desc.flags & EXENV_FLAG_SATURATED) return launch_step_t<Pmsm<T, true>, EXENV_TSIT5, T, false>(plan, x, st);
  return launch_step_t<Pmsm<T, false>, EXENV_TSIT5, T, false>(plan, x, st);
}
#endif

template <typename T>
static int step_dispatch(const ExenvPlan &plan, const ExenvStepArgs &x, cudaStream_t st) {
  switch (plan.desc.env) {
    case EXENV_PENDULUM: return step_dispatch2<Pendulum<T>, T>(plan, x, st);
    case EXENV_MASS_SPRING_DAMPER: return step_dispatch2<MassSpringDamper<T>, T>(plan, x, st);
    case EXENV_FLUID_TANK: return step_dispatch2<FluidTank<T>, T>(plan, x, st);
    case EXENV_CART_POLE: return step_dispatch2<CartPole<T>, T>(plan, x, st);
    case EXENV_ACROBOT: return step_dispatch2<Acrobot<T>, T>(plan, x, st);
    case EXENV_PMSM:
#if EXENV_STEP_SOLVER == 1
      return EXENV_CAT(launch_step_tsit5_pmsm_, EXENV_SFX)(plan, x, st);
#else
      if (plan.desc.flags & EXENV_FLAG_SATURATED) return step_dispatch2<Pmsm<T, true>, T>(plan, x, st);
      return step_dispatch2<Pmsm<T, false>, T>(plan, x, st);
#endif
  }
  set_error("unknown env %d", plan.desc.env);
  return EXENV_ERR_INVALID;
}

#if EXENV_STEP_SOLVER == 1
int EXENV_CAT(launch_step_tsit5_, EXENV_SFX)(const ExenvPlan &plan, const ExenvStepArgs &a, cudaStream_t stream) {
  return step_dispatch<EXENV_REAL>(plan, a, stream);
}
#else
int EXENV_CAT(launch_step_euler_, EXENV_SFX)(const ExenvPlan &plan, const ExenvStepArgs &a, cudaStream_t stream) {
  return step_dispatch<EXENV_REAL>(plan, a, stream);
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

int EXENV_CAT(launch_reset_, EXENV_SFX)(const ExenvPlan &plan, const ExenvResetArgs &a, cudaStream_t stream) {
  return reset_dispatch<EXENV_REAL>(plan, a, stream);
}

#endif  // Euler translation unit: reset + epilogue

}  // namespace exenv
