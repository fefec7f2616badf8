// exenv_kernel_common.cuh -- pieces shared by the step / reset / sim_ahead kernels: kernel-argument parameter
// block, per-environment parameter fetch, 128-bit vector access helpers, the one-environment step routine and the
// (env, solver, dtype, per-env) dispatch.
#pragma once

#include <cuda_runtime.h>

#include "exenv_envs.cuh"
#include "exenv_internal.h"

namespace exenv {

// Environment parameters as kernel arguments: shared scalars live in the constant bank (v[]); a non-NULL pp[i]
// makes slot i a per-environment (B,) array (reference core_env.py:253-277, in_axes 0 vs None).
template <typename T>
struct KParams {
  T v[EXENV_MAX_PARAMS];
  const T *pp[EXENV_MAX_PARAMS];
};

template <class Env>
constexpr int num_params() {
  return Env::NSTATIC + 2 * Env::D + 2 * Env::A;
}

template <typename T>
inline KParams<T> make_kparams(const ExenvPlan &plan) {
  KParams<T> k;
  for (int i = 0; i < EXENV_MAX_PARAMS; ++i) {
    k.v[i] = (T)plan.desc.p[i];
    k.pp[i] = (const T *)plan.desc.pp[i];
  }
  return k;
}

template <typename T>
inline LutGeom<T> make_lut_geom(const ExenvPlan &plan) {
  LutGeom<T> g{};
  g.bytes = (unsigned)plan.lut.bytes;
  g.off_gy = (unsigned)plan.lut.off_gy;
  g.off_cells = (unsigned)plan.lut.off_cells;
  g.off_wx = (unsigned)plan.lut.off_wx;
  g.off_wy = (unsigned)plan.lut.off_wy;
  g.nx = plan.lut.nx;
  g.ny = plan.lut.ny;
  g.regular = plan.lut.regular;
  g.inv_dx = (T)plan.lut.inv_dx;
  g.inv_dy = (T)plan.lut.inv_dy;
  g.gx0 = (T)plan.lut.gx0;
  g.gy0 = (T)plan.lut.gy0;
  g.dx = (T)plan.lut.dx;
  g.dy = (T)plan.lut.dy;
  return g;
}
template <typename T>
inline LutView<T> make_lut_view(const ExenvPlan &plan) {
  LutView<T> L{};
  if (plan.lut.blob) L = lut_view_at<T>(plan.lut.blob, make_lut_geom<T>(plan));
  return L;
}

#if defined(__CUDACC__)

template <class Env, typename T, bool PER_ENV>
__device__ __forceinline__ void load_env_params(const KParams<T> &kp, long b, T *p) {
  constexpr int NP = num_params<Env>();
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    if constexpr (PER_ENV) {
      p[i] = kp.pp[i] ? __ldg(kp.pp[i] + b) : kp.v[i];
    } else {
      p[i] = kp.v[i];
    }
  }
}

// N consecutive elements as one 16-byte (or narrower) access.
template <typename T, int N>
struct alignas(sizeof(T) * N) VecT {
  T v[N];
};
template <typename T, int N>
__device__ __forceinline__ void vload(const T *p, T *out) {
  if constexpr (N == 1) {
    out[0] = __ldg(p);
  } else {
    const VecT<T, N> t = *reinterpret_cast<const VecT<T, N> *>(p);
#pragma unroll
    for (int i = 0; i < N; ++i) out[i] = t.v[i];
  }
}
template <typename T, int N>
__device__ __forceinline__ void vstore(T *p, const T *in) {
  if constexpr (N == 1) {
    p[0] = in[0];
  } else {
    VecT<T, N> t;
#pragma unroll
    for (int i = 0; i < N; ++i) t.v[i] = in[i];
    *reinterpret_cast<VecT<T, N> *>(p) = t;
  }
}

#endif  // __CUDACC__

// One environment, one control step (reference CoreEnvironment.step core_env.py:393-425; PMSM.step
// pmsm_env.py:851-883): de-normalise (PMSM: clip + dead time) -> solver step over [0, tau] -> wrap/clip -> torque.
// s[D] physical state in/out, a[A] normalised action, f_last[NODE] = Tsit5 FSAL derivative (solver_state leaf).
// Tsit5 evaluates a fresh k1 on every call (the reference re-runs solver.init each step).
template <class Env, int SOLVER, typename T>
EX_HD void env_step_one(T *s, const T *a, const T *p, typename Env::Aux &aux, T tau, T pmsm_adv, int deadtime,
                        T *f_last) {
  constexpr int NODE = Env::NODE, A = Env::A;
  const T *an = p + Env::NSTATIC + 2 * Env::D;
  T u[A], y0[NODE], y1[NODE];
  if constexpr (Env::IS_PMSM) {
    T uc0, uc1;
    pmsm_clip<T>(a[0], a[1], s[2], s[6], p[5], pmsm_adv, an, &uc0, &uc1);
    if (deadtime > 0) {  // pmsm_env.py:866-875
      u[0] = s[0];
      u[1] = s[1];
      s[0] = uc0;
      s[1] = uc1;
    } else {
      u[0] = uc0;
      u[1] = uc1;
    }
    aux.omega_el = s[6];
    y0[0] = s[3];
    y0[1] = s[4];
    y0[2] = s[2];
  } else {
#pragma unroll
    for (int i = 0; i < A; ++i) u[i] = denorm_mm<T>(a[i], an[2 * i], an[2 * i + 1]);
#pragma unroll
    for (int i = 0; i < NODE; ++i) y0[i] = s[i];
  }
  // Tsit5 (instruction bound): the FAST right-hand sides (prepared reciprocals, Newton-refined divisions, exenv_math.cuh)
  // in both precisions.  Euler: fp64 keeps the plain IEEE operators (the reference's golden vectors replay on this
  // path); fp32 uses the FAST forms for cart-pole and acrobot, whose IEEE divisions and cosines made the fp32 Euler step
  // instruction bound (0.73 / 0.65 of HBM; profiles/README.md round 2) -- the same forms the trajectory kernels use,
  // inside the 1e-6 per-step tolerance.
  constexpr bool FD = SOLVER == EXENV_TSIT5 || (sizeof(T) == 4 && (Env::ID == 3 || Env::ID == 4));
  T f0[NODE];
  const typename Env::Pre pre = Env::precompute(p);
  Env::template rhs<FD>(y0, u, p, pre, aux, f0);
  if constexpr (SOLVER == EXENV_EULER) {  // diffrax.Euler: y1 = y0 + (t1 - t0) * f(t0, y0)
#pragma unroll
    for (int i = 0; i < NODE; ++i) {
      y1[i] = y0[i] + f0[i] * tau;
      f_last[i] = f0[i];
    }
  } else {
    T k[7][NODE];
    tsit5_step<T, NODE>(
        y0, f0, tau, [&](int, const T *ys, T *f) { Env::template rhs<FD>(ys, u, p, pre, aux, f); }, y1, f_last, k);
  }
  Env::post(y1);
  if constexpr (Env::IS_PMSM) {
    s[3] = y1[0];
    s[4] = y1[1];
    s[2] = y1[2];
    s[5] = Env::template torque<FD>(y1[0], y1[1], p, aux);
  } else {
#pragma unroll
    for (int i = 0; i < NODE; ++i) s[i] = y1[i];
  }
}

// (env, solver, dtype, per-env params) -> template instantiation.  F is a generic callable invoked as
// f(EnvTag<Env>{}, SolverTag<S>{}, TypeTag<T>{}, BoolTag<PER_ENV>{}).
template <class E>
struct EnvTag {
  using type = E;
};
template <int S>
struct IntTag {
  static constexpr int value = S;
};
template <bool B>
struct BoolTag {
  static constexpr bool value = B;
};

}  // namespace exenv
