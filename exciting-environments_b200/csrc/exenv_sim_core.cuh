// exenv_sim_core.cuh -- the per-environment body of the persistent sim_ahead kernel, free of any memory-layout
// knowledge: action preparation (de-normalisation; PMSM voltage-hexagon clip + dead-time shift), one solver step
// driven by a host-built schedule word, and the dense-output saves.  exenv_sim.cu wires it to shared-memory tiles;
// tests/hostcheck compiles the very same functions for the CPU to validate them against the oracle without a GPU.
#pragma once

#include <type_traits>

#include "../../include/exenv_b200.h"
#include "exenv_envs.cuh"

namespace exenv {

// Schedule word (built by exenv_api.cu: build_schedule): bits 0-31 smallest ZOH action index of the step,
// bits 32-45 seven 2-bit per-stage offsets from it, bits 46-49 number of saves falling into the step,
// bits 50-51 the largest of the seven offsets.
EX_HD int sched_base(uint64_t w) { return (int)(unsigned)(w & 0xffffffffull); }
EX_HD int sched_delta(uint64_t w, int stage) { return (int)((w >> (32 + 2 * stage)) & 3ull); }
EX_HD int sched_saves(uint64_t w) { return (int)((w >> 46) & 15ull); }
EX_HD int sched_dmax(uint64_t w) { return (int)((w >> 50) & 3ull); }  // largest stage offset of the step

// Number of values staged per save: the post-processed ODE state; PMSM: eps, i_d, i_q, torque, u_d_buffer,
// u_q_buffer, cos(eps), sin(eps) (the two observation entries that are not plain normalisations are evaluated once
// by the owner thread instead of divergently by the row writer).
template <class Env>
constexpr int staged_values() {
  return Env::IS_PMSM ? 8 : Env::NODE;
}

// Per-save dense-output record built by the host (make_saverec below; exenv_api.cu, tests/hostcheck): Tsit5 -> the
// seven weights b_i(theta) of diffrax's _Tsit5Interpolation, evaluated in T exactly like tsit5_dense_weights and
// multiplied by the length dt of the step the save falls into (the kernels keep derivatives, not increments:
// y(theta) = y0 + sum_i (b_i(theta) dt) f_i); Euler -> v[0] = theta.
template <typename T>
struct alignas(16) SaveRec {
  T v[8];
};
template <typename T>
inline void make_saverec(bool tsit5, double theta, double dt_step, SaveRec<T> &r) {
  for (int j = 0; j < 8; ++j) r.v[j] = (T)0;
  if (tsit5) {
    tsit5_dense_weights<T>((T)theta, r.v);
    for (int j = 0; j < 7; ++j) r.v[j] = r.v[j] * (T)dt_step;
  } else {
    r.v[0] = (T)theta;
  }
}
// Records of a whole schedule: save s belongs to the step that counts it in bits 46-49 of its word.
template <typename T, class Vec>
inline void make_saverecs(bool tsit5, long n_steps, const uint64_t *words, const double *dts, const double *thetas,
                          Vec &recs) {
  long s = 0;
  for (long n = 0; n < n_steps; ++n) {
    const int ns = (int)((words[n] >> 46) & 15ull);
    for (int q = 0; q < ns; ++q, ++s) make_saverec<T>(tsit5, thetas[s], dts[n], recs[(size_t)s]);
  }
}

// Per-step schedule entry of the time-major kernels: the schedule word, the step length and the dense-output record of
// the step's FIRST save (zeros when the step saves nothing; further saves of the same step -- rare -- are read from the
// per-save SaveRec array).  48 bytes in fp32, 96 in fp64; staged through shared memory 32 steps at a time.
template <typename T>
struct alignas(16) TmStep {
  uint64_t w;
  T dt;
  T pad;
  T wts[8];
};
constexpr int TM_CHUNK = 32;  // steps per staged chunk (one record per lane)

// Action of stage `s` of the step described by word `w`.  A fetch object may provide stage(w, s, u) and decode the
// 2-bit stage offset itself (time-major kernel: its window starts at the step's base index); plain callables
// fetch(idx, u) receive the absolute zero-order-hold index.
template <class F, typename T>
EX_HD auto fetch_stage_(F &f, uint64_t w, int s, T *u, int) -> decltype(f.stage(w, s, u), void()) {
  f.stage(w, s, u);
}
template <class F, typename T>
EX_HD void fetch_stage_(F &f, uint64_t w, int s, T *u, long) {
  f(sched_base(w) + sched_delta(w, s), u);
}

// Physical action for index `q` of the zero-order-hold sequence the solver sees.
//  generic envs : denormalize_action of raw action q                               (core_env.py:468)
//  PMSM         : q < deadtime -> the initial u_dq buffer; else raw action j = q - deadtime clipped to the voltage
//                 hexagon at eps_j = eps_0 + linspace(0, tau*(L-1), L)[j] * omega_el   (pmsm_env.py:709-777)
//  pre-scaled  : environments marked env_sim_prescaled keep Env::act_sim(raw, pre) instead (the de-normalised action
//                 already divided by the parameter the right-hand side divides it by), consumed by Env::rhs_sim
// `raw` = normalised raw action (A values; ignored for buffer entries), s0 = initial physical state.
template <class Env, typename T>
EX_HD void prepare_action(long q, const T *raw, const T *s0, const T *p, const typename Env::Pre &pre, int deadtime,
                          long n_actions, T pmsm_adv, T pmsm_lin_stop, T *u) {
  const T *an = p + Env::NSTATIC + 2 * Env::D;
  if constexpr (env_sim_prescaled<Env>::value) {
#pragma unroll
    for (int c = 0; c < Env::A; ++c) u[c] = Env::act_sim(raw[c], pre);
  } else if constexpr (Env::IS_PMSM) {
    if (q < deadtime) {
      u[0] = s0[0];
      u[1] = s0[1];
    } else {
      const long j = q - deadtime;
      T lin;
      if (n_actions == 1) lin = (T)0;
      else if (j == n_actions - 1) lin = pmsm_lin_stop;
      else lin = pmsm_lin_stop * ((T)j / (T)(n_actions - 1));
      const T eps_j = s0[2] + lin * s0[6];
      pmsm_clip<T>(raw[0], raw[1], eps_j, s0[6], p[5], pmsm_adv, an, &u[0], &u[1]);
    }
  } else {
#pragma unroll
    for (int c = 0; c < Env::A; ++c) u[c] = denorm_mm<T>(raw[c], an[2 * c], an[2 * c + 1]);
  }
}

// Right-hand side as the trajectory kernels evaluate it: `u` is what prepare_action stored.
template <class Env, typename T>
EX_HD void sim_rhs(const T *y, const T *u, const T *p, const typename Env::Pre &pre, typename Env::Aux &aux, T *dy) {
  if constexpr (env_sim_prescaled<Env>::value) Env::rhs_sim(y, u, pre, dy);
  else Env::template rhs<true>(y, u, p, pre, aux, dy);
}

// One solver step + its saves.
//   y[NODE]      ODE state, updated in place
//   f[NODE]      Tsit5: FSAL derivative at (t0, y); valid on entry (the caller evaluates it once before the first
//                step, as diffrax does on `first_step`), replaced by the stage-7 derivative on exit
//   fetch(idx,u) physical action of ZOH index idx (absolute)
//   sink(save_index, staged[NST]) receives every saved sample, already wrapped / clipped (+ torque, PMSM buffers)
//   rec0 / recs  per-save dense-output records: rec0 = recs[save_idx] (prefetched by the caller one step ahead so its
//                latency is off the critical path), further saves of the same step read recs[] directly
// Divisions by parameter-only denominators use the prepared reciprocals (div_inv<true>), see exenv_math.cuh.
template <class Env, int SOLVER, typename T, class Fetch, typename RT>
EX_HD void sim_save_one(const SaveRec<RT> &rec, const T *y, const T *y1, const T (*k)[Env::NODE], const T *p,
                        typename Env::Aux &aux, long save_idx, Fetch &&fetch, T *st) {
  constexpr int NODE = Env::NODE, A = Env::A;
  T ysv[NODE];
  if constexpr (SOLVER == EXENV_EULER) {  // LocalLinearInterpolation
#pragma unroll
    for (int i = 0; i < NODE; ++i) ysv[i] = y[i] + rec.v[0] * (y1[i] - y[i]);
  } else {  // _Tsit5Interpolation on the stage derivatives k = f (weights carry the step length, see SaveRec)
#pragma unroll
    for (int i = 0; i < NODE; ++i) {
      T acc = fma_(T(rec.v[0]), k[0][i], y[i]);
#pragma unroll
      for (int j = 1; j < 7; ++j) acc = fma_(T(rec.v[j]), k[j][i], acc);
      ysv[i] = acc;
    }
  }
  Env::post(ysv);
  if constexpr (Env::IS_PMSM) {
    st[0] = ysv[2];
    st[1] = ysv[0];
    st[2] = ysv[1];
    st[3] = Env::template torque<true>(ysv[0], ysv[1], p, aux);
    // u_dq buffer leaves of saved state j: vstack([buffer]*deadtime, clipped actions)[j]; zeros when
    // deadtime == 0 (pmsm_env.py:785-791) -- the same dead-time shifted sequence the solver reads
    T ub[A];
    ub[0] = ub[1] = (T)0;
    if (aux.deadtime > 0) fetch((int)save_idx, ub);
    st[4] = ub[0];
    st[5] = ub[1];
    Mth<T>::sincos(ysv[2], &st[7], &st[6]);  // obs entries cos(eps), sin(eps) (pmsm_env.py:900-902)
  } else {
#pragma unroll
    for (int i = 0; i < NODE; ++i) st[i] = ysv[i];
  }
}

// T is the value type (float, double, or the packed f32x2: two environments per value); the dense-output records are
// plain scalars of type RT whatever T is.
template <class Env, int SOLVER, typename T, class Fetch, class Sink, typename RT>
EX_HD void sim_one_step(uint64_t w, T dt, T *y, T *f, const T *p, const typename Env::Pre &pre,
                        typename Env::Aux &aux, const SaveRec<RT> &rec0, const SaveRec<RT> *recs, long &save_idx,
                        Fetch &&fetch, Sink &&sink) {
  constexpr int NODE = Env::NODE, A = Env::A, NST = staged_values<Env>();
  const int base = sched_base(w);
  T y1[NODE], k[7][NODE];
  if constexpr (SOLVER == EXENV_EULER) {  // diffrax.Euler
    T u[A], f0[NODE];
    fetch_stage_(fetch, w, 0, u, 0);
    sim_rhs<Env, T>(y, u, p, pre, aux, f0);
#pragma unroll
    for (int i = 0; i < NODE; ++i) {
      k[0][i] = f0[i] * dt;
      y1[i] = y[i] + k[0][i];
    }
  } else {  // diffrax.Tsit5 with FSAL, derivative form (exenv_math.cuh: tsit5_step_f): k holds f_1 .. f_7
#pragma unroll
    for (int i = 0; i < NODE; ++i) k[0][i] = f[i];
    auto stage_rhs = [&](int stage, const T *ys, T *fo) {
      T u[A];
      fetch_stage_(fetch, w, stage, u, 0);
      sim_rhs<Env, T>(ys, u, p, pre, aux, fo);
    };
    tsit5_step_f<T, NODE>(y, dt, stage_rhs, y1, k);
#pragma unroll
    for (int i = 0; i < NODE; ++i) f[i] = k[6][i];
  }
  // SaveAt(ts): dense output at the scheduled abscissae; one save per step is the common case
  const int ns = sched_saves(w);
  if (ns > 0) {
    T st[NST];
    sim_save_one<Env, SOLVER, T>(rec0, y, y1, k, p, aux, save_idx, fetch, st);
    sink(save_idx, st);
    ++save_idx;
#pragma unroll 1
    for (int q = 1; q < ns; ++q) {
      const SaveRec<RT> rec = recs[save_idx];
      sim_save_one<Env, SOLVER, T>(rec, y, y1, k, p, aux, save_idx, fetch, st);
      sink(save_idx, st);
      ++save_idx;
    }
  }
#pragma unroll
  for (int i = 0; i < NODE; ++i) y[i] = y1[i];
}

// FSAL derivative for the very first step (diffrax evaluates it when `first_step` is set).
template <class Env, typename T, class Fetch>
EX_HD void sim_first_derivative(uint64_t w0, const T *y, T *f, const T *p, const typename Env::Pre &pre,
                                typename Env::Aux &aux, Fetch &&fetch) {
  T u[Env::A];
  fetch_stage_(fetch, w0, 0, u, 0);
  sim_rhs<Env, T>(y, u, p, pre, aux, f);
}

}  // namespace exenv
