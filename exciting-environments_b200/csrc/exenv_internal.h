// exenv_internal.h -- host-side structures shared by the C-ABI translation unit (exenv_api.cu) and the kernel
// launchers (exenv_step.cu, exenv_sim.cu).  Not part of the public ABI.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/exenv_b200.h"

namespace exenv {

void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launch_count;

// One tile of the persistent sim_ahead kernel: a run of solver steps whose ZOH actions and saved samples are
// staged in shared memory together.
struct SimTile {
  int step_begin, step_count;
  int act_lo, act_cnt;    // staged action-index range (indices into the ZOH'ed sequence)
  int save_begin, save_cnt;
};

// Environment-independent emulation of diffrax.diffeqsolve's constant-step loop in the plan dtype:
// per solver step the step length, the ZOH action index of every stage and the saves that fall into it.
struct SimSchedule {
  int64_t n_actions = 0, n_save = 0, n_steps = 0;
  int tile_steps = 0, n_tiles = 0, act_cap = 0, save_cap = 0;
  int win_need = 1;  // time-major kernel: consecutive staged action indices a single solver step reads (from its base)
  bool uploaded = false;  // every device copy below is complete
  // device copies
  uint64_t *d_words = nullptr;  // bits 0-31 min action index, 32-45 seven 2-bit stage deltas, 46-49 #saves
  void *d_dt = nullptr;         // T[n_steps]   tnext - tprev
  void *d_theta = nullptr;      // SaveRec<T>[n_save]  dense-output record of every save, in save order
  SimTile *d_tiles = nullptr;
  void *d_steprec = nullptr;    // TmStep<T>[chunks of 32, padded]: (word, dt, first-save record) per step, time-major kernels
  // host copies (tests / shape queries)
  std::vector<uint64_t> words;
  std::vector<double> dt, theta;
  std::vector<SimTile> tiles;
};

struct DevLut {
  void *blob = nullptr;  // [gx | gy | cells (7 words per node) | wx | wy (4 words per cell)] in the plan dtype, 16-byte padded sections
  size_t bytes = 0, off_gy = 0, off_cells = 0, off_wx = 0, off_wy = 0;
  int nx = 0, ny = 0;
  int regular = 0;       // both grids equidistant up to rounding (cell guess at most one off)
  double inv_dx = 0, inv_dy = 0, gx0 = 0, gy0 = 0, dx = 0, dy = 0;
  // fp32 plans whose grids are equidistant with exactly representable nodes: the coefficient form of the maps
  // (exenv_envs.cuh: LUT_COEF_STRIDE words per cell), read by the time-major trajectory kernels of Pmsm<float, 2>
  void *coef = nullptr;
  size_t coef_bytes = 0;
};

}  // namespace exenv

struct ExenvPlan {
  ExenvDesc desc;
  ExenvDims dims;
  int device = 0;
  int deadtime = 0;        // PMSM static_params.deadtime
  double pmsm_adv = 0;     // (deadtime + 0.5) * tau
  bool per_env = false;    // any pp[i] != NULL
  exenv::DevLut lut;
  std::mutex mu;
  std::map<std::tuple<int64_t, double, double, int>, exenv::SimSchedule> schedules;
  // exenv_sim_ahead_host: three copy/compute streams and their device workspaces, created on first use, kept for the
  // life of the plan (grown when a call needs more)
  std::mutex host_mu;
  cudaStream_t host_stream[3] = {nullptr, nullptr, nullptr};
  void *host_ws[3] = {nullptr, nullptr, nullptr};
  size_t host_ws_bytes = 0;
};

namespace exenv {

// per-dtype translation units (exenv_step_f32.cu, exenv_step_f64.cu, exenv_sim_f32.cu, exenv_sim_f64.cu)
int launch_step_euler_f32(const ExenvPlan &plan, const ExenvStepArgs &a, cudaStream_t stream);
int launch_step_euler_f64(const ExenvPlan &plan, const ExenvStepArgs &a, cudaStream_t stream);
int launch_step_tsit5_f32(const ExenvPlan &plan, const ExenvStepArgs &a, cudaStream_t stream);
int launch_step_tsit5_f64(const ExenvPlan &plan, const ExenvStepArgs &a, cudaStream_t stream);
int launch_reset_f32(const ExenvPlan &plan, const ExenvResetArgs &a, cudaStream_t stream);
int launch_reset_f64(const ExenvPlan &plan, const ExenvResetArgs &a, cudaStream_t stream);
int launch_epilogue_f32(const ExenvPlan &plan, const ExenvEpilogueArgs &a, cudaStream_t stream);
int launch_epilogue_f64(const ExenvPlan &plan, const ExenvEpilogueArgs &a, cudaStream_t stream);
int launch_gym_epilogue_f32(const ExenvPlan &plan, const ExenvGymArgs &a, cudaStream_t stream);
int launch_gym_epilogue_f64(const ExenvPlan &plan, const ExenvGymArgs &a, cudaStream_t stream);
int launch_sim_f32(const ExenvPlan &plan, const ExenvSimArgs &a, const SimSchedule &sch, cudaStream_t stream);
int launch_sim_f64(const ExenvPlan &plan, const ExenvSimArgs &a, const SimSchedule &sch, cudaStream_t stream);
int launch_simtm_f32(const ExenvPlan &plan, const ExenvSimArgs &a, const SimSchedule &sch, cudaStream_t stream);
int launch_simtm_f64(const ExenvPlan &plan, const ExenvSimArgs &a, const SimSchedule &sch, cudaStream_t stream);
// saturated PMSM, coefficient-form flux maps (exenv_simtm_pmsm_coef_f32.cu); plan.lut.coef must exist
int launch_simtm_pmsm_coef_f32(const ExenvPlan &plan, const ExenvSimArgs &a, const SimSchedule &sch, cudaStream_t stream);
int launch_simtm_packed_f32(const ExenvPlan &plan, const ExenvSimArgs &a, const SimSchedule &sch, cudaStream_t stream,
                            int envs_per_thread);
size_t sim_smem_bytes_f32(const ExenvPlan &plan, int step_cap, int act_cap, int save_cap);
size_t sim_smem_bytes_f64(const ExenvPlan &plan, int step_cap, int act_cap, int save_cap);

inline int launch_step(const ExenvPlan &plan, const ExenvStepArgs &a, cudaStream_t s) {
  if (plan.desc.solver == EXENV_TSIT5)
    return plan.desc.dtype == EXENV_F64 ? launch_step_tsit5_f64(plan, a, s) : launch_step_tsit5_f32(plan, a, s);
  return plan.desc.dtype == EXENV_F64 ? launch_step_euler_f64(plan, a, s) : launch_step_euler_f32(plan, a, s);
}
inline int launch_reset(const ExenvPlan &plan, const ExenvResetArgs &a, cudaStream_t s) {
  return plan.desc.dtype == EXENV_F64 ? launch_reset_f64(plan, a, s) : launch_reset_f32(plan, a, s);
}
inline int launch_sim(const ExenvPlan &plan, const ExenvSimArgs &a, const SimSchedule &sch, cudaStream_t s) {
  return plan.desc.dtype == EXENV_F64 ? launch_sim_f64(plan, a, sch, s) : launch_sim_f32(plan, a, sch, s);
}
inline int launch_simtm(const ExenvPlan &plan, const ExenvSimArgs &a, const SimSchedule &sch, cudaStream_t s) {
  return plan.desc.dtype == EXENV_F64 ? launch_simtm_f64(plan, a, sch, s) : launch_simtm_f32(plan, a, sch, s);
}
// Shared-memory bytes the sim kernel needs for tiles of (step_cap steps, act_cap actions, save_cap saves).
inline size_t sim_smem_bytes(const ExenvPlan &plan, int step_cap, int act_cap, int save_cap) {
  return plan.desc.dtype == EXENV_F64 ? sim_smem_bytes_f64(plan, step_cap, act_cap, save_cap)
                                      : sim_smem_bytes_f32(plan, step_cap, act_cap, save_cap);
}
int sim_block_envs(int64_t batch, int requested);

}  // namespace exenv
