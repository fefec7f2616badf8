/*
 * exenv_b200 -- C ABI of the B200-native batched dynamics path of exciting-environments.
 *
 * This is the drop-in boundary: plain C, raw device pointers and sizes, no torch / JAX types.  Every entry point
 * names the reference interface it replaces (paths relative to the reference's exciting_environments/ package):
 *
 *   exenv_step            <- CoreEnvironment.vmap_step            core_env.py:533-569  (-> step :393-425,
 *                            PMSM.step pmsm/pmsm_env.py:851-883, <Env>._ode_solver_step e.g. pendulum_env.py:152-194)
 *                            + the GymWrapper epilogue            gym_wrapper.py:117-126
 *   exenv_sim_ahead       <- CoreEnvironment.vmap_sim_ahead       core_env.py:571-616  (-> sim_ahead :427-488,
 *                            PMSM.sim_ahead pmsm_env.py:746-801, <Env>._ode_solver_simulate_ahead e.g.
 *                            pendulum_env.py:196-259) + vmap_generate_rew_trunc_term_ahead core_env.py:618-647
 *   exenv_reset_pack      <- CoreEnvironment.vmap_reset           core_env.py:664-687  (-> reset :361-391, init_state
 *                            e.g. pendulum_env.py:261-295) and vmap_generate_state_from_observation :689-705
 *   exenv_epilogue        <- vmap(generate_reward / truncated / terminated) gym_wrapper.py:117-126 and
 *                            vmap_generate_rew_trunc_term_ahead core_env.py:618-647 on materialised states
 *   exenv_gym_epilogue    <- GymWrapper.gym_step after its vmap_step: update_ref / generate_new_ref (gym_wrapper.py:
 *                            170-192, per-environment hold counters, device-side counter RNG) + the three epilogue vmaps
 *   exenv_sim_ahead_host  <- the same vmap_sim_ahead call made with HOST buffers (what a CPU-JAX caller holds);
 *                            pipelines H2D / kernel / D2H over batch chunks.
 *
 * The reference has no FFI of its own (pure Python on JAX); these symbols are what a `jax.ffi.ffi_call` custom-call
 * target binds (see INTEGRATION.md and csrc/xla_ffi_shim.cc).
 *
 * Conventions
 *   - dtype: all floating buffers of one plan share one dtype (EXENV_F32 / EXENV_F64), like jax_enable_x64.
 *   - State leaves are structure-of-arrays: one (B,) array per PhysicalState field, in the reference's dataclass
 *     field order.  action (B, A), obs (B, O'), actions (B, T, A), observations (B, N+1, O'), trajectory leaves
 *     (B, N+1): C-contiguous, batch-major, exactly the reference's array shapes.  O' = O + n_control.
 *     exenv_sim_ahead additionally offers EXENV_LAYOUT_TIME_MAJOR (structure of arrays over the batch: actions
 *     (T, B, A), observations (N+1, B, O'), trajectory leaves (N+1, B), reward / terminated (N, B), truncated
 *     (N+1, B, n_trunc)): the same logical arrays with the batch axis innermost, so that every warp access is one
 *     contiguous segment.  The host mirror returns them as permuted views with the reference's shapes.
 *   - Outputs may alias the matching inputs leaf-by-leaf (each thread reads its element before writing it).
 *   - Kernels never allocate or retain caller pointers past the call; launches are asynchronous on `stream`
 *     (a cudaStream_t passed as void*), no host synchronisation, CUDA-graph capturable (exenv_sim_ahead: after
 *     exenv_sim_ahead_prepare for the call shape; inside a capture an unprepared shape is an error, never a malloc).
 *     Every entry point runs on the plan's device and restores the calling thread's current device.
 *   - Return value: EXENV_OK or a negative ExenvStatus; exenv_last_error() gives the thread-local message.
 *     There is NO CPU fallback: without a CUDA device every compute entry point returns EXENV_ERR_CUDA.
 */
#ifndef EXENV_B200_H
#define EXENV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EXENV_ABI_VERSION 3
#define EXENV_MAX_PARAMS 32
#define EXENV_MAX_CONTROL 8
#define EXENV_MAX_PHYS 8
#define EXENV_MAX_ODE 4

/* EnvironmentRegistry members (registration.py:12-18) */
typedef enum ExenvEnv {
  EXENV_PENDULUM = 0,
  EXENV_MASS_SPRING_DAMPER = 1,
  EXENV_FLUID_TANK = 2,
  EXENV_CART_POLE = 3,
  EXENV_ACROBOT = 4,
  EXENV_PMSM = 5
} ExenvEnv;

/* diffrax.Euler() / diffrax.Tsit5() (the `solver=` constructor kwarg, core_env.py:41) */
typedef enum ExenvSolver { EXENV_EULER = 0, EXENV_TSIT5 = 1 } ExenvSolver;
typedef enum ExenvDtype { EXENV_F32 = 0, EXENV_F64 = 1 } ExenvDtype;

typedef enum ExenvStatus {
  EXENV_OK = 0,
  EXENV_ERR_INVALID = -1,     /* bad enum / shape / unsupported combination */
  EXENV_ERR_NULL = -2,        /* required pointer is NULL */
  EXENV_ERR_ALIGN = -3,       /* pointer not aligned to the element size */
  EXENV_ERR_UNSUPPORTED = -4, /* e.g. PMSM sim_ahead with obs_stepsize != action_stepsize (pmsm_env.py:787) */
  EXENV_ERR_CUDA = -5,        /* CUDA runtime error (message holds cudaGetErrorString) */
  EXENV_ERR_MAX_STEPS = -6,   /* diffrax `max_steps` reached (reference: RuntimeError from diffeqsolve) */
  EXENV_ERR_NOMEM = -7
} ExenvStatus;

#define EXENV_FLAG_SATURATED 1u /* PMSM(saturated=True) */

/*
 * Environment description = the reference's `EnvProperties` (core_env.py:245-251) + constructor kwargs.
 * Parameter slots, all in reference dataclass field order:
 *   p[0 .. n_static)                      StaticParams
 *   p[n_static + 2*i], p[.. + 2*i + 1]    physical_normalizations field i  (min, max)
 *   p[n_static + 2*D + 2*j], [.. + 1]     action_normalizations field j    (min, max)
 * StaticParams order: pendulum g,l,m | mass_spring_damper d,k,m | fluid_tank base_area,orifice_area,c_d,g |
 *   cart_pole mu_p,mu_c,l,m_p,m_c,g | acrobot g,l_1,l_2,m_1,m_2,l_c1,l_c2,I_1,I_2 | pmsm p,r_s,l_d,l_q,psi_p,u_dc,deadtime
 * pp[i] != NULL makes slot i a per-environment (B,) DEVICE array of the plan dtype (the reference's "jnp.array with
 * the length of batch_size", core_env.py:253-277); the pointer must stay valid while the plan is used.
 */
typedef struct ExenvDesc {
  int32_t env;       /* ExenvEnv */
  int32_t solver;    /* ExenvSolver */
  int32_t dtype;     /* ExenvDtype */
  uint32_t flags;    /* EXENV_FLAG_* */
  double tau;        /* control step, seconds (core_env.py:52) */
  int32_t n_control; /* len(control_state) */
  int32_t control[EXENV_MAX_CONTROL]; /* PhysicalState field indices, in control_state order */
  int32_t time_rule; /* sim_ahead step-size rule: 0 = tnext + dt0 (diffrax>=0.5), 1 = tnext + (tnext - tprev) */
  double p[EXENV_MAX_PARAMS];
  const void *pp[EXENV_MAX_PARAMS];
  /* PMSM saturated flux maps, HOST pointers copied at plan creation: grids gx (i_d, nx), gy (i_q, ny) and six maps
   * L_dd, L_dq, L_qd, L_qq, Psi_d, Psi_q, each [nx][ny] (= RegularGridInterpolator values[x_idx][y_idx],
   * pmsm_env.py:352-361), all float64. */
  const double *lut_gx;
  const double *lut_gy;
  const double *lut_maps;
  int32_t lut_nx;
  int32_t lut_ny;
} ExenvDesc;

typedef struct ExenvPlan *exenv_plan_t;

/* Static facts about an environment (D, ODE dim, A, O, number of static params, truncated width). */
typedef struct ExenvDims {
  int32_t n_phys, n_ode, n_act, n_obs, n_static, n_trunc;
} ExenvDims;

int exenv_abi_version(void);
/* Copies the calling thread's last error message (NUL terminated) and returns its length. */
size_t exenv_last_error(char *buf, size_t cap);
int exenv_env_dims(int32_t env, ExenvDims *out);
/* Number of CUDA kernels launched by this library in this process so far (bench.py's gpu_launches). */
uint64_t exenv_launch_count(void);

/* Plans own the device copies of LUTs and cached time schedules.  `device` = CUDA ordinal. */
int exenv_plan_create(const ExenvDesc *desc, int device, exenv_plan_t *out);
int exenv_plan_destroy(exenv_plan_t plan);

/* ---- vmap_step ------------------------------------------------------------------------------------------------ */
typedef struct ExenvStepArgs {
  int64_t batch;
  const void *state_in[EXENV_MAX_PHYS]; /* D leaves (B,) */
  void *state_out[EXENV_MAX_PHYS];      /* D leaves (B,) */
  const void *action;                   /* (B, A) normalised */
  void *obs;                            /* (B, O') */
  void *k_out[EXENV_MAX_ODE];           /* Tsit5: solver_state derivative leaves (B,), optional (NULL) */
  const void *ref[EXENV_MAX_PHYS];      /* reference leaves (B,), required iff n_control > 0 */
  void *reward;                         /* optional (B,)    -- gym_wrapper.py:117 */
  uint8_t *truncated;                   /* optional (B, n_trunc) */
  uint8_t *terminated;                  /* optional (B,) */
} ExenvStepArgs;
int exenv_step(exenv_plan_t plan, const ExenvStepArgs *args, void *stream);

/* ---- vmap_sim_ahead ------------------------------------------------------------------------------------------- */
typedef enum ExenvLayout {
  EXENV_LAYOUT_BATCH_MAJOR = 0, /* the reference's C-contiguous arrays: (B, T, A), (B, N+1, O'), (B, N+1) */
  EXENV_LAYOUT_TIME_MAJOR = 1   /* batch axis innermost: (T, B, A), (N+1, B, O'), (N+1, B); reward/terminated (N, B),
                                   truncated (N+1, B, n_trunc) */
} ExenvLayout;

typedef struct ExenvSimArgs {
  int64_t batch;
  int64_t n_actions;      /* T = actions.shape[1] */
  double obs_stepsize;    /* solver step and save interval (dt0) */
  double action_stepsize; /* zero-order-hold interval */
  int64_t max_steps;      /* diffrax default 4096; <= 0 lifts the limit */
  const void *state_in[EXENV_MAX_PHYS]; /* D leaves (B,) */
  const void *actions;                  /* (B, T, A) normalised */
  void *obs;                            /* (B, N+1, O') */
  void *traj[EXENV_MAX_PHYS];           /* D leaves (B, N+1); a NULL leaf is skipped */
  void *last[EXENV_MAX_PHYS];           /* D leaves (B,), optional */
  void *k_last[EXENV_MAX_ODE];          /* Tsit5: solver.init derivative at (t1, y_last), optional */
  const void *ref[EXENV_MAX_PHYS];
  void *reward;        /* optional (B, N)            -- vmap_generate_rew_trunc_term_ahead */
  uint8_t *truncated;  /* optional (B, N+1, n_trunc) */
  uint8_t *terminated; /* optional (B, N) */
  int32_t tile_steps;  /* solver steps staged per shared-memory tile; 0 = auto */
  int32_t block_envs;  /* batch-major: reserved (128 environments per CTA); time-major: environments per thread
                          (1, 2 or 4; 0 = auto) */
  int32_t layout;      /* ExenvLayout of actions / obs / traj / reward / truncated / terminated */
  int32_t reserved_;
} ExenvSimArgs;
/* N+1 = 1 + int(action_stepsize * T / obs_stepsize) (pendulum_env.py:222-225) and the solver-step count of the
 * emulated diffeqsolve loop in the plan dtype. */
int exenv_sim_ahead_shape(exenv_plan_t plan, int64_t n_actions, double obs_stepsize, double action_stepsize,
                          int64_t *n_save, int64_t *n_steps);
/* Builds and uploads the time schedule of this call shape (host work, cudaMalloc, one synchronous cudaMemcpy) ahead of
 * time, so that exenv_sim_ahead itself only launches: required before the first exenv_sim_ahead of a shape that is issued
 * inside a CUDA-graph capture / an XLA custom call (which must not allocate or synchronise); otherwise optional --
 * outside a capture exenv_sim_ahead prepares lazily.  tile_steps as in ExenvSimArgs (0 = auto).  Also returns the shape
 * like exenv_sim_ahead_shape (pointers may be NULL). */
int exenv_sim_ahead_prepare(exenv_plan_t plan, int64_t n_actions, double obs_stepsize, double action_stepsize,
                            int32_t tile_steps, int64_t *n_save, int64_t *n_steps);
int exenv_sim_ahead(exenv_plan_t plan, const ExenvSimArgs *args, void *stream);
/* Same call with HOST buffers (pinned or pageable); chunk_envs environments per pipeline stage (0 = auto).
 * Host buffers are always in the reference's batch-major layout (layout must be EXENV_LAYOUT_BATCH_MAJOR). */
int exenv_sim_ahead_host(exenv_plan_t plan, const ExenvSimArgs *host_args, int64_t chunk_envs);

/* ---- reward / truncated / terminated of given states -------------------------------------------------------------
 * generate_reward / generate_truncated / generate_terminated vmapped over `batch` states (gym_wrapper.py:117-126) and,
 * applied to the flattened (B * (N+1),) leaves of a trajectory, CoreEnvironment.vmap_generate_rew_trunc_term_ahead
 * (core_env.py:618-647; PMSM pmsm_env.py:803-849).  Shared parameters only. */
typedef struct ExenvEpilogueArgs {
  int64_t batch;
  const void *state[EXENV_MAX_PHYS]; /* D leaves (batch,) */
  const void *ref[EXENV_MAX_PHYS];   /* D reference leaves (batch,), required iff n_control > 0 */
  void *reward;                      /* optional (batch,) */
  uint8_t *truncated;                /* optional (batch, n_trunc) */
  uint8_t *terminated;               /* optional (batch,) */
} ExenvEpilogueArgs;
int exenv_epilogue(exenv_plan_t plan, const ExenvEpilogueArgs *args, void *stream);

/* ---- GymWrapper.gym_step after vmap_step: reference generator + reward / terminated / truncated ------------------------
 * gym_wrapper.py:111-126: `update_ref` (:170-175) -- for every environment whose hold counter is 0 the tracked
 * (control_state) reference fields become the physical state of a freshly drawn random initial state and a new hold time
 * is drawn in [hold_steps_min, hold_steps_max) (`generate_new_ref`, :177-192), then every counter is decremented --
 * followed by generate_reward / generate_terminated / generate_truncated against the UPDATED reference (the observation
 * returned by the step itself still shows the previous reference, as in the reference).  Random numbers: Philox4x32-10,
 * one stream per environment -- rng[2 b] is the environment's key, rng[2 b + 1] its draw counter (advanced by the
 * kernel).  Distributions are those of each environment's init_state(rng); the bit stream is not JAX's threefry.
 * hold_steps == NULL switches the generator off (plain epilogue).  force_new != 0: draw for every environment and do not
 * decrement (GymWrapper.reset with rng_ref, :156-160).  Shared parameters only. */
typedef struct ExenvGymArgs {
  int64_t batch;
  const void *state[EXENV_MAX_PHYS]; /* D leaves (B,) */
  void *ref[EXENV_MAX_PHYS];         /* D reference leaves (B,), control_state leaves updated in place */
  int32_t *hold_steps;               /* (B,) in/out, or NULL */
  uint64_t *rng;                     /* (B, 2) in/out, required iff hold_steps != NULL */
  int32_t hold_steps_min, hold_steps_max;
  int32_t force_new, reserved_;
  void *reward;        /* optional (B,) */
  uint8_t *truncated;  /* optional (B, n_trunc) */
  uint8_t *terminated; /* optional (B,) */
} ExenvGymArgs;
int exenv_gym_epilogue(exenv_plan_t plan, const ExenvGymArgs *args, void *stream);

/* Host-only diagnostic (needs no CUDA device): the emulated diffrax.diffeqsolve time bookkeeping for `desc` -- per
 * solver step the packed ZOH-index word and step length, per saved sample the dense-output abscissa, and the tile
 * table (6 ints per tile: step_begin, step_count, act_lo, act_cnt, save_begin, save_cnt).  Output arrays may be NULL
 * to query the sizes first.  tile_steps <= 0 selects 32. */
typedef struct ExenvScheduleInfo {
  int64_t n_save, n_steps;
  int32_t n_tiles, act_cap, save_cap, tile_steps;
} ExenvScheduleInfo;
int exenv_schedule_build(const ExenvDesc *desc, int64_t n_actions, double obs_stepsize, double action_stepsize,
                         int32_t tile_steps, ExenvScheduleInfo *info, uint64_t *words, double *dt, double *theta,
                         int32_t *tiles);

/* Host-only diagnostic (needs no CUDA device): the two thresholds the fused trajectory epilogue compares a state with
 * instead of evaluating `abs(2 * (x - min) / (max - min) - 1) > 1` (generate_truncated, e.g. pendulum_env.py:381-385 with
 * utils.py:13-14): the flag is `x >= *hi_first || x <= *lo_last`, found by bisection over the floating-point numbers of
 * `dtype` with the reference expression itself.  Returns EXENV_ERR_INVALID for a degenerate range. */
int exenv_trunc_thresholds(int32_t dtype, double min, double max, double *lo_last, double *hi_first);

/* ---- vmap_reset state packing / generate_state_from_observation ----------------------------------------------- */
typedef enum ExenvResetMode {
  EXENV_RESET_DEFAULT = 0,    /* init_state(rng=None) defaults */
  EXENV_RESET_FROM_NORM = 1,  /* values = (B, D) normalised physical state drawn by the host RNG (PMSM: physical) */
  EXENV_RESET_FROM_OBS = 2,   /* values = (B, O') observation -> state (generate_state_from_observation) */
  EXENV_RESET_FROM_STATE = 3  /* values unused; state_out holds initial_state already: only obs is generated */
} ExenvResetMode;
typedef struct ExenvResetArgs {
  int64_t batch;
  int32_t mode;
  const void *values;
  void *state_out[EXENV_MAX_PHYS]; /* D leaves (B,) */
  void *ref_out[EXENV_MAX_PHYS];   /* D leaves (B,), optional: NaN or de-normalised references */
  void *obs;                       /* (B, O') */
} ExenvResetArgs;
int exenv_reset_pack(exenv_plan_t plan, const ExenvResetArgs *args, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* EXENV_B200_H */
