// exenv_api.cu -- the C ABI (include/exenv_b200.h): plan management, argument validation, the host-side emulation of
// diffrax.diffeqsolve's time bookkeeping (build_schedule) and the host-buffer pipeline used for end-to-end numbers.
// No CPU compute fallback exists: every compute entry point needs a CUDA device and launches the sm_100a kernels of
// exenv_step.cu / exenv_sim.cu.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

#include "exenv_internal.h"
#include "exenv_sim_core.cuh"

namespace exenv {

static thread_local char g_err[512] = {0};
std::atomic<uint64_t> g_launch_count{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

#define EX_CUDA(call)                                                        \
  do {                                                                       \
    cudaError_t e_ = (call);                                                 \
    if (e_ != cudaSuccess) {                                                 \
      set_error("%s failed: %s", #call, cudaGetErrorString(e_));             \
      return EXENV_ERR_CUDA;                                                 \
    }                                                                        \
  } while (0)

// Every entry point that launches, allocates or frees runs on the plan's device and leaves the calling thread's current
// device as it found it (the caller -- torch, XLA -- owns that state).
struct DeviceGuard {
  int prev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) err = cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};
#define EX_ON_DEVICE(plan)                                                                              \
  DeviceGuard guard_((plan)->device);                                                                   \
  if (guard_.err != cudaSuccess) {                                                                      \
    set_error("cannot select CUDA device %d: %s", (plan)->device, cudaGetErrorString(guard_.err));      \
    return EXENV_ERR_CUDA;                                                                              \
  }
// No C++ exception crosses the C ABI (std::vector / std::map allocations of the schedule builder can throw).
#define EX_ABI_TRY try {
#define EX_ABI_CATCH                                             \
  }                                                              \
  catch (const std::bad_alloc &) {                               \
    set_error("out of host memory");                             \
    return EXENV_ERR_NOMEM;                                      \
  }                                                              \
  catch (const std::exception &e) {                              \
    set_error("internal error: %s", e.what());                   \
    return EXENV_ERR_INVALID;                                    \
  }                                                              \
  catch (...) {                                                  \
    set_error("internal error");                                 \
    return EXENV_ERR_INVALID;                                    \
  }

static const ExenvDims kDims[6] = {
    /* pendulum   */ {2, 2, 1, 2, 3, 2},
    /* msd        */ {2, 2, 1, 2, 3, 2},
    /* fluid tank */ {1, 1, 1, 1, 4, 1},
    /* cart pole  */ {4, 4, 1, 4, 6, 4},
    /* acrobot    */ {4, 4, 1, 4, 9, 4},
    /* pmsm       */ {7, 3, 2, 8, 7, 1},
};

// ---- diffeqsolve time bookkeeping, emulated in the solve dtype T --------------------------------------------------
// Restates (diffrax==0.7.0 is not vendored by the reference):
//   t1 = action_stepsize * n_actions; ts = jnp.linspace(0, t1, 1 + int(t1 / obs_stepsize))   (pendulum_env.py:222-225)
//   tprev = 0; tnext = clip_to_end(0 + dt0)
//   while tprev < t1: step over [tprev, tnext]; save every ts[i] <= tnext by dense output over [tprev, tnext];
//                     tprev' = min(tnext, t1); tnext' = clip_to_end(tnext + dt0)   (ConstantStepSize, dt0 = obs_stepsize)
//   stage i of the solver evaluates the ZOH closure at tprev + c_i * (tnext - tprev):
//                     index = int(t / action_stepsize), clamped to the last action (pendulum_env.py:215-216)
// Everything is plain IEEE arithmetic in T (this file is compiled with -ffp-contract=off).
template <typename T>
static int build_schedule_t(const ExenvPlan &plan, int64_t n_actions, double obs_dt_d, double act_dt_d,
                            int tile_steps, SimSchedule &out) {
  const bool tsit = plan.desc.solver == EXENV_TSIT5;
  const bool pmsm = plan.desc.env == EXENV_PMSM;
  if (!(obs_dt_d > 0) || !(act_dt_d > 0) || n_actions < 1) {
    set_error("sim_ahead: step sizes must be positive and n_actions >= 1");
    return EXENV_ERR_INVALID;
  }
  const double t1_d = act_dt_d * (double)n_actions;
  const int64_t N = (int64_t)(t1_d / obs_dt_d);
  if (N < 1 || N > (int64_t)1 << 30) {
    set_error("sim_ahead: unsupported number of saved samples %lld", (long long)(N + 1));
    return EXENV_ERR_INVALID;
  }
  const T t1 = (T)t1_d, dt0 = (T)obs_dt_d, adt = (T)act_dt_d;
  const T tol = sizeof(T) == 8 ? (T)1e-10 : (T)1e-6;
  auto clip_to_end = [&](T tn) -> T { return (tn > t1 - tol) ? t1 : tn; };
  std::vector<T> ts((size_t)N + 1);
  for (int64_t i = 0; i < N; ++i) {
    const T s = (T)i / (T)N;
    ts[(size_t)i] = (T)0 * ((T)1 - s) + t1 * s;
  }
  ts[(size_t)N] = t1;
  static const double c_d[7] = {0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0};

  out = SimSchedule();
  out.n_actions = n_actions;
  out.n_save = N + 1;
  T tprev = (T)0, tnext = clip_to_end((T)0 + dt0);
  int64_t si = 0;
  const int64_t hard_cap = 4 * (N + 8);
  std::vector<int> save_count;
  while (tprev < t1) {
    if ((int64_t)out.words.size() > hard_cap) {
      set_error("sim_ahead: the emulated solver loop does not terminate (dt0 too small for the dtype?)");
      return EXENV_ERR_INVALID;
    }
    const T dt = tnext - tprev;
    int64_t idx[7];
    const int n_st = tsit ? 7 : 1;
    for (int s = 0; s < 7; ++s) {
      if (s < n_st) {
        const T ti = tprev + (T)c_d[s] * dt;
        int64_t q = (int64_t)(ti / adt);
        if (q < 0) q = 0;
        if (q > n_actions - 1) q = n_actions - 1;
        idx[s] = q;
      } else {
        idx[s] = idx[0];
      }
    }
    int64_t base = idx[0];
    for (int s = 1; s < 7; ++s) base = std::min(base, idx[s]);
    uint64_t w = (uint64_t)(uint32_t)base;
    int64_t dmax = 0;
    for (int s = 0; s < 7; ++s) {
      const int64_t d = idx[s] - base;
      if (d > 3) {
        set_error("sim_ahead: a solver step spans more than 4 action intervals (obs_stepsize > action_stepsize?)");
        return EXENV_ERR_UNSUPPORTED;
      }
      w |= (uint64_t)d << (32 + 2 * s);
      dmax = std::max(dmax, d);
    }
    w |= (uint64_t)dmax << 50;
    int ns = 0;
    while (si <= N && ts[(size_t)si] <= tnext) {
      const T th = (tprev == tnext) ? (T)0 : (ts[(size_t)si] - tprev) / (tnext - tprev);
      out.theta.push_back((double)th);
      ++si;
      ++ns;
    }
    if (ns > 15) {
      set_error("sim_ahead: more than 15 saves inside one solver step");
      return EXENV_ERR_UNSUPPORTED;
    }
    w |= (uint64_t)ns << 46;
    {  // window of consecutive staged indices this step reads, counted from `base` (time-major kernel)
      int64_t need = dmax + 1;
      if (pmsm && plan.deadtime > 0 && ns > 0) {  // buffer leaves of saved state j read staged index j
        const int64_t first = si - ns;
        if (first < base) {
          set_error("sim_ahead: a saved sample precedes the zero-order-hold index of its step");
          return EXENV_ERR_UNSUPPORTED;
        }
        need = std::max<int64_t>(need, si - 1 - base + 1);
      }
      out.win_need = std::max(out.win_need, (int)need);
    }
    out.words.push_back(w);
    out.dt.push_back((double)dt);
    save_count.push_back(ns);
    T ntprev = tnext;
    T ntnext = plan.desc.time_rule ? tnext + (tnext - tprev) : tnext + dt0;
    if (ntprev > t1) ntprev = t1;
    ntnext = clip_to_end(ntnext);
    tprev = ntprev;
    tnext = ntnext;
  }
  out.n_steps = (int64_t)out.words.size();
  if (si != N + 1) {
    set_error("sim_ahead: internal error, %lld of %lld samples scheduled", (long long)si, (long long)(N + 1));
    return EXENV_ERR_INVALID;
  }

  // ---- tiles: at most `tile_steps` solver steps AND at most `tile_steps` saves each (a step with several saves,
  // e.g. the first one which also saves ts[0], must not push a tile over the staging capacity the fast row
  // writer is built for)
  out.tile_steps = tile_steps;
  int64_t save_run = 0;
  for (int64_t n0 = 0; n0 < out.n_steps;) {
    SimTile t{};
    t.step_begin = (int)n0;
    int64_t lo = INT64_MAX, hi = -1;
    int sc = 0, cntsteps = 0;
    for (int64_t n = n0; n < out.n_steps && cntsteps < tile_steps; ++n) {
      const int add = save_count[(size_t)n];
      if (cntsteps > 0 && sc + add > tile_steps) break;
      const uint64_t w = out.words[(size_t)n];
      const int64_t base = (int64_t)(w & 0xffffffffull);
      int64_t mx = base;
      for (int s = 0; s < 7; ++s) mx = std::max(mx, base + (int64_t)((w >> (32 + 2 * s)) & 3));
      lo = std::min(lo, base);
      hi = std::max(hi, mx);
      sc += add;
      ++cntsteps;
    }
    t.step_count = cntsteps;
    t.save_begin = (int)save_run;
    t.save_cnt = sc;
    if (pmsm && plan.deadtime > 0 && sc > 0) {  // buffer leaves of the saved states index the same staged sequence
      lo = std::min<int64_t>(lo, save_run);
      hi = std::max<int64_t>(hi, save_run + sc - 1);
    }
    n0 += cntsteps;
    const bool is_last = n0 >= out.n_steps;
    if (is_last) hi = std::max<int64_t>(hi, n_actions - 1);  // solver.init at t1 reads the clamped last action
    save_run += sc;
    t.act_lo = (int)lo;
    t.act_cnt = (int)(hi - lo + 1);
    out.act_cap = std::max(out.act_cap, t.act_cnt);
    out.save_cap = std::max(out.save_cap, t.save_cnt);
    out.tiles.push_back(t);
  }
  out.n_tiles = (int)out.tiles.size();
  if (out.save_cap < 1) out.save_cap = 1;
  return EXENV_OK;
}

template <typename T>
static int upload_records(const ExenvPlan &plan, SimSchedule &s) {
  std::vector<SaveRec<T>> recs(s.theta.size() + 1);  // + one look-ahead entry
  for (int j = 0; j < 8; ++j) recs.back().v[j] = (T)0;
  make_saverecs<T>(plan.desc.solver == EXENV_TSIT5, (long)s.n_steps, s.words.data(), s.dt.data(), s.theta.data(), recs);
  std::vector<T> dt(s.dt.begin(), s.dt.end());
  dt.push_back((T)0);  // look-ahead padding (the batch-major kernel fetches step n+1 while integrating step n)
  // time-major kernels: one record per step, padded to whole 32-step chunks plus one chunk the staging runs ahead
  const size_t n_rec = ((size_t)s.n_steps / TM_CHUNK + 2) * TM_CHUNK;
  std::vector<TmStep<T>> steps(n_rec);
  memset(steps.data(), 0, steps.size() * sizeof(TmStep<T>));
  size_t save = 0;
  for (size_t i = 0; i < (size_t)s.n_steps; ++i) {
    steps[i].w = s.words[i];
    steps[i].dt = (T)s.dt[i];
    const int ns = (int)((s.words[i] >> 46) & 15ull);
    if (ns > 0)
      for (int j = 0; j < 8; ++j) steps[i].wts[j] = recs[save].v[j];
    save += (size_t)ns;
  }
  EX_CUDA(cudaMalloc(&s.d_steprec, steps.size() * sizeof(TmStep<T>)));
  EX_CUDA(cudaMemcpy(s.d_steprec, steps.data(), steps.size() * sizeof(TmStep<T>), cudaMemcpyHostToDevice));
  EX_CUDA(cudaMalloc(&s.d_dt, dt.size() * sizeof(T)));
  EX_CUDA(cudaMalloc(&s.d_theta, recs.size() * sizeof(SaveRec<T>)));
  EX_CUDA(cudaMemcpy(s.d_dt, dt.data(), dt.size() * sizeof(T), cudaMemcpyHostToDevice));
  EX_CUDA(cudaMemcpy(s.d_theta, recs.data(), recs.size() * sizeof(SaveRec<T>), cudaMemcpyHostToDevice));
  return EXENV_OK;
}

static int upload_schedule(const ExenvPlan &plan, SimSchedule &s) {
  std::vector<uint64_t> words(s.words);
  words.push_back(0);  // look-ahead padding
  EX_CUDA(cudaMalloc(&s.d_words, words.size() * sizeof(uint64_t)));
  EX_CUDA(cudaMalloc(&s.d_tiles, s.tiles.size() * sizeof(SimTile)));
  EX_CUDA(cudaMemcpy(s.d_words, words.data(), words.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
  EX_CUDA(cudaMemcpy(s.d_tiles, s.tiles.data(), s.tiles.size() * sizeof(SimTile), cudaMemcpyHostToDevice));
  return plan.desc.dtype == EXENV_F64 ? upload_records<double>(plan, s) : upload_records<float>(plan, s);
}

static void free_schedule(SimSchedule &s) {
  cudaFree(s.d_words);
  cudaFree(s.d_dt);
  cudaFree(s.d_theta);
  cudaFree(s.d_tiles);
  cudaFree(s.d_steprec);
  s.d_steprec = nullptr;
  s.uploaded = false;
  s.d_words = nullptr;
  s.d_dt = s.d_theta = nullptr;
  s.d_tiles = nullptr;
}

static int build_schedule(const ExenvPlan &plan, int64_t n_actions, double obs_dt, double act_dt, int tile_steps,
                          SimSchedule &out) {
  return plan.desc.dtype == EXENV_F64 ? build_schedule_t<double>(plan, n_actions, obs_dt, act_dt, tile_steps, out)
                                      : build_schedule_t<float>(plan, n_actions, obs_dt, act_dt, tile_steps, out);
}

// Tile length.  Longer tiles mean longer contiguous row segments in HBM (128 B per leaf row at 32 fp32 saves) but
// more shared memory per CTA, i.e. fewer resident warps.  Measured on B200 (tools/sweep_tiles.sh, profiles/README.md):
// light right-hand sides (mass-spring-damper, fluid tank, any Euler run) are bound by the memory access pattern and
// want 32-step tiles at 4 CTAs/SM; transcendental-heavy ones (pendulum, cart-pole Tsit5) want more resident warps;
// acrobot and PMSM are instruction bound and take the smallest tiles.
static int auto_tile_steps(const ExenvPlan &plan, int /*E*/) {
  const int env = plan.desc.env;
  const bool tsit = plan.desc.solver == EXENV_TSIT5;
  size_t budget = 56 * 1024;
  if (env == EXENV_ACROBOT || env == EXENV_PMSM)
    budget = 36 * 1024;
  else if (tsit && (env == EXENV_PENDULUM || env == EXENV_CART_POLE))
    budget = 44 * 1024;
  const int cand[] = {32, 24, 16, 12, 8};
  for (int k : cand)
    if (sim_smem_bytes(plan, k, k + 3, k) <= budget) return k;
  const int small[] = {8, 4, 2, 1};
  for (int k : small)
    if (sim_smem_bytes(plan, k, k + 3, k) <= 200 * 1024) return k;
  return 1;
}

// The schedule of a (n_actions, obs_stepsize, action_stepsize) shape is built on the host once and, for the launching
// entry points, uploaded once (cudaMalloc + synchronous cudaMemcpy).  `stream` != NULL names the stream a launch is about
// to use: if it is being captured into a CUDA graph the upload must have happened already (exenv_sim_ahead_prepare) --
// nothing allocates or synchronises inside a capture.
static int get_schedule(ExenvPlan *plan, int64_t n_actions, double obs_dt, double act_dt, int tile_steps,
                        bool need_device, SimSchedule **out, const cudaStream_t *stream = nullptr) {
  std::lock_guard<std::mutex> lock(plan->mu);
  auto key = std::make_tuple(n_actions, obs_dt, act_dt, tile_steps);
  auto it = plan->schedules.find(key);
  if (it == plan->schedules.end()) {
    SimSchedule s;
    int rc = build_schedule(*plan, n_actions, obs_dt, act_dt, tile_steps, s);
    if (rc != EXENV_OK) return rc;
    it = plan->schedules.emplace(key, std::move(s)).first;
  }
  if (need_device && !it->second.uploaded) {
    if (stream) {
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      if (cudaStreamIsCapturing(*stream, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) {
        set_error("exenv_sim_ahead inside a stream capture needs exenv_sim_ahead_prepare(plan, %lld, %g, %g) first "
                  "(the time schedule of a new shape is uploaded with cudaMalloc + cudaMemcpy)",
                  (long long)n_actions, obs_dt, act_dt);
        return EXENV_ERR_INVALID;
      }
    }
    int rc = upload_schedule(*plan, it->second);
    if (rc != EXENV_OK) {  // a half-uploaded schedule must not be launched with: release it, the next call retries
      free_schedule(it->second);
      return rc;
    }
    it->second.uploaded = true;
  }
  *out = &it->second;
  return EXENV_OK;
}

int sim_block_envs(int64_t batch, int requested) {
  int E = requested > 0 ? ((requested + 31) / 32) * 32 : 128;
  if (batch < E) E = (int)((batch + 31) / 32) * 32;
  if (E < 32) E = 32;
  if (E > 256) E = 256;
  return E;
}

static bool aligned_to(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

}  // namespace exenv

using namespace exenv;

extern "C" {

int exenv_abi_version(void) { return EXENV_ABI_VERSION; }

size_t exenv_last_error(char *buf, size_t cap) {
  const size_t n = strlen(g_err);
  if (buf && cap) {
    const size_t m = std::min(n, cap - 1);
    memcpy(buf, g_err, m);
    buf[m] = 0;
  }
  return n;
}

int exenv_env_dims(int32_t env, ExenvDims *out) {
  if (env < 0 || env > 5 || !out) {
    set_error("exenv_env_dims: bad env id %d", env);
    return EXENV_ERR_INVALID;
  }
  *out = kDims[env];
  return EXENV_OK;
}

uint64_t exenv_launch_count(void) { return g_launch_count.load(); }

int exenv_plan_create(const ExenvDesc *desc, int device, exenv_plan_t *out) {
  EX_ABI_TRY
  if (!desc || !out) {
    set_error("exenv_plan_create: NULL argument");
    return EXENV_ERR_NULL;
  }
  if (desc->env < 0 || desc->env > 5 || (desc->solver != EXENV_EULER && desc->solver != EXENV_TSIT5) ||
      (desc->dtype != EXENV_F32 && desc->dtype != EXENV_F64)) {
    set_error("exenv_plan_create: bad env/solver/dtype (%d, %d, %d)", desc->env, desc->solver, desc->dtype);
    return EXENV_ERR_INVALID;
  }
  const ExenvDims dm = kDims[desc->env];
  if (desc->n_control < 0 || desc->n_control > EXENV_MAX_CONTROL) {
    set_error("exenv_plan_create: n_control out of range");
    return EXENV_ERR_INVALID;
  }
  for (int j = 0; j < desc->n_control; ++j)
    if (desc->control[j] < 0 || desc->control[j] >= dm.n_phys) {
      set_error("exenv_plan_create: control_state index %d is not a PhysicalState field", desc->control[j]);
      return EXENV_ERR_INVALID;
    }
  if (!(desc->tau > 0)) {
    set_error("exenv_plan_create: tau must be positive");
    return EXENV_ERR_INVALID;
  }
  const bool sat = desc->flags & EXENV_FLAG_SATURATED;
  if (sat && desc->env != EXENV_PMSM) {
    set_error("exenv_plan_create: the saturated flag is only valid for PMSM");
    return EXENV_ERR_INVALID;
  }
  ExenvPlan *p = new (std::nothrow) ExenvPlan();
  if (!p) return EXENV_ERR_NOMEM;
  p->desc = *desc;
  p->dims = dm;
  p->dims.n_trunc = (desc->env == EXENV_FLUID_TANK || desc->env == EXENV_PMSM) ? 1 : dm.n_obs + desc->n_control;
  p->device = device;
  const int np = dm.n_static + 2 * dm.n_phys + 2 * dm.n_act;
  for (int i = 0; i < np; ++i) p->per_env |= desc->pp[i] != nullptr;
  for (int i = np; i < EXENV_MAX_PARAMS; ++i) p->desc.pp[i] = nullptr;
  if (desc->env == EXENV_PMSM) {
    if (p->per_env) {  // the reference hashes PMSM env_properties as a static jit argument (pmsm_env.py:525, :851)
      set_error("PMSM does not support per-environment parameter arrays");
      delete p;
      return EXENV_ERR_UNSUPPORTED;
    }
    const double dead = desc->p[6];
    p->deadtime = (int)dead;
    p->pmsm_adv = (dead + 0.5) * desc->tau;
    if (p->deadtime < 0 || p->deadtime > 1) {
      set_error("PMSM deadtime must be 0 or 1 (the reference's buffers hold one action)");
      delete p;
      return EXENV_ERR_UNSUPPORTED;
    }
  }
  DeviceGuard guard(device);  // restores the caller's current device on every return path
  if (guard.err != cudaSuccess) {
    set_error("exenv_plan_create: no usable CUDA device %d (%s); this library has no CPU fallback", device,
              cudaGetErrorString(guard.err));
    delete p;
    return EXENV_ERR_CUDA;
  }
  if (sat) {
    if (!desc->lut_gx || !desc->lut_gy || !desc->lut_maps || desc->lut_nx < 2 || desc->lut_ny < 2) {
      set_error("saturated PMSM needs lut_gx / lut_gy / lut_maps");
      delete p;
      return EXENV_ERR_NULL;
    }
    const int nx = desc->lut_nx, ny = desc->lut_ny;
    const size_t es = desc->dtype == EXENV_F64 ? 8 : 4;
    auto pad16 = [](size_t n) { return (n + 15) & ~(size_t)15; };
    DevLut &L = p->lut;
    L.nx = nx;
    L.ny = ny;
    L.off_gy = pad16((size_t)nx * es);
    L.off_cells = L.off_gy + pad16((size_t)ny * es);
    constexpr int CS = LUT_BLOB_CELL_STRIDE;
    L.off_wx = L.off_cells + pad16((size_t)nx * ny * CS * es);
    L.off_wy = L.off_wx + pad16((size_t)(nx - 1) * 4 * es);
    L.bytes = L.off_wy + pad16((size_t)(ny - 1) * 4 * es);
    L.gx0 = desc->lut_gx[0];
    L.gy0 = desc->lut_gy[0];
    L.dx = desc->lut_gx[1] - desc->lut_gx[0];
    L.dy = desc->lut_gy[1] - desc->lut_gy[0];
    L.inv_dx = 1.0 / L.dx;
    L.inv_dy = 1.0 / L.dy;
    std::vector<unsigned char> blob(L.bytes, 0);
    auto put = [&](size_t off, size_t i, double v) {
      if (es == 8)
        reinterpret_cast<double *>(blob.data() + off)[i] = v;
      else
        reinterpret_cast<float *>(blob.data() + off)[i] = (float)v;
    };
    for (int i = 0; i < nx; ++i) put(0, (size_t)i, desc->lut_gx[i]);
    for (int j = 0; j < ny; ++j) put(L.off_gy, (size_t)j, desc->lut_gy[j]);
    // per-cell {width, reciprocal} in the plan dtype, computed like the kernel would: d = RN(g[i+1] - g[i]) on the
    // dtype-rounded grid values, r = RN(1 / d)
    auto widths = [&](const double *g, int n, size_t off) {
      bool regular = true;
      const double d0 = (g[n - 1] - g[0]) / (n - 1);
      for (int i = 0; i + 1 < n; ++i) {
        put(off, (size_t)4 * i, g[i]);
        if (es == 8) {
          const double d = g[i + 1] - g[i];
          put(off, (size_t)4 * i + 1, d);
          put(off, (size_t)4 * i + 2, 1.0 / d);
        } else {
          const float d = (float)g[i + 1] - (float)g[i];
          put(off, (size_t)4 * i + 1, d);
          put(off, (size_t)4 * i + 2, 1.0f / d);
        }
        regular = regular && std::fabs(g[i] - (g[0] + i * d0)) < 0.25 * std::fabs(d0);
      }
      return regular && d0 > 0;
    };
    const bool rx = widths(desc->lut_gx, nx, L.off_wx), ry = widths(desc->lut_gy, ny, L.off_wy);
    L.regular = rx && ry;
    for (int q = 0; q < 6; ++q)
      for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j)
          put(L.off_cells, ((size_t)i * ny + j) * CS + q, desc->lut_maps[((size_t)q * nx + i) * ny + j]);
    if (cudaMalloc(&L.blob, L.bytes) != cudaSuccess ||
        cudaMemcpy(L.blob, blob.data(), L.bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
      set_error("exenv_plan_create: LUT upload failed: %s", cudaGetErrorString(cudaGetLastError()));
      delete p;
      return EXENV_ERR_CUDA;
    }
    if (es == 4 && L.regular && lut_grid_exact_f32(desc->lut_gx, nx) && lut_grid_exact_f32(desc->lut_gy, ny)) {
      // coefficient form for the fp32 time-major trajectory kernels; used when it fits one CTA's shared memory next to
      // the schedule buffers (launch_simtm_pmsm_coef_f32 decides)
      std::vector<float> coef((size_t)(nx - 1) * (ny - 1) * LUT_COEF_STRIDE);
      lut_coef_fill(desc->lut_maps, nx, ny, coef.data());
      L.coef_bytes = coef.size() * sizeof(float);
      if (cudaMalloc(&L.coef, L.coef_bytes) != cudaSuccess ||
          cudaMemcpy(L.coef, coef.data(), L.coef_bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("exenv_plan_create: LUT upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (L.coef) cudaFree(L.coef);
        cudaFree(L.blob);
        delete p;
        return EXENV_ERR_CUDA;
      }
    }
    p->desc.lut_gx = p->desc.lut_gy = p->desc.lut_maps = nullptr;  // host pointers are not retained
  }
  *out = p;
  return EXENV_OK;
  EX_ABI_CATCH
}

int exenv_plan_destroy(exenv_plan_t plan) {
  if (!plan) return EXENV_OK;
  {
    DeviceGuard guard(plan->device);  // may run from a garbage collector at any time: leave the current device alone
    for (auto &kv : plan->schedules) free_schedule(kv.second);
    if (plan->lut.blob) cudaFree(plan->lut.blob);
    if (plan->lut.coef) cudaFree(plan->lut.coef);
    for (int i = 0; i < 3; ++i) {
      if (plan->host_ws[i]) cudaFree(plan->host_ws[i]);
      if (plan->host_stream[i]) cudaStreamDestroy(plan->host_stream[i]);
    }
  }
  delete plan;
  return EXENV_OK;
}

int exenv_step(exenv_plan_t plan, const ExenvStepArgs *a, void *stream) {
  if (!plan || !a) {
    set_error("exenv_step: NULL plan/args");
    return EXENV_ERR_NULL;
  }
  EX_ON_DEVICE(plan)
  if (a->batch < 0) {
    set_error("exenv_step: negative batch");
    return EXENV_ERR_INVALID;
  }
  if (a->batch == 0) return EXENV_OK;
  const size_t es = plan->desc.dtype == EXENV_F64 ? 8 : 4;
  for (int d = 0; d < plan->dims.n_phys; ++d) {
    if (!a->state_in[d] || !a->state_out[d]) {
      set_error("exenv_step: state leaf %d is NULL", d);
      return EXENV_ERR_NULL;
    }
    if (plan->desc.n_control > 0 && !a->ref[d]) {
      set_error("exenv_step: control_state is set but reference leaf %d is NULL", d);
      return EXENV_ERR_NULL;
    }
    if (!aligned_to(a->state_in[d], es) || !aligned_to(a->state_out[d], es)) {
      set_error("exenv_step: state leaf %d is not aligned to its element size", d);
      return EXENV_ERR_ALIGN;
    }
  }
  if (!a->action || !a->obs) {
    set_error("exenv_step: action/obs is NULL");
    return EXENV_ERR_NULL;
  }
  if (!aligned_to(a->action, es) || !aligned_to(a->obs, es)) {
    set_error("exenv_step: action/obs not aligned to the element size");
    return EXENV_ERR_ALIGN;
  }
  return launch_step(*plan, *a, (cudaStream_t)stream);
}

int exenv_reset_pack(exenv_plan_t plan, const ExenvResetArgs *a, void *stream) {
  if (!plan || !a) {
    set_error("exenv_reset_pack: NULL plan/args");
    return EXENV_ERR_NULL;
  }
  EX_ON_DEVICE(plan)
  if (a->batch < 0 || a->mode < 0 || a->mode > 3) {
    set_error("exenv_reset_pack: bad batch/mode");
    return EXENV_ERR_INVALID;
  }
  if (a->batch == 0) return EXENV_OK;
  for (int d = 0; d < plan->dims.n_phys; ++d)
    if (!a->state_out[d]) {
      set_error("exenv_reset_pack: state leaf %d is NULL", d);
      return EXENV_ERR_NULL;
    }
  if (!a->obs || ((a->mode == EXENV_RESET_FROM_NORM || a->mode == EXENV_RESET_FROM_OBS) && !a->values)) {
    set_error("exenv_reset_pack: obs/values is NULL");
    return EXENV_ERR_NULL;
  }
  return launch_reset(*plan, *a, (cudaStream_t)stream);
}

int exenv_epilogue(exenv_plan_t plan, const ExenvEpilogueArgs *a, void *stream) {
  if (!plan || !a) {
    set_error("exenv_epilogue: NULL plan/args");
    return EXENV_ERR_NULL;
  }
  EX_ON_DEVICE(plan)
  if (a->batch < 0) {
    set_error("exenv_epilogue: negative batch");
    return EXENV_ERR_INVALID;
  }
  if (a->batch == 0) return EXENV_OK;
  if (plan->per_env) {
    set_error("exenv_epilogue supports shared parameters only");
    return EXENV_ERR_UNSUPPORTED;
  }
  for (int d = 0; d < plan->dims.n_phys; ++d) {
    if (!a->state[d] || (plan->desc.n_control > 0 && !a->ref[d])) {
      set_error("exenv_epilogue: state / reference leaf %d is NULL", d);
      return EXENV_ERR_NULL;
    }
  }
  return plan->desc.dtype == EXENV_F64 ? launch_epilogue_f64(*plan, *a, (cudaStream_t)stream)
                                       : launch_epilogue_f32(*plan, *a, (cudaStream_t)stream);
}

int exenv_gym_epilogue(exenv_plan_t plan, const ExenvGymArgs *a, void *stream) {
  if (!plan || !a) {
    set_error("exenv_gym_epilogue: NULL plan/args");
    return EXENV_ERR_NULL;
  }
  EX_ON_DEVICE(plan)
  if (a->batch < 0) {
    set_error("exenv_gym_epilogue: negative batch");
    return EXENV_ERR_INVALID;
  }
  if (a->batch == 0) return EXENV_OK;
  if (plan->per_env) {
    set_error("exenv_gym_epilogue supports shared parameters only");
    return EXENV_ERR_UNSUPPORTED;
  }
  for (int d = 0; d < plan->dims.n_phys; ++d)
    if (!a->state[d] || (plan->desc.n_control > 0 && !a->ref[d])) {
      set_error("exenv_gym_epilogue: state / reference leaf %d is NULL", d);
      return EXENV_ERR_NULL;
    }
  if (a->hold_steps && !a->rng) {
    set_error("exenv_gym_epilogue: hold_steps without rng");
    return EXENV_ERR_NULL;
  }
  if (a->hold_steps && a->hold_steps_max <= a->hold_steps_min) {
    set_error("exenv_gym_epilogue: hold_steps_max must exceed hold_steps_min");
    return EXENV_ERR_INVALID;
  }
  return plan->desc.dtype == EXENV_F64 ? launch_gym_epilogue_f64(*plan, *a, (cudaStream_t)stream)
                                       : launch_gym_epilogue_f32(*plan, *a, (cudaStream_t)stream);
}

int exenv_sim_ahead_shape(exenv_plan_t plan, int64_t n_actions, double obs_stepsize, double action_stepsize,
                          int64_t *n_save, int64_t *n_steps) {
  if (!plan) {
    set_error("exenv_sim_ahead_shape: NULL plan");
    return EXENV_ERR_NULL;
  }
  EX_ABI_TRY
  SimSchedule *s = nullptr;
  const int E = 128;
  int rc = get_schedule(plan, n_actions, obs_stepsize, action_stepsize, auto_tile_steps(*plan, E), false, &s);
  if (rc != EXENV_OK) return rc;
  if (n_save) *n_save = s->n_save;
  if (n_steps) *n_steps = s->n_steps;
  return EXENV_OK;
  EX_ABI_CATCH
}

int exenv_sim_ahead_prepare(exenv_plan_t plan, int64_t n_actions, double obs_stepsize, double action_stepsize,
                            int32_t tile_steps, int64_t *n_save, int64_t *n_steps) {
  if (!plan) {
    set_error("exenv_sim_ahead_prepare: NULL plan");
    return EXENV_ERR_NULL;
  }
  EX_ABI_TRY
  EX_ON_DEVICE(plan)
  SimSchedule *s = nullptr;
  const int K = tile_steps > 0 ? tile_steps : auto_tile_steps(*plan, 128);
  int rc = get_schedule(plan, n_actions, obs_stepsize, action_stepsize, K, true, &s);
  if (rc != EXENV_OK) return rc;
  if (n_save) *n_save = s->n_save;
  if (n_steps) *n_steps = s->n_steps;
  return EXENV_OK;
  EX_ABI_CATCH
}

int exenv_schedule_build(const ExenvDesc *desc, int64_t n_actions, double obs_stepsize, double action_stepsize,
                         int32_t tile_steps, ExenvScheduleInfo *info, uint64_t *words, double *dt, double *theta,
                         int32_t *tiles) {
  if (!desc || !info) {
    set_error("exenv_schedule_build: NULL argument");
    return EXENV_ERR_NULL;
  }
  EX_ABI_TRY
  ExenvPlan tmp;
  tmp.desc = *desc;
  tmp.deadtime = desc->env == EXENV_PMSM ? (int)desc->p[6] : 0;
  SimSchedule s;
  int rc = build_schedule(tmp, n_actions, obs_stepsize, action_stepsize, tile_steps > 0 ? tile_steps : 32, s);
  if (rc != EXENV_OK) return rc;
  info->n_save = s.n_save;
  info->n_steps = s.n_steps;
  info->n_tiles = s.n_tiles;
  info->act_cap = s.act_cap;
  info->save_cap = s.save_cap;
  info->tile_steps = s.tile_steps;
  if (words) memcpy(words, s.words.data(), s.words.size() * sizeof(uint64_t));
  if (dt) memcpy(dt, s.dt.data(), s.dt.size() * sizeof(double));
  if (theta) memcpy(theta, s.theta.data(), s.theta.size() * sizeof(double));
  if (tiles) memcpy(tiles, s.tiles.data(), s.tiles.size() * sizeof(SimTile));
  return EXENV_OK;
  EX_ABI_CATCH
}

int exenv_trunc_thresholds(int32_t dtype, double mn, double mx, double *lo_last, double *hi_first) {
  if (!lo_last || !hi_first) {
    set_error("exenv_trunc_thresholds: NULL argument");
    return EXENV_ERR_NULL;
  }
  bool ok;
  if (dtype == EXENV_F64) {
    TruncThr<double> t;
    ok = make_trunc_thresholds<double>(mn, mx, t);
    *lo_last = t.lo_last;
    *hi_first = t.hi_first;
  } else if (dtype == EXENV_F32) {
    TruncThr<float> t;
    ok = make_trunc_thresholds<float>((float)mn, (float)mx, t);
    *lo_last = t.lo_last;
    *hi_first = t.hi_first;
  } else {
    set_error("exenv_trunc_thresholds: bad dtype %d", dtype);
    return EXENV_ERR_INVALID;
  }
  if (!ok) {
    set_error("exenv_trunc_thresholds: degenerate range [%g, %g]", mn, mx);
    return EXENV_ERR_INVALID;
  }
  return EXENV_OK;
}

static int validate_sim(const ExenvPlan *plan, const ExenvSimArgs *a) {
  if (!plan || !a) {
    set_error("exenv_sim_ahead: NULL plan/args");
    return EXENV_ERR_NULL;
  }
  if (a->batch < 0 || a->n_actions < 1) {
    set_error("exenv_sim_ahead: bad batch / n_actions");
    return EXENV_ERR_INVALID;
  }
  if (a->obs_stepsize > a->action_stepsize) {  // core_env.py:591-593
    set_error("The action stepsize should be greater or equal to the observation stepsize.");
    return EXENV_ERR_INVALID;
  }
  if (plan->desc.env == EXENV_PMSM && a->obs_stepsize != a->action_stepsize) {
    set_error("PMSM sim_ahead requires obs_stepsize == action_stepsize (reference pmsm_env.py:787)");
    return EXENV_ERR_UNSUPPORTED;
  }
  for (int d = 0; d < plan->dims.n_phys; ++d) {
    if (!a->state_in[d]) {
      set_error("exenv_sim_ahead: state leaf %d is NULL", d);
      return EXENV_ERR_NULL;
    }
    if (plan->desc.n_control > 0 && !a->ref[d]) {
      set_error("exenv_sim_ahead: control_state is set but reference leaf %d is NULL", d);
      return EXENV_ERR_NULL;
    }
  }
  if (!a->actions || !a->obs) {
    set_error("exenv_sim_ahead: actions/obs is NULL");
    return EXENV_ERR_NULL;
  }
  if (a->layout != EXENV_LAYOUT_BATCH_MAJOR && a->layout != EXENV_LAYOUT_TIME_MAJOR) {
    set_error("exenv_sim_ahead: unknown layout %d", a->layout);
    return EXENV_ERR_INVALID;
  }
  return EXENV_OK;
}

static int prepare_sim(ExenvPlan *plan, const ExenvSimArgs *a, int *E_out, SimSchedule **sch, cudaStream_t stream) {
  const int E = 128;  // environments per CTA (compile-time constant of the sim kernel)
  int K = a->tile_steps > 0 ? a->tile_steps : auto_tile_steps(*plan, E);
  int rc = get_schedule(plan, a->n_actions, a->obs_stepsize, a->action_stepsize, K, true, sch, &stream);
  if (rc != EXENV_OK) return rc;
  const SimSchedule &s = **sch;
  const int64_t max_steps = a->max_steps;
  if (max_steps > 0 && s.n_steps > max_steps) {  // diffrax: RuntimeError from diffeqsolve(max_steps=4096)
    set_error("The maximum number of solver steps was reached. Try increasing `max_steps`. (%lld > %lld)",
              (long long)s.n_steps, (long long)max_steps);
    return EXENV_ERR_MAX_STEPS;
  }
  const size_t smem = sim_smem_bytes(*plan, s.tile_steps, s.act_cap, s.save_cap);
  if (smem > 227 * 1024) {
    set_error("sim_ahead tile (%d steps x %d envs) needs %zu bytes of shared memory (> 227 KB)", K, E, smem);
    return EXENV_ERR_INVALID;
  }
  *E_out = E;
  return EXENV_OK;
}

int exenv_sim_ahead(exenv_plan_t plan, const ExenvSimArgs *a, void *stream) {
  int rc = validate_sim(plan, a);
  if (rc != EXENV_OK) return rc;
  if (a->batch == 0) return EXENV_OK;
  EX_ABI_TRY
  EX_ON_DEVICE(plan)
  int E = 0;
  SimSchedule *sch = nullptr;
  rc = prepare_sim(plan, a, &E, &sch, (cudaStream_t)stream);
  if (rc != EXENV_OK) return rc;
  if (a->layout == EXENV_LAYOUT_TIME_MAJOR) return launch_simtm(*plan, *a, *sch, (cudaStream_t)stream);
  ExenvSimArgs b = *a;
  b.block_envs = E;
  return launch_sim(*plan, b, *sch, (cudaStream_t)stream);
  EX_ABI_CATCH
}

// Host-buffer variant: what a caller holding CPU arrays (the reference's CPU-JAX user) pays end to end.  The batch is
// cut into chunks; each chunk's inputs go host->device, the persistent kernel runs, outputs come back device->host, on
// NSTREAM streams so copies in both directions overlap the kernel.
int exenv_sim_ahead_host(exenv_plan_t plan, const ExenvSimArgs *h, int64_t chunk_envs) {
  int rc = validate_sim(plan, h);
  if (rc != EXENV_OK) return rc;
  if (h->batch == 0) return EXENV_OK;
  if (h->layout != EXENV_LAYOUT_BATCH_MAJOR) {
    set_error("exenv_sim_ahead_host: host buffers are batch-major (the reference's layout)");
    return EXENV_ERR_UNSUPPORTED;
  }
  EX_ABI_TRY
  EX_ON_DEVICE(plan)
  const size_t es = plan->desc.dtype == EXENV_F64 ? 8 : 4;
  const int D = plan->dims.n_phys, A = plan->dims.n_act, NODE = plan->dims.n_ode;
  const int OW = plan->dims.n_obs + plan->desc.n_control, nt = plan->dims.n_trunc;
  int64_t n_save = 0, n_steps = 0;
  rc = exenv_sim_ahead_prepare(plan, h->n_actions, h->obs_stepsize, h->action_stepsize, h->tile_steps, &n_save, &n_steps);
  if (rc != EXENV_OK) return rc;
  const int64_t T = h->n_actions;
  // bytes per environment on the device
  size_t per_env = es * ((size_t)D + (size_t)T * A + (size_t)n_save * OW);
  for (int d = 0; d < D; ++d) {
    if (h->traj[d]) per_env += es * (size_t)n_save;
    if (h->last[d]) per_env += es;
    if (h->ref[d]) per_env += es;
  }
  for (int d = 0; d < NODE; ++d)
    if (h->k_last[d]) per_env += es;
  if (h->reward) per_env += es * (size_t)(n_save - 1);
  if (h->truncated) per_env += (size_t)n_save * nt;
  if (h->terminated) per_env += (size_t)(n_save - 1);
  constexpr int NSTREAM = 3;
  if (chunk_envs <= 0) {
    const size_t target = (size_t)768 << 20;  // ~0.75 GiB of device workspace per pipeline stage
    chunk_envs = (int64_t)std::max<size_t>(128, (target / per_env) / 128 * 128);
  }
  chunk_envs = std::min<int64_t>(chunk_envs, h->batch);
  // every sub-buffer starts on a 256-byte boundary: at most 8 + 3 * D + NODE of them
  const size_t ws_bytes = per_env * (size_t)chunk_envs + (size_t)256 * (size_t)(12 + 3 * D + NODE);
  // The three streams and their workspaces live in the plan: created on the first call, grown when a later call needs
  // more, released by exenv_plan_destroy.  One host-buffer call at a time per plan.
  std::lock_guard<std::mutex> pipe_lock(plan->host_mu);
  for (int i = 0; i < NSTREAM; ++i)
    if (!plan->host_stream[i]) EX_CUDA(cudaStreamCreateWithFlags(&plan->host_stream[i], cudaStreamNonBlocking));
  if (plan->host_ws_bytes < ws_bytes) {
    for (int i = 0; i < NSTREAM; ++i) {
      if (plan->host_ws[i]) cudaFree(plan->host_ws[i]);
      plan->host_ws[i] = nullptr;
    }
    plan->host_ws_bytes = 0;
    for (int i = 0; i < NSTREAM; ++i) EX_CUDA(cudaMalloc(&plan->host_ws[i], ws_bytes));
    plan->host_ws_bytes = ws_bytes;
  }
  cudaStream_t *st = plan->host_stream;
  int result = EXENV_OK;
  cudaError_t copy_err = cudaSuccess;
  auto copy = [&](void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s) {
    const cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, s);
    if (e != cudaSuccess && copy_err == cudaSuccess) copy_err = e;
  };
  int slot = 0;
  for (int64_t b0 = 0; b0 < h->batch && result == EXENV_OK && copy_err == cudaSuccess;
       b0 += chunk_envs, slot = (slot + 1) % NSTREAM) {
    const int64_t nb = std::min<int64_t>(chunk_envs, h->batch - b0);
    cudaStream_t s = st[slot];
    char *cur = (char *)plan->host_ws[slot];
    auto take = [&](size_t bytes) {
      char *p = cur;
      cur += (bytes + 255) & ~(size_t)255;
      return (void *)p;
    };
    ExenvSimArgs d = *h;
    d.batch = nb;
    struct Back {
      void *host;
      void *dev;
      size_t bytes;
    };
    std::vector<Back> back;
    auto in = [&](const void *host, size_t per, size_t off_bytes) -> const void * {
      void *dv = take(per * nb);
      copy(dv, (const char *)host + off_bytes, per * nb, cudaMemcpyHostToDevice, s);
      return dv;
    };
    auto outb = [&](void *host, size_t per) -> void * {
      void *dv = take(per * nb);
      back.push_back({(char *)host + per * b0, dv, per * (size_t)nb});
      return dv;
    };
    for (int i = 0; i < D; ++i) {
      d.state_in[i] = in(h->state_in[i], es, es * b0);
      if (h->ref[i]) d.ref[i] = in(h->ref[i], es, es * b0);
      if (h->traj[i]) d.traj[i] = outb(h->traj[i], es * (size_t)n_save);
      if (h->last[i]) d.last[i] = outb(h->last[i], es);
    }
    for (int i = 0; i < NODE; ++i)
      if (h->k_last[i]) d.k_last[i] = outb(h->k_last[i], es);
    d.actions = in(h->actions, es * (size_t)T * A, es * (size_t)T * A * b0);
    d.obs = outb(h->obs, es * (size_t)n_save * OW);
    if (h->reward) d.reward = outb(h->reward, es * (size_t)(n_save - 1));
    if (h->truncated) d.truncated = (uint8_t *)outb(h->truncated, (size_t)n_save * nt);
    if (h->terminated) d.terminated = (uint8_t *)outb(h->terminated, (size_t)(n_save - 1));
    result = exenv_sim_ahead(plan, &d, s);
    if (result != EXENV_OK) break;
    for (const Back &bk : back) copy(bk.host, bk.dev, bk.bytes, cudaMemcpyDeviceToHost, s);
  }
  // one exit path: the streams are drained whatever happened above (the workspaces stay with the plan)
  for (int i = 0; i < NSTREAM; ++i) {
    const cudaError_t e = cudaStreamSynchronize(st[i]);
    if (e != cudaSuccess && copy_err == cudaSuccess) copy_err = e;
  }
  if (result == EXENV_OK && copy_err != cudaSuccess) {
    set_error("exenv_sim_ahead_host: %s", cudaGetErrorString(copy_err));
    result = EXENV_ERR_CUDA;
  }
  return result;
  EX_ABI_CATCH
}

}  // extern "C"
