// exenv_simtm.inl (included by exenv_simtm_f32.cu / exenv_simtm_f64.cu with EXENV_REAL / EXENV_SFX) -- vmap_sim_ahead
// for the TIME-MAJOR (structure-of-arrays) trajectory layout, EXENV_LAYOUT_TIME_MAJOR of include/exenv_b200.h.
//
// Physical layout: actions (T, B, A) in; observations (N+1, B, O'), trajectory leaves (N+1, B), reward / terminated
// (N, B), truncated (N+1, B, n_trunc) out.  The host mirror hands them out as permuted views with the reference's
// logical shapes ((B, T, A), (B, N+1, O'), (B, N+1), ... -- core_env.py:571-616), so values and indexing are those
// of the reference; only the strides differ.
//
// With the batch axis innermost nothing has to be transposed on chip: thread e of a CTA owns environment
// blockIdx.x * blockDim.x + e for the whole trajectory (ODE state, FSAL derivative, stage increments and a 4-entry
// window of prepared zero-order-hold actions in registers) and every access of a warp is one contiguous segment:
//   action of ZOH index q        32 lanes x A scalars  = 128 B * A  (loaded one solver step ahead of its first use)
//   trajectory leaf, one save    32 lanes x 1 scalar   = 128 B
//   observation row, one save    32 lanes x O' scalars = 128 B * O' (one 8/16-byte vector store per lane if O' = O)
// No shared memory is used except for the PMSM flux maps (staged once per CTA by a TMA bulk copy), so occupancy is
// bound by registers only and the kernel streams S*(A + O + D) bytes per env-step at whatever rate the solver
// arithmetic allows.  The per-step time bookkeeping (ZOH indices per stage, step lengths, dense-output records) is
// the same host-built schedule the batch-major kernel uses (exenv_api.cu: build_schedule_t), read through L1 one
// step ahead; the per-environment arithmetic is the same exenv_sim_core.cuh code, so both layouts produce identical
// values.
#include <cstdio>
#include <cstring>

#include "exenv_kernel_common.cuh"
#include "exenv_sim_core.cuh"

#define EXENV_CAT_(a, b) a##b
#define EXENV_CAT(a, b) EXENV_CAT_(a, b)

namespace exenv {

constexpr int SIMTM_THREADS = 128;

template <typename T>
struct SimTmK {
  long B, n_actions, n_save, n_steps;
  const T *sin[EXENV_MAX_PHYS];
  const T *actions;  // (T, B, A)
  T *obs;            // (N+1, B, O')
  T *traj[EXENV_MAX_PHYS];
  T *last[EXENV_MAX_PHYS];
  T *klast[EXENV_MAX_ODE];
  const T *ref[EXENV_MAX_PHYS];
  T *reward;       // (N, B)
  uint8_t *trunc;  // (N+1, B, n_trunc)
  uint8_t *term;   // (N, B)
  const StepRec<T> *steps;  // packed (schedule word, step length) per solver step + one look-ahead entry
  const SaveRec<T> *recs;
  int n_control;
  int control[EXENV_MAX_CONTROL];
  int deadtime;
  T pmsm_adv, pmsm_lin_stop;
  unsigned lut_bytes, lut_off_gy, lut_off_cells;
  const void *lut_blob;
  int lut_nx, lut_ny;
  T lut_inv_dx, lut_inv_dy;
};

__device__ __forceinline__ unsigned tm_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// Streaming (evict-first) vector access of BYTES = 4 / 8 / 16 bytes.
template <int BYTES>
__device__ __forceinline__ void stcs_chunk(void *dst, const void *src) {
  if constexpr (BYTES == 16) {
    float4 t;
    memcpy(&t, src, 16);
    __stcs(reinterpret_cast<float4 *>(dst), t);
  } else if constexpr (BYTES == 8) {
    float2 t;
    memcpy(&t, src, 8);
    __stcs(reinterpret_cast<float2 *>(dst), t);
  } else {
    float t;
    memcpy(&t, src, 4);
    __stcs(reinterpret_cast<float *>(dst), t);
  }
}
template <int BYTES>
__device__ __forceinline__ void ldcs_chunk(void *dst, const void *src) {
  if constexpr (BYTES == 16) {
    const float4 t = __ldcs(reinterpret_cast<const float4 *>(src));
    memcpy(dst, &t, 16);
  } else if constexpr (BYTES == 8) {
    const float2 t = __ldcs(reinterpret_cast<const float2 *>(src));
    memcpy(dst, &t, 8);
  } else {
    const float t = __ldcs(reinterpret_cast<const float *>(src));
    memcpy(dst, &t, 4);
  }
}

// p + n elements with one IMAD.WIDE.U32 (running row pointers: the batch-axis stride fits 32 bits).
template <typename P>
__device__ __forceinline__ P *ptr_add(P *p, unsigned n) {
  unsigned long long r;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(n), "r"((unsigned)sizeof(P)), "l"((unsigned long long)p));
  return reinterpret_cast<P *>(r);
}

// O consecutive scalars with the widest vector stores (16 bytes at most).
template <typename T, int O>
__device__ __forceinline__ void store_row_vec(T *dst, const T *v) {
  constexpr int ROW = O * (int)sizeof(T);
  constexpr int CH = ROW < 16 ? ROW : 16;
#pragma unroll
  for (int i = 0; i < ROW / CH; ++i)
    stcs_chunk<CH>(reinterpret_cast<char *>(dst) + i * CH, reinterpret_cast<const char *>(v) + i * CH);
}

// The window of prepared actions as the `fetch` object of exenv_sim_core.cuh.  Slot 0 holds the step's base index
// (the kernel slides the window before every step), so a stage's action is selected by its 2-bit offset alone.
template <typename T, int A, int W>
struct TmFetch {
  T uw[W][A];
  int wbase;
  __device__ __forceinline__ void pick(int pos, T *u) const {
#pragma unroll
    for (int c = 0; c < A; ++c) {
      if constexpr (W == 2) {
        u[c] = (pos & 1) ? uw[1][c] : uw[0][c];
      } else {
        const T lo2 = (pos & 1) ? uw[1][c] : uw[0][c];
        const T hi2 = (pos & 1) ? uw[3][c] : uw[2][c];
        u[c] = (pos & 2) ? hi2 : lo2;
      }
    }
  }
  __device__ __forceinline__ void stage(uint64_t w, int s, T *u) const { pick((int)(unsigned)(w >> (32 + 2 * s)), u); }
  __device__ __forceinline__ void operator()(int idx, T *u) const { pick(idx - wbase, u); }
};

// GENERIC: reference entries in the observation (control_state != []), the fused reward / truncated / terminated
// outputs (vmap_generate_rew_trunc_term_ahead, core_env.py:618-647) or per-environment parameters.
template <class Env, int SOLVER, typename T, bool PER_ENV, bool GENERIC, int W>
__global__ void __launch_bounds__(SIMTM_THREADS)
    simtm_kernel(const __grid_constant__ SimTmK<T> a, const __grid_constant__ KParams<T> kp) {
  constexpr int D = Env::D, A = Env::A, NODE = Env::NODE, O = Env::O, NP = num_params<Env>();
  constexpr int NST = staged_values<Env>();
  extern __shared__ __align__(16) unsigned char smem_raw[];

  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = b < a.B;

  typename Env::Aux aux{};
  if constexpr (Env::IS_PMSM && Env::SATURATED) {
    // stage the flux maps once per CTA: TMA bulk copy + mbarrier (SASS: UBLKCP)
    unsigned char *sm_lut = smem_raw;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(sm_lut + a.lut_bytes);
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tm_smem_u32(mbar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tm_smem_u32(mbar)), "r"(a.lut_bytes)
                   : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       tm_smem_u32(sm_lut)),
                   "l"(a.lut_blob), "r"(a.lut_bytes), "r"(tm_smem_u32(mbar))
                   : "memory");
    }
    __syncthreads();
    unsigned done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(tm_smem_u32(mbar))
          : "memory");
    }
    aux.lut.gx = reinterpret_cast<const T *>(sm_lut);
    aux.lut.gy = reinterpret_cast<const T *>(sm_lut + a.lut_off_gy);
    aux.lut.cells = reinterpret_cast<const T *>(sm_lut + a.lut_off_cells);
    aux.lut.nx = a.lut_nx;
    aux.lut.ny = a.lut_ny;
    aux.lut.inv_dx = a.lut_inv_dx;
    aux.lut.inv_dy = a.lut_inv_dy;
  }
  if (!valid) return;  // no block-wide synchronisation below

  // ---- per-environment constants and initial state ------------------------------------------------------------
  T pl[NP];
  const T *p = kp.v;
  if constexpr (PER_ENV) {
    load_env_params<Env, T, true>(kp, b, pl);
    p = pl;
  }
  const typename Env::Pre pre = Env::precompute(p);
  const T *pn = p + Env::NSTATIC;
  T s0[D];
#pragma unroll
  for (int d = 0; d < D; ++d) s0[d] = __ldg(a.sin[d] + b);
  T y[NODE], f[NODE];
  if constexpr (Env::IS_PMSM) {
    y[0] = s0[3];
    y[1] = s0[4];
    y[2] = s0[2];
    aux.omega_el = s0[6];
    aux.deadtime = a.deadtime;
  } else {
#pragma unroll
    for (int i = 0; i < NODE; ++i) y[i] = s0[i];
  }
#pragma unroll
  for (int i = 0; i < NODE; ++i) f[i] = (T)0;
  T last[NST];
#pragma unroll
  for (int i = 0; i < NST; ++i) last[i] = (T)0;

  // observation normalisation: prepared reciprocal of every span (correctly rounded quotient, exenv_math.cuh)
  InvDiv<T> sdiv[D];
#pragma unroll
  for (int d = 0; d < D; ++d) sdiv[d] = make_invdiv<T>(pn[2 * d + 1] - pn[2 * d]);
  T om_norm = (T)0;
  if constexpr (Env::IS_PMSM) om_norm = norm_inv<true, T>(s0[6], pn[12], sdiv[6]);
  T rf[D];
  if constexpr (GENERIC) {
#pragma unroll
    for (int d = 0; d < D; ++d) rf[d] = (a.n_control > 0) ? __ldg(a.ref[d] + b) : (T)0;
  }
  const int OW = GENERIC ? O + a.n_control : O;
  const int nt = trunc_dim<Env>(GENERIC ? a.n_control : 0);

  // ---- window of prepared ZOH actions: uw[k] = physical action of staged index wbase + k, k < W -------------------
  // Slot 0 follows the step's smallest stage index (`base`); the window slides by whole entries, the new last entry
  // is prepared from a raw action whose load was issued one slide earlier (>= one solver step of latency hiding).
  // Every staged index is loaded and prepared exactly once.
  const int dead = Env::IS_PMSM ? a.deadtime : 0;
  const int q_max = (int)a.n_actions - 1 + dead;  // last staged index that exists
  TmFetch<T, A, W> win;
  T pend[A];
  win.wbase = 0;
  // staged indices are loaded strictly in order: ap points at the raw action of the next one (staged q -> raw action
  // q - dead; the PMSM buffer entry q < dead carries no load; past the end the last action is re-read, never used)
  const T *ap = a.actions + b * A;
  int qn = 0;
  const unsigned act_stride = (unsigned)(a.B * A);
  auto load_next = [&](T *raw) {
    if (!Env::IS_PMSM || qn >= dead) {
      ldcs_chunk<A * (int)sizeof(T)>(raw, ap);
      if (qn < q_max) ap = ptr_add(ap, act_stride);
    } else {
#pragma unroll
      for (int c = 0; c < A; ++c) raw[c] = (T)0;
    }
    ++qn;
  };
  auto prepare = [&](int q, const T *raw, T *u) {
    q = q < q_max ? q : q_max;
    prepare_action<Env, T>((long)q, raw, s0, p, a.deadtime, a.n_actions, a.pmsm_adv, a.pmsm_lin_stop, u);
  };
#pragma unroll
  for (int k = 0; k < W; ++k) {
    T raw[A];
    load_next(raw);
    prepare(k, raw, win.uw[k]);
  }
  load_next(pend);
  auto slide_to = [&](int base) {
#pragma unroll 1
    while (win.wbase < base) {
#pragma unroll
      for (int k = 0; k + 1 < W; ++k)
#pragma unroll
        for (int c = 0; c < A; ++c) win.uw[k][c] = win.uw[k + 1][c];
      ++win.wbase;
      prepare(win.wbase + W - 1, pend, win.uw[W - 1]);
      load_next(pend);
    }
  };

  // ---- save sink: saves arrive in order (0, 1, 2, ...), so every output keeps a running per-thread pointer ---------
  const unsigned Bu = (unsigned)a.B;
  T *tp[D];
#pragma unroll
  for (int d = 0; d < D; ++d) tp[d] = a.traj[d] ? a.traj[d] + b : nullptr;
  T *op = a.obs + b * OW;
  T *rwp = (GENERIC && a.reward) ? a.reward + b : nullptr;
  uint8_t *trp = (GENERIC && a.trunc) ? a.trunc + b * nt : nullptr;
  uint8_t *tmp = (GENERIC && a.term) ? a.term + b : nullptr;
  auto sink = [&](long sidx, const T *st) {
#pragma unroll
    for (int i = 0; i < NST; ++i) last[i] = st[i];
    T sv[D];
    if constexpr (Env::IS_PMSM) {
      sv[0] = st[4];
      sv[1] = st[5];
      sv[2] = st[0];
      sv[3] = st[1];
      sv[4] = st[2];
      sv[5] = st[3];
      sv[6] = s0[6];
    } else {
#pragma unroll
      for (int d = 0; d < D; ++d) sv[d] = st[d];
    }
#pragma unroll
    for (int d = 0; d < D; ++d) {
      if (!GENERIC || tp[d]) {
        __stcs(tp[d], sv[d]);
        tp[d] = ptr_add(tp[d], Bu);
      }
    }
    if constexpr (!GENERIC) {
      T ob[O];
      if constexpr (Env::IS_PMSM) {
        ob[0] = norm_inv<true, T>(sv[3], pn[6], sdiv[3]);
        ob[1] = norm_inv<true, T>(sv[4], pn[8], sdiv[4]);
        ob[2] = om_norm;
        ob[3] = norm_inv<true, T>(sv[5], pn[10], sdiv[5]);
        ob[4] = st[6];
        ob[5] = st[7];
        ob[6] = norm_inv<true, T>(sv[0], pn[0], sdiv[0]);
        ob[7] = norm_inv<true, T>(sv[1], pn[2], sdiv[1]);
      } else {
#pragma unroll
        for (int o = 0; o < O; ++o) ob[o] = norm_inv<true, T>(sv[o], pn[2 * o], sdiv[o]);
      }
      store_row_vec<T, O>(op, ob);
      op = ptr_add(op, Bu * (unsigned)O);
    } else {
      T ob[O + EXENV_MAX_CONTROL];
      observe<Env, T>(sv, p, ob);
#pragma unroll
      for (int j = 0; j < EXENV_MAX_CONTROL; ++j)
        if (j < a.n_control) ob[O + j] = norm_field<Env, T>(a.control[j], rf, p);
#pragma unroll
      for (int o = 0; o < O + EXENV_MAX_CONTROL; ++o)
        if (o < OW) __stcs(op + o, ob[o]);
      op += (size_t)Bu * (unsigned)OW;
      if (rwp || trp || tmp) {
        bool t0 = false;
#pragma unroll
        for (int i = 0; i < O + EXENV_MAX_CONTROL; ++i)
          if (i < nt) {
            const bool tv = trunc_entry<Env, T>(i, sv, ob, p);
            if (i == 0) t0 = tv;
            if (trp) trp[i] = tv;
          }
        if (trp) trp += (size_t)Bu * (unsigned)nt;
        if (sidx >= 1) {
          const T rw = a.n_control > 0 ? reward_of<Env, T>(sv, rf, p, a.n_control, a.control) : (T)0;
          if (rwp) {
            __stcs(rwp, rw);
            rwp += Bu;
          }
          if (tmp) {
            *tmp = terminated_of<Env, T>(rw, t0);
            tmp += Bu;
          }
        }
      }
    }
  };

  // ---- the solver loop: the packed (word, dt) entry of step i+1 is fetched while step i is integrated; the dense-
  // output record of a step is only needed after its last stage, so its load is issued at the top of the same step
  long save_idx = 0;
  const StepRec<T> *sp = a.steps;
  StepRec<T> s_next = *sp;
  slide_to(sched_base(s_next.w));
  if constexpr (SOLVER == EXENV_TSIT5) sim_first_derivative<Env, T>(s_next.w, y, f, p, pre, aux, win);
  const int n_steps = (int)a.n_steps;
#pragma unroll 1
  for (int i = 0; i < n_steps; ++i) {
    const uint64_t w = s_next.w;
    const T dt = s_next.dt;
    const SaveRec<T> rec = a.recs[save_idx];
    s_next = *++sp;  // one look-ahead entry is appended by the host
    slide_to(sched_base(w));
    sim_one_step<Env, SOLVER, T>(w, dt, y, f, p, pre, aux, rec, a.recs, save_idx, win, sink);
  }
  // ---- last_state (core_env.py:484-488) and the fresh solver.init derivative at (t1, y_last) --------------------
  T sl[D];
  if constexpr (Env::IS_PMSM) {
    sl[0] = last[4];
    sl[1] = last[5];
    sl[2] = last[0];
    sl[3] = last[1];
    sl[4] = last[2];
    sl[5] = last[3];
    sl[6] = s0[6];
  } else {
#pragma unroll
    for (int i = 0; i < NODE; ++i) sl[i] = last[i];
  }
#pragma unroll
  for (int d = 0; d < D; ++d)
    if (a.last[d]) a.last[d][b] = sl[d];
  if (SOLVER == EXENV_TSIT5 && a.klast[0]) {
    // solver.init(term, t1, t1 + tau, y_last, args): derivative with the ZOH closure at t1 -> clamped last action
    const int q = (int)(a.n_actions - 1);
    if (q > win.wbase + W - 1) slide_to(q - (W - 1));
    T u[A], yl[NODE], fi[NODE];
    win(q, u);
    if constexpr (Env::IS_PMSM) {
      yl[0] = last[1];
      yl[1] = last[2];
      yl[2] = last[0];
    } else {
#pragma unroll
      for (int i = 0; i < NODE; ++i) yl[i] = last[i];
    }
    Env::template rhs<true>(yl, u, p, pre, aux, fi);
#pragma unroll
    for (int i = 0; i < NODE; ++i) a.klast[i][b] = fi[i];
  }
}

// ---- launcher ----------------------------------------------------------------------------------------------------
template <class Env, int SOLVER, typename T, bool PER_ENV, bool GENERIC, int W>
static int launch_simtm_w(const ExenvPlan &plan, const ExenvSimArgs &x, const SimSchedule &sch, cudaStream_t stream) {
  SimTmK<T> a{};
  a.B = x.batch;
  a.n_actions = x.n_actions;
  a.n_save = sch.n_save;
  a.n_steps = sch.n_steps;
  for (int d = 0; d < Env::D; ++d) {
    a.sin[d] = (const T *)x.state_in[d];
    a.traj[d] = (T *)x.traj[d];
    a.last[d] = (T *)x.last[d];
    a.ref[d] = (const T *)x.ref[d];
  }
  for (int d = 0; d < Env::NODE; ++d) a.klast[d] = (T *)x.k_last[d];
  a.actions = (const T *)x.actions;
  a.obs = (T *)x.obs;
  a.reward = (T *)x.reward;
  a.trunc = x.truncated;
  a.term = x.terminated;
  a.steps = (const StepRec<T> *)sch.d_steprec;
  a.recs = (const SaveRec<T> *)sch.d_theta;
  a.n_control = plan.desc.n_control;
  for (int j = 0; j < EXENV_MAX_CONTROL; ++j) a.control[j] = plan.desc.control[j];
  a.deadtime = plan.deadtime;
  a.pmsm_adv = (T)plan.pmsm_adv;
  a.pmsm_lin_stop = (T)(plan.desc.tau * (double)(x.n_actions - 1));
  a.lut_bytes = (unsigned)plan.lut.bytes;
  a.lut_off_gy = (unsigned)plan.lut.off_gy;
  a.lut_off_cells = (unsigned)plan.lut.off_cells;
  a.lut_blob = plan.lut.blob;
  a.lut_nx = plan.lut.nx;
  a.lut_ny = plan.lut.ny;
  a.lut_inv_dx = (T)plan.lut.inv_dx;
  a.lut_inv_dy = (T)plan.lut.inv_dy;
  if (Env::A == 2 && reinterpret_cast<uintptr_t>(x.actions) % (2 * sizeof(T)) != 0) {
    set_error("exenv_sim_ahead (time-major): actions must be aligned to %zu bytes", 2 * sizeof(T));
    return EXENV_ERR_ALIGN;
  }
  KParams<T> kp = make_kparams<T>(plan);
  size_t smem = 0;
  if (Env::IS_PMSM && (plan.desc.flags & EXENV_FLAG_SATURATED)) smem = plan.lut.bytes + 16;
  auto kern = simtm_kernel<Env, SOLVER, T, PER_ENV, GENERIC, W>;
  if (smem > 48 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("sim_ahead (time-major) needs %zu bytes of shared memory per CTA: %s", smem, cudaGetErrorString(e));
      return EXENV_ERR_CUDA;
    }
  }
  const unsigned grid = (unsigned)((x.batch + SIMTM_THREADS - 1) / SIMTM_THREADS);
  kern<<<grid, SIMTM_THREADS, smem, stream>>>(a, kp);
  g_launch_count++;
  const cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("exenv_sim_ahead (time-major) launch failed: %s", cudaGetErrorString(e));
    return EXENV_ERR_CUDA;
  }
  return EXENV_OK;
}

// Window length: 2 covers every schedule whose steps read at most two consecutive ZOH indices (always the case up to
// floating-point lag of the accumulated solver time); 4 is the general case (build_schedule_t rejects more).
template <class Env, int SOLVER, typename T, bool PER_ENV, bool GENERIC>
static int launch_simtm_t(const ExenvPlan &plan, const ExenvSimArgs &x, const SimSchedule &sch, cudaStream_t stream) {
  if (sch.win_need <= 2) return launch_simtm_w<Env, SOLVER, T, PER_ENV, GENERIC, 2>(plan, x, sch, stream);
  if (sch.win_need <= 4) return launch_simtm_w<Env, SOLVER, T, PER_ENV, GENERIC, 4>(plan, x, sch, stream);
  set_error("sim_ahead (time-major): a solver step needs %d consecutive actions (> 4)", sch.win_need);
  return EXENV_ERR_UNSUPPORTED;
}

template <class Env, typename T>
static int simtm_dispatch2(const ExenvPlan &plan, const ExenvSimArgs &x, const SimSchedule &s, cudaStream_t st) {
  const bool tsit = plan.desc.solver == EXENV_TSIT5;
  // anything beyond the plain call takes the GENERIC instantiation: reference entries, the fused epilogue, skipped
  // (NULL) trajectory leaves, observation rows that cannot be written with aligned vector stores
  bool generic = plan.desc.n_control > 0 || x.reward || x.truncated || x.terminated;
  for (int d = 0; d < Env::D; ++d) generic |= x.traj[d] == nullptr;
  {
    const size_t row_bytes = (size_t)Env::O * sizeof(T);
    const size_t vec = row_bytes < 16 ? row_bytes : 16;
    generic |= reinterpret_cast<uintptr_t>(x.obs) % vec != 0;
  }
  if constexpr (Env::IS_PMSM) {
    if (generic)
      return tsit ? launch_simtm_t<Env, EXENV_TSIT5, T, false, true>(plan, x, s, st)
                  : launch_simtm_t<Env, EXENV_EULER, T, false, true>(plan, x, s, st);
    return tsit ? launch_simtm_t<Env, EXENV_TSIT5, T, false, false>(plan, x, s, st)
                : launch_simtm_t<Env, EXENV_EULER, T, false, false>(plan, x, s, st);
  } else {
    if (plan.per_env)  // per-environment parameters always take the generic instantiation
      return tsit ? launch_simtm_t<Env, EXENV_TSIT5, T, true, true>(plan, x, s, st)
                  : launch_simtm_t<Env, EXENV_EULER, T, true, true>(plan, x, s, st);
    if (generic)
      return tsit ? launch_simtm_t<Env, EXENV_TSIT5, T, false, true>(plan, x, s, st)
                  : launch_simtm_t<Env, EXENV_EULER, T, false, true>(plan, x, s, st);
    return tsit ? launch_simtm_t<Env, EXENV_TSIT5, T, false, false>(plan, x, s, st)
                : launch_simtm_t<Env, EXENV_EULER, T, false, false>(plan, x, s, st);
  }
}

template <typename T>
static int simtm_dispatch(const ExenvPlan &plan, const ExenvSimArgs &x, const SimSchedule &s, cudaStream_t st) {
  switch (plan.desc.env) {
    case EXENV_PENDULUM: return simtm_dispatch2<Pendulum<T>, T>(plan, x, s, st);
    case EXENV_MASS_SPRING_DAMPER: return simtm_dispatch2<MassSpringDamper<T>, T>(plan, x, s, st);
    case EXENV_FLUID_TANK: return simtm_dispatch2<FluidTank<T>, T>(plan, x, s, st);
    case EXENV_CART_POLE: return simtm_dispatch2<CartPole<T>, T>(plan, x, s, st);
    case EXENV_ACROBOT: return simtm_dispatch2<Acrobot<T>, T>(plan, x, s, st);
    case EXENV_PMSM:
      if (plan.desc.flags & EXENV_FLAG_SATURATED) return simtm_dispatch2<Pmsm<T, true>, T>(plan, x, s, st);
      return simtm_dispatch2<Pmsm<T, false>, T>(plan, x, s, st);
  }
  set_error("unknown env %d", plan.desc.env);
  return EXENV_ERR_INVALID;
}

int EXENV_CAT(launch_simtm_, EXENV_SFX)(const ExenvPlan &plan, const ExenvSimArgs &a, const SimSchedule &sch,
                                         cudaStream_t stream) {
  return simtm_dispatch<EXENV_REAL>(plan, a, sch, stream);
}

}  // namespace exenv
