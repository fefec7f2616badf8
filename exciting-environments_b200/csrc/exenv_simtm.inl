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
#include "exenv_simtm_common.cuh"

#define EXENV_CAT_(a, b) a##b
#define EXENV_CAT(a, b) EXENV_CAT_(a, b)

namespace exenv {

// Multiprocessors of a device (grid size of the persistent flavour), queried once per device.
static int device_sm_count(int device) {
  static std::mutex mu;
  static int cache[64] = {};
  std::lock_guard<std::mutex> lk(mu);
  int &n = cache[device & 63];
  if (n <= 0 && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) n = 148;
  return n > 0 ? n : 148;
}

// EXENV_PMSM_COEF=0 keeps the node-form flux maps in the saturated-PMSM fp32 trajectory kernels (A/B measurements).
static bool simtm_coef_enabled() {
  static const bool on = [] {
    const char *e = getenv("EXENV_PMSM_COEF");
    return !(e && e[0] == '0');
  }();
  return on;
}
// Whether launch_simtm_pmsm_coef_f32 can run this plan: coefficient table built (fp32, exact equidistant grids) and
// table + mbarrier + schedule buffers of the persistent CTA within the 227 KB of one CTA.
static bool simtm_coef_fits(const ExenvPlan &plan) {
  return plan.lut.coef && simtm_coef_enabled() &&
         ((plan.lut.coef_bytes + 16 + 15) & ~(size_t)15) + tm_sched_smem_bytes<float>(SIMTM_COEF_THREADS) <= 227 * 1024;
}

template <class Env, int SOLVER, typename T, bool PER_ENV, bool GENERIC, int W, int V, bool EPI>
__global__ void __launch_bounds__(simtm_cta_threads<Env>())
    simtm_kernel(const __grid_constant__ SimTmK<T> a, const __grid_constant__ KParams<T> kp,
                 const __grid_constant__ SimTmHostPre<Env, T> hp) {
  constexpr int D = Env::D, A = Env::A, NODE = Env::NODE, O = Env::O, NP = num_params<Env>();
  constexpr int NST = staged_values<Env>();
  static_assert(!GENERIC || V == 1, "the generic instantiation handles one environment per thread");
  static_assert(!(GENERIC && EPI), "EPI is the fused-epilogue flavour of the fast instantiation");
  extern __shared__ __align__(16) unsigned char smem_raw[];

  // First environment of this thread.  A warp entirely beyond the batch leaves; in the warp that straddles its end the
  // surplus lanes shadow the last group of V environments (same loads, same arithmetic, same values stored to the same
  // addresses): every warp that runs the loop runs it with all 32 lanes, which the shared-memory schedule staging needs.
  // A job = the 32 * V consecutive environments of one warp.  One job per warp, CTA-major -- or, for the persistent
  // coefficient-form PMSM flavour, a strided sequence of jobs per warp (exenv_simtm_common.cuh: simtm_is_coef).
  constexpr bool PERSIST = simtm_is_coef<Env>();
  const unsigned lane = threadIdx.x & 31u, warp_in_cta = threadIdx.x >> 5;
  long job = PERSIST ? (long)warp_in_cta * gridDim.x + blockIdx.x : (long)blockIdx.x * (blockDim.x >> 5) + warp_in_cta;

  typename Env::Aux aux[V];
  // dynamic shared memory: [PMSM flux maps + mbarrier (saturated only) | per-warp schedule buffers]
  const unsigned sched_off = (Env::IS_PMSM && Env::SATURATED) ? ((a.lut.bytes + 16 + 15) & ~15u) : 0u;
  if constexpr (Env::IS_PMSM && Env::SATURATED) {
    // stage the flux maps once per CTA: TMA bulk copy + mbarrier (SASS: UBLKCP)
    unsigned char *sm_lut = smem_raw;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(sm_lut + a.lut.bytes);
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tm_smem_u32(mbar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tm_smem_u32(mbar)), "r"(a.lut.bytes)
                   : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       tm_smem_u32(sm_lut)),
                   "l"(a.lut_blob), "r"(a.lut.bytes), "r"(tm_smem_u32(mbar))
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
#pragma unroll
    for (int v = 0; v < V; ++v) {
      if constexpr (PERSIST) aux[v].lut = lut_view_coef<T>(sm_lut, a.lut);
      else aux[v].lut = lut_view_at<T>(sm_lut, a.lut);
    }
  }
  // no block-wide synchronisation below
  for (;; job += (long)(blockDim.x >> 5) * gridDim.x) {
  if (job * (32 * V) >= a.B) return;
  const long b_own = (job * 32 + lane) * V;
  const long b = b_own < a.B ? b_own : a.B - V;

  // ---- per-environment constants and initial state ------------------------------------------------------------
  T pl[NP];
  const T *p = kp.v;
  if constexpr (PER_ENV) {
    load_env_params<Env, T, true>(kp, b, pl);
    p = pl;
  }
  constexpr bool HOST_PRE = !PER_ENV && simtm_use_host_pre<Env, SOLVER>();
  typename Env::Pre pre_thread;
  if constexpr (!HOST_PRE) pre_thread = Env::precompute(p);
  const typename Env::Pre &pre = HOST_PRE ? hp.pre : pre_thread;
  const T *pn = p + Env::NSTATIC;
  T s0[V][D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    T tmp[V];
    load_vec<T, V>(tmp, a.sin[d] + b);
#pragma unroll
    for (int v = 0; v < V; ++v) s0[v][d] = tmp[v];
  }
  T y[V][NODE], f[V][NODE], last[V][NST];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    if constexpr (Env::IS_PMSM) {
      y[v][0] = s0[v][3];
      y[v][1] = s0[v][4];
      y[v][2] = s0[v][2];
      aux[v].omega_el = s0[v][6];
      aux[v].deadtime = a.deadtime;
    } else {
#pragma unroll
      for (int i = 0; i < NODE; ++i) y[v][i] = s0[v][i];
    }
#pragma unroll
    for (int i = 0; i < NODE; ++i) f[v][i] = (T)0;
#pragma unroll
    for (int i = 0; i < NST; ++i) last[v][i] = (T)0;
  }

  // observation normalisation: one FMA per entry with prepared coefficients (norm_fma, exenv_math.cuh); every
  // instantiation of this kernel (vector, fused-epilogue, generic) and the packed kernel use the same form
  NormFma<T> nf_thread[D];
  if constexpr (!HOST_PRE) {
#pragma unroll
    for (int d = 0; d < D; ++d) nf_thread[d] = make_normfma<T>(pn[2 * d], pn[2 * d + 1]);
  }
  const NormFma<T> *nf = HOST_PRE ? hp.nf : nf_thread;
  T om_norm[V];
#pragma unroll
  for (int v = 0; v < V; ++v) om_norm[v] = Env::IS_PMSM ? norm_fma<T>(s0[v][D - 1], nf[D - 1]) : (T)0;
  // The truncation flags (EPI / GENERIC instantiations only) are bit-exact statements about the REFERENCE's observation,
  // abs(2 (x - min) / (max - min) - 1) > 1, so they are evaluated on the correctly rounded quotient (div_inv), not on
  // the one-FMA value that is returned: a state exactly on the edge of its range must not be flagged.
  // GENERIC: the reference expression itself (IEEE division); EPI: threshold compares prepared on the host with that
  // expression (hp.thr, exenv_math.cuh: make_trunc_thresholds) -- the same flags, two compares per entry.
  auto obs_for_flags = [&](const T *sv, T *obx) {
    if constexpr (!Env::IS_PMSM) {  // PMSM's flag is computed from the state (trunc_entry)
#pragma unroll
      for (int o = 0; o < O; ++o) obx[o] = norm_mm<T>(sv[o], pn[2 * o], pn[2 * o + 1]);
    }
  };
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
  // Every staged index is loaded and prepared exactly once.  Staged indices are loaded strictly in order: ap points
  // at the raw action of the next one (staged q -> raw action q - dead; the PMSM buffer entry q < dead carries no
  // load; past the end the last action is re-read, never used).
  const int dead = Env::IS_PMSM ? a.deadtime : 0;
  const int q_max = (int)a.n_actions - 1 + dead;  // last staged index that exists
  TmFetch<T, A, W> win[V];
  T pend[V][A];
  int wbase = 0;
  const T *ap = a.actions + b * A;
  int qn = 0;
  const unsigned act_stride = (unsigned)(a.B * A);
  constexpr int PF_AHEAD = 6;  // solver steps between the L2 prefetch of an action row segment and its load
  auto load_next = [&](T (*raw)[A]) {
    if (!Env::IS_PMSM || qn >= dead) {
      load_vec<T, V * A>(&raw[0][0], ap);
      // the segment this thread will load PF_AHEAD slides from now starts its trip from DRAM to L2 here (no register,
      // no scoreboard entry), so the load above finds its data in L2
      if (qn + PF_AHEAD <= q_max)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(ap + (size_t)PF_AHEAD * act_stride));
      if (qn < q_max) ap += act_stride;
    } else {
#pragma unroll
      for (int v = 0; v < V; ++v)
#pragma unroll
        for (int c = 0; c < A; ++c) raw[v][c] = (T)0;
    }
    ++qn;
  };
  auto prepare = [&](int q, const T (*raw)[A], int k) {
    q = q < q_max ? q : q_max;
#pragma unroll
    for (int v = 0; v < V; ++v)
      prepare_action<Env, T>((long)q, raw[v], s0[v], p, pre, a.deadtime, a.n_actions, a.pmsm_adv, a.pmsm_lin_stop,
                             win[v].uw[k]);
  };
#pragma unroll
  for (int k = 0; k < W; ++k) {
    T raw[V][A];
    load_next(raw);
    prepare(k, raw, k);
  }
  load_next(pend);
  bool pend_valid = true;  // `pend` holds the raw action of staged index wbase + W
  auto slide_to = [&](int base) {
#pragma unroll 1
    while (wbase < base) {
      if (!pend_valid) load_next(pend);  // only when one step slides the window by more than one entry
#pragma unroll
      for (int v = 0; v < V; ++v)
#pragma unroll
        for (int k = 0; k + 1 < W; ++k)
#pragma unroll
          for (int c = 0; c < A; ++c) win[v].uw[k][c] = win[v].uw[k + 1][c];
      ++wbase;
      prepare(wbase + W - 1, pend, W - 1);
      pend_valid = false;
    }
#pragma unroll
    for (int v = 0; v < V; ++v) win[v].wbase = wbase;
  };

  // ---- saves ------------------------------------------------------------------------------------------------------
  // Fast path: the first save of a step (the only one, except for the step that also saves ts[0]) is stashed per
  // environment and written after all V environments of the thread have stepped: one vector store per trajectory
  // leaf and per observation row group, through running row pointers.  Further saves of the same step and the
  // GENERIC instantiation store element-wise.
  const unsigned Bu = (unsigned)a.B;
  // ONE running element index addresses every output of the fast path: (save step_first_save, environment b) of an
  // (N+1, B) array; leaf / observation / flag addresses are formed from the kernel-argument base pointers at the store
  // (the same two integer instructions a running pointer per array would spend on its increment, and 2 (D + 3) fewer
  // live registers).  reward / terminated rows are those of states[1:]: one row behind.
  size_t srow = (size_t)b;
  constexpr int NT0 = trunc_dim<Env>(0);  // truncated width without reference entries
  auto flags_fast = [&](const T *sv, const T *, uint8_t *tr, uint8_t &te) {  // control_state == []
    bool t0 = false;
#pragma unroll
    for (int i = 0; i < NT0; ++i) {
      bool tv;
      if constexpr (Env::IS_PMSM || Env::ID == 2) tv = trunc_entry<Env, T>(i, sv, sv, p);  // from the state / constant
      else tv = trunc_by_threshold<T>(sv[i], hp.thr[i]);
      if (i == 0) t0 = tv;
      tr[i] = tv;
    }
    te = terminated_of<Env, T>((T)0, t0);
  };
  T stash[V][NST];
  long step_first_save = 0;  // save index the running pointers address
  auto obs_fast = [&](int v, const T *st, const T *sv, T *ob) {
    if constexpr (Env::IS_PMSM) {
      ob[0] = norm_fma<T>(sv[3], nf[3]);
      ob[1] = norm_fma<T>(sv[4], nf[4]);
      ob[2] = om_norm[v];
      ob[3] = norm_fma<T>(sv[5], nf[5]);
      ob[4] = st[6];
      ob[5] = st[7];
      ob[6] = norm_fma<T>(sv[0], nf[0]);
      ob[7] = norm_fma<T>(sv[1], nf[1]);
    } else {
#pragma unroll
      for (int o = 0; o < O; ++o) ob[o] = norm_fma<T>(sv[o], nf[o]);
    }
  };
  auto sink_v = [&](int v, long sidx, const T *st) {
    if constexpr (GENERIC) {  // (the fast path reads last_state back from the trajectory rows it has written)
#pragma unroll
      for (int i = 0; i < NST; ++i) last[v][i] = st[i];
    }
    if constexpr (!GENERIC) {
      if (sidx == step_first_save) {
#pragma unroll
        for (int i = 0; i < NST; ++i) stash[v][i] = st[i];
      } else {  // rare: a later save of the same step
        const size_t row = srow + (size_t)(sidx - step_first_save) * Bu + v;
        T sv[D], ob[O];
        staged_to_fields<Env, T>(st, s0[v][D - 1], sv);
#pragma unroll
        for (int d = 0; d < D; ++d) __stcs(a.traj[d] + row, sv[d]);
        obs_fast(v, st, sv, ob);
#pragma unroll
        for (int o = 0; o < O; ++o) __stcs(a.obs + row * O + o, ob[o]);
        if constexpr (EPI) {
          uint8_t tr[NT0], te;
          flags_fast(sv, ob, tr, te);
#pragma unroll
          for (int i = 0; i < NT0; ++i) a.trunc[row * NT0 + i] = tr[i];
          // reward / terminated rows run one save behind (states[1:]); sidx > step_first_save >= 0 here
          __stcs(a.reward + (row - Bu), (T)0);
          a.term[row - Bu] = te;
        }
      }
    } else {
      const size_t row = (size_t)sidx * Bu + b;
      T sv[D];
      staged_to_fields<Env, T>(st, s0[0][D - 1], sv);
#pragma unroll
      for (int d = 0; d < D; ++d)
        if (a.traj[d]) __stcs(a.traj[d] + row, sv[d]);
      T ob[O + EXENV_MAX_CONTROL];
      obs_fast(0, st, sv, ob);
#pragma unroll
      for (int j = 0; j < EXENV_MAX_CONTROL; ++j)
        if (j < a.n_control) ob[O + j] = norm_field<Env, T>(a.control[j], rf, p);
      T *dst = a.obs + row * OW;
#pragma unroll
      for (int o = 0; o < O + EXENV_MAX_CONTROL; ++o)
        if (o < OW) __stcs(dst + o, ob[o]);
      if (a.reward || a.trunc || a.term) {
        bool t0 = false;
        T obx[O + EXENV_MAX_CONTROL];
#pragma unroll
        for (int o = O; o < O + EXENV_MAX_CONTROL; ++o) obx[o] = ob[o];  // reference entries: exact already
        obs_for_flags(sv, obx);
#pragma unroll
        for (int i = 0; i < O + EXENV_MAX_CONTROL; ++i)
          if (i < nt) {
            const bool tv = trunc_entry<Env, T>(i, sv, obx, p);
            if (i == 0) t0 = tv;
            if (a.trunc) a.trunc[row * nt + i] = tv;
          }
        if (sidx >= 1) {
          const T rw = a.n_control > 0 ? reward_of<Env, T>(sv, rf, p, a.n_control, a.control) : (T)0;
          const size_t r1 = row - Bu;
          if (a.reward) __stcs(a.reward + r1, rw);
          if (a.term) a.term[r1] = terminated_of<Env, T>(rw, t0);
        }
      }
    }
  };
  auto flush = [&](int ns) {  // vector stores of the stashed first save, then the row pointers move on by ns saves
    if constexpr (!GENERIC) {
      T sv[V][D], ob[V][O];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        staged_to_fields<Env, T>(stash[v], s0[v][D - 1], sv[v]);
        obs_fast(v, stash[v], sv[v], ob[v]);
      }
#pragma unroll
      for (int d = 0; d < D; ++d) {
        T col[V];
#pragma unroll
        for (int v = 0; v < V; ++v) col[v] = sv[v][d];
        store_vec<T, V>(a.traj[d] + srow, col);
      }
      store_vec<T, V * O>(a.obs + srow * O, &ob[0][0]);
      if constexpr (EPI) {
        uint8_t tr[V][NT0], te[V];
        T zero[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
          flags_fast(sv[v], ob[v], tr[v], te[v]);
          zero[v] = (T)0;
        }
        store_vec<uint8_t, V * NT0>(a.trunc + srow * NT0, &tr[0][0]);
        if (step_first_save > 0) {  // save 0 has no reward / terminated entry
          store_vec<T, V>(a.reward + (srow - Bu), zero);
          store_vec<uint8_t, V>(a.term + (srow - Bu), te);
        }
      }
      srow += (size_t)Bu * (unsigned)ns;
    }
  };

  // ---- the solver loop: the schedule record of a step (word, step length, dense-output record of its first save)
  // comes out of this warp's shared-memory copy (TmSchedReader); the only global load of the loop is the action stream's
  long save_idx = 0;
  TmSchedReader<T> sched;
  sched.init(a.steps, smem_raw + sched_off);
  {
    const uint64_t w0 = sched.chunk(0)->w;
    slide_to(sched_base(w0));
    if constexpr (SOLVER == EXENV_TSIT5) {
#pragma unroll
      for (int v = 0; v < V; ++v) sim_first_derivative<Env, T>(w0, y[v], f[v], p, pre, aux[v], win[v]);
    }
  }
  const int n_steps = (int)a.n_steps;
#pragma unroll 1
  for (int i0 = 0; i0 < n_steps; i0 += TM_CHUNK) {  // 32 steps per staged chunk of schedule records
  const TmStep<T> *sr = sched.chunk(i0 / TM_CHUNK);
  const int n_in = min(TM_CHUNK, n_steps - i0);
#pragma unroll 1
  for (int k = 0; k < n_in; ++k, ++sr) {
    const uint64_t w = sr->w;
    const T dt = sr->dt;
    // the window slide consumes the action that left for DRAM one step ago; the next action load is issued right after
    slide_to(sched_base(w));
    SaveRec<T> rec;
#pragma unroll
    for (int j = 0; j < (SOLVER == EXENV_EULER ? 1 : 7); ++j) rec.v[j] = sr->wts[j];
    if (!pend_valid) {
      load_next(pend);
      pend_valid = true;
    }
    step_first_save = save_idx;
    long si = save_idx;
    if (W > 2 && SOLVER == EXENV_TSIT5 && simtm_dual_path<Env>() && sched_dmax(w) <= 1) {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        si = save_idx;
        const TmFetchPair<T, A, W> pair{win[v]};
        sim_one_step<Env, SOLVER, T>(w, dt, y[v], f[v], p, pre, aux[v], rec, a.recs, si, pair,
                                     [&](long sidx, const T *st) { sink_v(v, sidx, st); });
      }
    } else {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        si = save_idx;
        sim_one_step<Env, SOLVER, T>(w, dt, y[v], f[v], p, pre, aux[v], rec, a.recs, si, win[v],
                                     [&](long sidx, const T *st) { sink_v(v, sidx, st); });
      }
    }
    const int ns = sched_saves(w);
    if (ns > 0) flush(ns);
    save_idx = si;
  }
  }

  // ---- last_state (core_env.py:484-488) and the fresh solver.init derivative at (t1, y_last) --------------------
  const int q_last = (int)(a.n_actions - 1);
  const bool want_k = SOLVER == EXENV_TSIT5 && a.klast[0];
  if (want_k && q_last > wbase + W - 1) slide_to(q_last - (W - 1));
  (void)pend_valid;
  T sl[V][D], fi[V][NODE];
  if constexpr (!GENERIC) {  // the fields of the last save, as this thread has just written them
#pragma unroll
    for (int d = 0; d < D; ++d) {
      T col[V];
      load_vec<T, V>(col, a.traj[d] + ((size_t)(a.n_save - 1) * Bu + (size_t)b));
#pragma unroll
      for (int v = 0; v < V; ++v) sl[v][d] = col[v];
    }
  }
#pragma unroll
  for (int v = 0; v < V; ++v) {
    if constexpr (GENERIC) staged_to_fields<Env, T>(last[v], s0[v][D - 1], sl[v]);
    if (want_k) {
      // solver.init(term, t1, t1 + tau, y_last, args): derivative with the ZOH closure at t1 -> clamped last action
      T u[A], yl[NODE];
      win[v](q_last, u);
      if constexpr (Env::IS_PMSM) {  // ODE state (i_d, i_q, epsilon) = fields 3, 4, 2
        yl[0] = sl[v][3];
        yl[1] = sl[v][4];
        yl[2] = sl[v][2];
      } else {
#pragma unroll
        for (int i = 0; i < NODE; ++i) yl[i] = sl[v][i];
      }
      sim_rhs<Env, T>(yl, u, p, pre, aux[v], fi[v]);
    }
  }
#pragma unroll
  for (int d = 0; d < D; ++d)
    if (a.last[d]) {
      T col[V];
#pragma unroll
      for (int v = 0; v < V; ++v) col[v] = sl[v][d];
      store_vec<T, V>(a.last[d] + b, col);
    }
  if (want_k) {
#pragma unroll
    for (int i = 0; i < NODE; ++i) {
      T col[V];
#pragma unroll
      for (int v = 0; v < V; ++v) col[v] = fi[v][i];
      store_vec<T, V>(a.klast[i] + b, col);
    }
  }
  if constexpr (!PERSIST) return;
  }  // next job of this warp
}

template <class Env, int SOLVER, typename T, bool PER_ENV, bool GENERIC, int W, int V, bool EPI>
static int launch_simtm_v(const ExenvPlan &plan, const ExenvSimArgs &x, const SimSchedule &sch, cudaStream_t stream) {
  SimTmK<T> a = simtm_args<Env, T>(plan, x, sch);
  KParams<T> kp = make_kparams<T>(plan);
  constexpr bool COEF = simtm_is_coef<Env>();
  if constexpr (COEF) {  // the kernel stages the coefficient-form table instead of the node blob
    a.lut_blob = plan.lut.coef;
    a.lut.bytes = (unsigned)plan.lut.coef_bytes;
  }
  size_t smem = 0;
  if (Env::IS_PMSM && (plan.desc.flags & EXENV_FLAG_SATURATED)) smem = ((size_t)a.lut.bytes + 16 + 15) & ~(size_t)15;
  smem += tm_sched_smem_bytes<T>(simtm_cta_threads<Env>());
  auto kern = simtm_kernel<Env, SOLVER, T, PER_ENV, GENERIC, W, V, EPI>;
  if (smem > 48 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("sim_ahead (time-major) needs %zu bytes of shared memory per CTA: %s", smem, cudaGetErrorString(e));
      return EXENV_ERR_CUDA;
    }
  }
  const long threads = (x.batch + V - 1) / V;
  const int nthr = COEF ? SIMTM_COEF_THREADS : simtm_block_threads();
  unsigned grid = (unsigned)((threads + nthr - 1) / nthr);
  if (COEF) grid = std::min<unsigned>(grid, (unsigned)device_sm_count(plan.device));  // persistent: one CTA per SM
  bool thr_ok = true;
  const SimTmHostPre<Env, T> hp = make_host_pre<Env, T>(kp, EPI ? &thr_ok : nullptr);
  if (!thr_ok) {
    set_error("sim_ahead (time-major): degenerate normalisation range for the fused truncation flags");
    return EXENV_ERR_INVALID;
  }
  kern<<<grid, nthr, smem, stream>>>(a, kp, hp);
  g_launch_count++;
  const cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("exenv_sim_ahead (time-major) launch failed: %s", cudaGetErrorString(e));
    return EXENV_ERR_CUDA;
  }
  return EXENV_OK;
}


// EXENV_SIMTM_PACKED=0 in the environment selects the scalar kernel where the packed one would run (A/B measurements).
static bool simtm_packed_enabled() {
  static const bool on = [] {
    const char *e = getenv("EXENV_SIMTM_PACKED");
    return !(e && e[0] == '0');
  }();
  return on;
}

// Window length: 2 covers every schedule whose steps read at most two consecutive ZOH indices (always the case up to
// floating-point lag of the accumulated solver time); 4 is the general case (build_schedule_t rejects more).
template <class Env, int SOLVER, typename T, bool PER_ENV, bool GENERIC, int V, bool EPI = false>
static int launch_simtm_w(const ExenvPlan &plan, const ExenvSimArgs &x, const SimSchedule &sch, cudaStream_t stream) {
  if (sch.win_need <= 2) return launch_simtm_v<Env, SOLVER, T, PER_ENV, GENERIC, 2, V, EPI>(plan, x, sch, stream);
  if (sch.win_need <= 4) return launch_simtm_v<Env, SOLVER, T, PER_ENV, GENERIC, 4, V, EPI>(plan, x, sch, stream);
  set_error("sim_ahead (time-major): a solver step needs %d consecutive actions (> 4)", sch.win_need);
  return EXENV_ERR_UNSUPPORTED;
}

// Environments per thread of the fast instantiation (tools/quick_bench.py simtm sweep on B200, profiles/README.md):
// light right-hand sides gain from sharing the per-step bookkeeping between more environments; heavy ones (acrobot,
// saturated PMSM) are bound by their own arithmetic and registers.
template <class Env, int SOLVER, typename T>
constexpr int simtm_auto_v() {
  constexpr bool f32 = sizeof(T) == 4, tsit = SOLVER == EXENV_TSIT5;
  if (Env::IS_PMSM) return 1;
  switch (Env::ID) {
    case 0: return tsit ? (f32 ? 4 : 1) : (f32 ? 4 : 1);  // pendulum (fp32: packed, profiles/README.md round 2)
    case 1: return tsit ? (f32 ? 4 : 1) : (f32 ? 2 : 1);  // mass-spring-damper
    case 2: return tsit ? (f32 ? 2 : 1) : (f32 ? 4 : 2);  // fluid tank
    case 3: return tsit ? (f32 ? 2 : 1) : 2;              // cart-pole (fp32 Tsit5: one packed pair per thread)
    default: return tsit ? (f32 ? 2 : 1) : (f32 ? 2 : 1);  // acrobot
  }
}

template <class Env, int SOLVER, typename T>
static int launch_simtm_fast(const ExenvPlan &plan, const ExenvSimArgs &x, const SimSchedule &sch, cudaStream_t stream) {
  // instantiated widths: 1, (2, 4).  Four fp64 environments of a four-component system per thread would need > 255
  // registers (the instantiation spilled and was never the tuned choice): those stop at two.
  constexpr int VMAX = Env::IS_PMSM ? 1 : ((sizeof(T) == 8 && Env::NODE >= 4) ? 2 : 4);
  int v = x.block_envs > 0 ? x.block_envs : simtm_auto_v<Env, SOLVER, T>();
  if (v > VMAX) v = VMAX;
  if (v == 3) v = 2;
  // small batches: threads first (one 148-SM wave of 1024 threads per SM), environments per thread second
  if (x.block_envs <= 0)
    while (v > 1 && x.batch / v < 148L * 1024) v >>= 1;
  // V consecutive environments per thread move as one vector: the batch must split evenly and every buffer must be
  // aligned to the vector (at most 16 bytes per access)
  auto ok = [&](int vv) {
    if (vv == 1) return true;
    if (x.batch % vv) return false;
    const size_t leaf = std::min<size_t>(16, (size_t)vv * sizeof(T));
    const size_t arow = std::min<size_t>(16, (size_t)vv * Env::A * sizeof(T));
    const size_t orow = std::min<size_t>(16, (size_t)vv * Env::O * sizeof(T));
    auto al = [](const void *p, size_t n) { return reinterpret_cast<uintptr_t>(p) % n == 0; };
    bool r = al(x.actions, arow) && al(x.obs, orow);
    for (int d = 0; d < Env::D; ++d) r = r && al(x.state_in[d], leaf) && al(x.traj[d], leaf) && al(x.last[d], leaf);
    for (int d = 0; d < Env::NODE; ++d) r = r && al(x.k_last[d], leaf);
    if (x.reward) {
      const size_t trow = std::min<size_t>(16, (size_t)vv * trunc_dim<Env>(0));
      r = r && al(x.reward, leaf) && al(x.truncated, trow) && al(x.terminated, (size_t)vv);
    }
    return r;
  };
  while (v > 1 && !ok(v)) v >>= 1;
  // packed fp32x2 arithmetic flavour (its own translation unit: exenv_simtm_packed_f32.cu)
  if constexpr (std::is_same<T, float>::value && env_packable<Env>::value) {
    if (simtm_packed_enabled() && (v == 2 || v == 4)) return launch_simtm_packed_f32(plan, x, sch, stream, v);
  }
  if (x.reward) {  // fused epilogue (control_state == []): the tuned width or one environment per thread
    constexpr int VA = simtm_auto_v<Env, SOLVER, T>() > VMAX ? VMAX : simtm_auto_v<Env, SOLVER, T>();
    if (v == VA) return launch_simtm_w<Env, SOLVER, T, false, false, VA, true>(plan, x, sch, stream);
    return launch_simtm_w<Env, SOLVER, T, false, false, 1, true>(plan, x, sch, stream);
  }
  if constexpr (VMAX >= 4) {
    if (v == 4) return launch_simtm_w<Env, SOLVER, T, false, false, 4>(plan, x, sch, stream);
  }
  if constexpr (VMAX >= 2) {
    if (v == 2) return launch_simtm_w<Env, SOLVER, T, false, false, 2>(plan, x, sch, stream);
  }
  return launch_simtm_w<Env, SOLVER, T, false, false, 1>(plan, x, sch, stream);
}

template <class Env, typename T>
static int simtm_dispatch2(const ExenvPlan &plan, const ExenvSimArgs &x, const SimSchedule &s, cudaStream_t st) {
  const bool tsit = plan.desc.solver == EXENV_TSIT5;
  // anything beyond the plain call (optionally with all three epilogue outputs) takes the GENERIC instantiation:
  // reference entries, a partial epilogue, skipped (NULL) trajectory leaves, observation rows that cannot be written
  // with aligned vector stores
  const bool epi_any = x.reward || x.truncated || x.terminated, epi_all = x.reward && x.truncated && x.terminated;
  bool generic = plan.desc.n_control > 0 || (epi_any && !epi_all);
  if (epi_all && !generic) generic = !simtm_thresholds_ok<Env, T>(plan);  // degenerate ranges: evaluate the expression
  for (int d = 0; d < Env::D; ++d) generic |= x.traj[d] == nullptr;
  {
    const size_t row_bytes = (size_t)Env::O * sizeof(T);
    const size_t vec = row_bytes < 16 ? row_bytes : 16;
    generic |= reinterpret_cast<uintptr_t>(x.obs) % vec != 0;
    generic |= reinterpret_cast<uintptr_t>(x.actions) % ((size_t)Env::A * sizeof(T)) != 0;
  }
  if constexpr (!Env::IS_PMSM) {
    if (plan.per_env)  // per-environment parameters always take the generic instantiation
      return tsit ? launch_simtm_w<Env, EXENV_TSIT5, T, true, true, 1>(plan, x, s, st)
                  : launch_simtm_w<Env, EXENV_EULER, T, true, true, 1>(plan, x, s, st);
  }
  if (generic)
    return tsit ? launch_simtm_w<Env, EXENV_TSIT5, T, false, true, 1>(plan, x, s, st)
                : launch_simtm_w<Env, EXENV_EULER, T, false, true, 1>(plan, x, s, st);
#if !defined(EXENV_SIMTM_COEF_TU)
  if constexpr (Env::IS_PMSM && Env::SATURATED && std::is_same<T, float>::value) {
    if (simtm_coef_fits(plan)) return launch_simtm_pmsm_coef_f32(plan, x, s, st);  // exenv_simtm_pmsm_coef_f32.cu
  }
#endif
  return tsit ? launch_simtm_fast<Env, EXENV_TSIT5, T>(plan, x, s, st) : launch_simtm_fast<Env, EXENV_EULER, T>(plan, x, s, st);
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

#if defined(EXENV_SIMTM_COEF_TU)
// trajectories of the saturated PMSM with coefficient-form flux maps, with and without the fused epilogue
int launch_simtm_pmsm_coef_f32(const ExenvPlan &plan, const ExenvSimArgs &a, const SimSchedule &sch, cudaStream_t stream) {
  if (plan.desc.solver == EXENV_TSIT5) return launch_simtm_fast<Pmsm<float, 2>, EXENV_TSIT5, float>(plan, a, sch, stream);
  return launch_simtm_fast<Pmsm<float, 2>, EXENV_EULER, float>(plan, a, sch, stream);
}
#else
int EXENV_CAT(launch_simtm_, EXENV_SFX)(const ExenvPlan &plan, const ExenvSimArgs &a, const SimSchedule &sch,
                                         cudaStream_t stream) {
  return simtm_dispatch<EXENV_REAL>(plan, a, sch, stream);
}
#endif

}  // namespace exenv
