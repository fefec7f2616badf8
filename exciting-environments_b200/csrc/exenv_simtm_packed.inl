// exenv_simtm_packed.inl (fp32 only; included inside namespace exenv by exenv_simtm_packed_f32.cu) -- the time-major
// trajectory kernel with PACKED fp32 arithmetic: a thread owns V = 2 or 4 consecutive environments as V / 2 values of
// type f32x2, so every solver operation is one FFMA2 / FMUL2 / FADD2 for two environments (exenv_math.cuh).
//
// Why: the scalar kernel is bound by instruction issue (profiles/README.md: 178 warp instructions per env-step at
// 86 % issue-slot utilisation, of which 127 are fp32 arithmetic, while the FMA pipe is 62 % busy).  Packed
// instructions run at the same arithmetic rate but take half the issue slots, which moves the limit from the
// schedulers to the FMA pipe itself.  Results are the same IEEE operations in the same order.
//
// Scope: environments marked env_packable (straight-line right-hand side: mass-spring-damper), shared parameters,
// control_state == [], optionally the fused reward / truncated / terminated outputs.  Everything else runs the scalar
// kernel of exenv_simtm.inl, which this file mirrors step for step.
//
// Zero-order-hold actions: a two-entry window {action q, action q + 1} of PREPARED actions in registers plus the raw
// action q + 2 in flight.  Every regular step reads at most two consecutive indices and advances q by 0 or 1; the rare
// steps that touch 3-4 indices or jump further (the accumulated fp32 solver time produces a handful per trajectory)
// take a second copy of the step body that loads its actions straight from global memory.
// Resident CTAs per SM the register allocation aims at: 128 registers for four environments per thread of the
// two-component systems and for ONE PAIR per thread of the four-component ones (cart-pole, acrobot: seven stage
// derivatives of four packed components = 56 registers per pair; with the parameter-only products of their right-hand
// sides prepared on the host the acrobot fits without a spill, the cart-pole with 32-56 bytes -- four CTAs per SM
// instead of three: +5 % / +4 %, profiles/README.md); two pairs per thread of those: two CTAs.
template <class Env, int V>
constexpr int packed_min_blocks() {
  return Env::NODE >= 4 ? (V == 4 ? 2 : 4) : (V == 4 ? 4 : 5);
}
// Solver steps per trip of the step loop: two for the mass-spring-damper (its loop-carried state then ping-pongs
// between two register sets instead of being copied back after every step: 44 MOVs per thread-step); one for the
// environments with sine / cosine / reciprocal code, whose step body is several times larger (instruction cache).
template <class Env>
constexpr int packed_unroll() {
  return Env::ID == 1 ? 2 : 1;
}
template <class Env, int SOLVER, int V, bool EPI>
__global__ void __launch_bounds__(SIMTM_THREADS, packed_min_blocks<Env, V>())
    simtm_packed_kernel(const __grid_constant__ SimTmK<float> a, const __grid_constant__ KParams<float> kp,
                        const __grid_constant__ SimTmHostPre<Env, float> hp) {
  using T = float;
  using P = f32x2;
  using EnvP = typename Env::template rebind<P>;
  static_assert(V % 2 == 0 && !Env::IS_PMSM, "pairs of environments, no PMSM");
  constexpr int D = Env::D, A = Env::A, NODE = Env::NODE, O = Env::O, NP = num_params<Env>();
  constexpr int NST = staged_values<Env>(), VV = V / 2;
  constexpr int NT0 = trunc_dim<Env>(0);

  // first environment of this thread; surplus lanes of the warp that straddles the end of the batch shadow the last
  // group of V environments (see exenv_simtm.inl)
  const long b_own = ((long)blockIdx.x * blockDim.x + threadIdx.x) * V;
  if (b_own - (long)(threadIdx.x & 31) * V >= a.B) return;
  const long b = b_own < a.B ? b_own : a.B - V;

  // shared parameters: scalars (constant bank); the packed copy is a splat the compiler folds into broadcast operands
  const T *p = kp.v;
  P pq[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) pq[i] = P(p[i]);
  constexpr bool HOST_PRE = simtm_use_host_pre<Env, SOLVER>();  // host-prepared scalars, broadcast operands
  typename EnvP::Pre pre_thread;
  if constexpr (HOST_PRE) pre_thread = EnvP::from_scalar(hp.pre);
  else pre_thread = EnvP::precompute(pq);
  const typename EnvP::Pre &pre = pre_thread;
  const P *pn = pq + Env::NSTATIC;
  typename EnvP::Aux aux{};

  P s0[VV][D], y[VV][NODE], f[VV][NODE];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    T tmp[V];
    load_vec<T, V>(tmp, a.sin[d] + b);
#pragma unroll
    for (int vv = 0; vv < VV; ++vv) s0[vv][d] = P(tmp[2 * vv], tmp[2 * vv + 1]);
  }
#pragma unroll
  for (int vv = 0; vv < VV; ++vv) {
#pragma unroll
    for (int i = 0; i < NODE; ++i) {
      y[vv][i] = s0[vv][i];
      f[vv][i] = P(0.0f);
    }
  }
  NormFma<P> nf[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    if constexpr (HOST_PRE) {
      nf[d].c1 = P(hp.nf[d].c1);
      nf[d].c0 = P(hp.nf[d].c0);
    } else {
      nf[d] = make_normfma<P>(pn[2 * d], pn[2 * d + 1]);
    }
  }

  // ---- window of prepared ZOH actions ----------------------------------------------------------------------------------
  const int q_max = (int)a.n_actions - 1;
  const unsigned act_stride = (unsigned)(a.B * A);
  const T *act0 = a.actions + b * A;  // this thread's V * A scalars of action 0; action q is q * act_stride further
  constexpr int PF_AHEAD = 6;         // solver steps between the L2 prefetch of an action row segment and its load
  auto load_raw = [&](int q, T (*raw)[A]) {  // indices past the end re-read the last action (never used)
    q = q < q_max ? q : q_max;
    load_vec<T, V * A>(&raw[0][0], act0 + (size_t)q * act_stride);
  };
  auto prep = [&](const T (*raw)[A], P (*u)[A]) {
#pragma unroll
    for (int vv = 0; vv < VV; ++vv) {
      P rp[A];
#pragma unroll
      for (int c = 0; c < A; ++c) rp[c] = P(raw[2 * vv][c], raw[2 * vv + 1][c]);
      prepare_action<EnvP, P>(0L, rp, s0[vv], pq, pre, 0, a.n_actions, P(0.0f), P(0.0f), u[vv]);
    }
  };
  P u0[VV][A], u1[VV][A];  // prepared actions wbase, wbase + 1
  T pend[V][A];            // raw action wbase + 2, requested one step before its first possible use
  int wbase = 0;
  auto resync = [&](int base) {  // window := {base, base + 1}, pend := base + 2, by direct loads
    T r0[V][A], r1[V][A];
    load_raw(base, r0);
    load_raw(base + 1, r1);
    load_raw(base + 2, pend);
    prep(r0, u0);
    prep(r1, u1);
    wbase = base;
  };
  auto advance_to = [&](int base) {  // warp-uniform
    if (base == wbase + 1) {
#pragma unroll
      for (int vv = 0; vv < VV; ++vv)
#pragma unroll
        for (int c = 0; c < A; ++c) u0[vv][c] = u1[vv][c];
      prep(pend, u1);
      ++wbase;
      load_raw(wbase + 2, pend);
      // the segment this thread will load PF_AHEAD steps from now starts its trip from DRAM to L2 here
      if (wbase + 2 + PF_AHEAD <= q_max)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(act0 + (size_t)(wbase + 2 + PF_AHEAD) * act_stride));
    } else if (base != wbase) {
      resync(base);
    }
  };
  // the two fetch objects of exenv_sim_core.cuh: the window (stage offset 0 / 1), and direct loads for the rare steps
  struct PairFetch {
    const P (*u0)[A];
    const P (*u1)[A];
    int vv, wbase;
    __device__ __forceinline__ void stage(uint64_t w, int s, P *u) const {
      const bool hi = (unsigned)(w >> (32 + 2 * s)) & 1u;
#pragma unroll
      for (int c = 0; c < A; ++c) u[c] = hi ? u1[vv][c] : u0[vv][c];
    }
    __device__ __forceinline__ void operator()(int idx, P *u) const {
#pragma unroll
      for (int c = 0; c < A; ++c) u[c] = idx > wbase ? u1[vv][c] : u0[vv][c];
    }
  };
  auto direct_fetch = [&](int vv) {
    return [&, vv](int idx, P *u) {
      idx = idx < q_max ? idx : q_max;
      T raw[2][A];
      load_vec<T, 2 * A>(&raw[0][0], act0 + (size_t)idx * act_stride + 2 * vv * A);
      P rp[A];
#pragma unroll
      for (int c = 0; c < A; ++c) rp[c] = P(raw[0][c], raw[1][c]);
      prepare_action<EnvP, P>(0L, rp, s0[vv], pq, pre, 0, a.n_actions, P(0.0f), P(0.0f), u);
    };
  };

  // ---- saves: stash the step's first save per pair, one vector store per leaf / observation group afterwards -----
  const unsigned Bu = (unsigned)a.B;
  // one running element index (save step_first_save, environment b) of an (N+1, B) array addresses every output; the
  // addresses are formed from the kernel-argument base pointers at the store (exenv_simtm.inl)
  size_t srow = (size_t)b;
  P stash[VV][NST];
  long step_first_save = 0;
  auto observe_pair = [&](const P *sv, P *ob) {
#pragma unroll
    for (int o = 0; o < O; ++o) ob[o] = norm_fma<P>(sv[o], nf[o]);
  };
  // flags: from the correctly rounded reference-form observation (see exenv_simtm.inl), per lane
  auto flags_lane = [&](const T *sv, const T *, uint8_t *tr, uint8_t &te) {  // control_state == []
    bool t0 = false;
#pragma unroll
    for (int i = 0; i < NT0; ++i) {
      bool tv;
      if constexpr (Env::ID == 2) tv = false;  // fluid tank: constant (fluid_tank_env.py:323-326)
      else tv = trunc_by_threshold<T>(sv[i], hp.thr[i]);
      if (i == 0) t0 = tv;
      tr[i] = tv;
    }
    te = terminated_of<Env, T>((T)0, t0);
  };
  auto sink_pair = [&](int vv, long sidx, const P *st) {
    if (sidx == step_first_save) {
#pragma unroll
      for (int i = 0; i < NST; ++i) stash[vv][i] = st[i];
    } else {  // rare: a later save of the same step, element-wise
      P svp[D], obp[O];
      staged_to_fields<EnvP, P>(st, s0[vv][D - 1], svp);
      observe_pair(svp, obp);
#pragma unroll
      for (int l = 0; l < 2; ++l) {
        const size_t row = srow + (size_t)(sidx - step_first_save) * Bu + 2 * vv + l;
        T sv[D], ob[O];
#pragma unroll
        for (int d = 0; d < D; ++d) {
          sv[d] = l ? svp[d].y : svp[d].x;
          __stcs(a.traj[d] + row, sv[d]);
        }
#pragma unroll
        for (int o = 0; o < O; ++o) {
          ob[o] = l ? obp[o].y : obp[o].x;
          __stcs(a.obs + row * O + o, ob[o]);
        }
        if constexpr (EPI) {
          uint8_t tr[NT0], te;
          flags_lane(sv, ob, tr, te);
#pragma unroll
          for (int i = 0; i < NT0; ++i) a.trunc[row * NT0 + i] = tr[i];
          __stcs(a.reward + (row - Bu), (T)0);  // rows of states[1:]: one behind (sidx >= 1 here)
          a.term[row - Bu] = te;
        }
      }
    }
  };
  auto flush = [&](int ns) {
    T sv[V][D], ob[V][O];
#pragma unroll
    for (int vv = 0; vv < VV; ++vv) {
      P svp[D], obp[O];
      staged_to_fields<EnvP, P>(stash[vv], s0[vv][D - 1], svp);
      observe_pair(svp, obp);
#pragma unroll
      for (int d = 0; d < D; ++d) {
        sv[2 * vv][d] = svp[d].x;
        sv[2 * vv + 1][d] = svp[d].y;
      }
#pragma unroll
      for (int o = 0; o < O; ++o) {
        ob[2 * vv][o] = obp[o].x;
        ob[2 * vv + 1][o] = obp[o].y;
      }
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
        flags_lane(sv[v], ob[v], tr[v], te[v]);
        zero[v] = (T)0;
      }
      store_vec<uint8_t, V * NT0>(a.trunc + srow * NT0, &tr[0][0]);
      if (step_first_save > 0) {  // save 0 has no reward / terminated entry
        store_vec<T, V>(a.reward + (srow - Bu), zero);
        store_vec<uint8_t, V>(a.term + (srow - Bu), te);
      }
    }
    srow += (size_t)Bu * (unsigned)ns;
  };

  // ---- the solver loop (schedule records from this warp's shared-memory copy, see exenv_simtm_common.cuh) ---------
  long save_idx = 0;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TmSchedReader<T> sched;
  sched.init(a.steps, smem_raw);
  {
    const uint64_t w0 = sched.chunk(0)->w;
    resync(sched_base(w0));
    if constexpr (SOLVER == EXENV_TSIT5) {
#pragma unroll
      for (int vv = 0; vv < VV; ++vv)
        sim_first_derivative<EnvP, P>(w0, y[vv], f[vv], pq, pre, aux, PairFetch{u0, u1, vv, wbase});
    }
  }
  const int n_steps = (int)a.n_steps;
#pragma unroll 1
  for (int i0 = 0; i0 < n_steps; i0 += TM_CHUNK) {  // 32 steps per staged chunk of schedule records
  const TmStep<T> *sr = sched.chunk(i0 / TM_CHUNK);
  const int n_in = min(TM_CHUNK, n_steps - i0);
  // two steps per trip: the loop-carried state (y, FSAL derivative, window) ping-pongs between two register sets
  // instead of being copied back at the end of every step (44 MOVs per thread-step in the single-step loop)
#pragma unroll(packed_unroll<Env>())
  for (int k = 0; k < n_in; ++k, ++sr) {
    const uint64_t w = sr->w;
    const P dt = P(sr->dt);
    // consumes the action that left for DRAM one step ago and requests the next one
    advance_to(sched_base(w));
    SaveRec<T> rec;
#pragma unroll
    for (int j = 0; j < (SOLVER == EXENV_EULER ? 1 : 7); ++j) rec.v[j] = sr->wts[j];
    step_first_save = save_idx;
    long si = save_idx;
    if (sched_dmax(w) <= 1) {
#pragma unroll
      for (int vv = 0; vv < VV; ++vv) {
        si = save_idx;
        sim_one_step<EnvP, SOLVER, P>(w, dt, y[vv], f[vv], pq, pre, aux, rec, a.recs, si, PairFetch{u0, u1, vv, wbase},
                                      [&](long sidx, const P *st) { sink_pair(vv, sidx, st); });
      }
    } else {  // rare: a step that reads three or four consecutive actions
#pragma unroll
      for (int vv = 0; vv < VV; ++vv) {
        si = save_idx;
        sim_one_step<EnvP, SOLVER, P>(w, dt, y[vv], f[vv], pq, pre, aux, rec, a.recs, si, direct_fetch(vv),
                                      [&](long sidx, const P *st) { sink_pair(vv, sidx, st); });
      }
    }
    const int ns = sched_saves(w);
    if (ns > 0) flush(ns);
    save_idx = si;
  }
  }

  // ---- last_state (core_env.py:484-488) = the last saved sample, read back from the trajectory rows this thread has
  // just written (nothing is carried through the loop for it), and the fresh solver.init derivative at (t1, y_last) ----
  const int q_last = (int)(a.n_actions - 1);
  const bool want_k = SOLVER == EXENV_TSIT5 && a.klast[0];
  T sl[D][V];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const T *src = a.traj[d] + (size_t)(a.n_save - 1) * Bu + b;
#pragma unroll
    for (int v = 0; v < V; ++v) sl[d][v] = src[v];
    if (a.last[d]) store_vec<T, V>(a.last[d] + b, sl[d]);
  }
  if (want_k) {
    T fi[NODE][V];
#pragma unroll
    for (int vv = 0; vv < VV; ++vv) {
      P u[A], yl[NODE], fp[NODE];
      direct_fetch(vv)(q_last, u);  // the ZOH closure at t1: clamped to the last action
#pragma unroll
      for (int i = 0; i < NODE; ++i) yl[i] = P(sl[i][2 * vv], sl[i][2 * vv + 1]);
      sim_rhs<EnvP, P>(yl, u, pq, pre, aux, fp);
#pragma unroll
      for (int i = 0; i < NODE; ++i) {
        fi[i][2 * vv] = fp[i].x;
        fi[i][2 * vv + 1] = fp[i].y;
      }
    }
#pragma unroll
    for (int i = 0; i < NODE; ++i) store_vec<T, V>(a.klast[i] + b, fi[i]);
  }
}

// Launch (the argument block is the scalar kernel's; the launcher in exenv_simtm.inl has already checked alignment and
// that the batch splits into groups of V).
template <class Env, int SOLVER, int V, bool EPI>
static int launch_simtm_packed_v(const SimTmK<float> &a, const KParams<float> &kp, long batch, cudaStream_t stream) {
  const long threads = (batch + V - 1) / V;
  const int nthr = simtm_block_threads();
  const unsigned grid = (unsigned)((threads + nthr - 1) / nthr);
  bool thr_ok = true;
  const SimTmHostPre<Env, float> hp = make_host_pre<Env, float>(kp, EPI ? &thr_ok : nullptr);
  if (!thr_ok) {
    set_error("sim_ahead (time-major, packed): degenerate normalisation range for the fused truncation flags");
    return EXENV_ERR_INVALID;
  }
  simtm_packed_kernel<Env, SOLVER, V, EPI><<<grid, nthr, tm_sched_smem_bytes<float>(SIMTM_THREADS), stream>>>(a, kp, hp);
  g_launch_count++;
  const cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("exenv_sim_ahead (time-major, packed) launch failed: %s", cudaGetErrorString(e));
    return EXENV_ERR_CUDA;
  }
  return EXENV_OK;
}

