// exenv_simtm_common.cuh -- pieces shared by the time-major trajectory kernels (exenv_simtm.inl: scalar arithmetic, every
// environment / dtype; exenv_simtm_packed.inl: packed fp32x2 arithmetic): kernel-argument block, explicit vector loads
// and streaming stores, the zero-order-hold action window, host-prepared constants.
#pragma once

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "exenv_kernel_common.cuh"
#include "exenv_sim_core.cuh"

namespace exenv {

constexpr int SIMTM_THREADS = 128;  // threads per CTA (launch bounds)
// EXENV_SIMTM_THREADS (32, 64 or 96) launches smaller CTAs for A/B measurements.  Measured on the B200 (profiles/
// README.md): a trajectory kernel's grid runs as a few long rounds of resident CTAs (2^20 environments, four per
// thread, 512 resident threads per SM: 3.46 rounds), but the partially filled last round costs nothing -- SMs with
// fewer resident warps finish their CTAs proportionally faster (a batch of exactly three rounds is not faster per
// environment), and CTA sizes that make the round count integral (224 x 2) lose 12 % to register allocation.
static int simtm_block_threads() {
  static const int n = [] {
    const char *e = getenv("EXENV_SIMTM_THREADS");
    const int v = e ? atoi(e) : 0;
    return (v >= 32 && v <= SIMTM_THREADS && v % 32 == 0) ? v : SIMTM_THREADS;
  }();
  return n;
}

// Saturated PMSM with coefficient-form flux maps (Pmsm<float, 2>): the table of a whole motor (BRUSA: 157 KB) fits one
// CTA's shared memory once per SM, so that flavour runs ONE persistent CTA per SM and hands whole warps their
// environments: warp w of CTA c takes the 32-environment jobs w * gridDim.x + c, + warps * gridDim.x, ... (warp-major,
// so that the jobs of a partially filled last round are spread over all SMs: an SM whose later warps have retired runs
// the others faster, as with the small CTAs of the other flavours).
// 512 threads = 16 warps at <= 128 registers (the kernel needs 113-119 once the six reads of a look-up are fenced,
// exenv_envs.cuh: lut_eval_coef).
template <class Env>
constexpr bool simtm_is_coef() {
  if constexpr (Env::IS_PMSM) return Env::COEF;
  else return false;
}
#ifndef EXENV_COEF_THREADS
#define EXENV_COEF_THREADS 512
#endif
constexpr int SIMTM_COEF_THREADS = EXENV_COEF_THREADS;
template <class Env>
constexpr int simtm_cta_threads() {
  return simtm_is_coef<Env>() ? SIMTM_COEF_THREADS : SIMTM_THREADS;
}

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
  const TmStep<T> *steps;  // per solver step: schedule word, step length, first-save record (staged through smem)
  const SaveRec<T> *recs;
  int n_control;
  int control[EXENV_MAX_CONTROL];
  int deadtime;
  T pmsm_adv, pmsm_lin_stop;
  const void *lut_blob;
  LutGeom<T> lut;
};

__device__ __forceinline__ unsigned tm_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// Streaming (evict-first) access of N consecutive scalars (N * sizeof(T) a power of two; 16-byte chunks at most).
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
  } else if constexpr (BYTES == 4) {
    float t;
    memcpy(&t, src, 4);
    __stcs(reinterpret_cast<float *>(dst), t);
  } else if constexpr (BYTES == 2) {
    unsigned short t;
    memcpy(&t, src, 2);
    __stcs(reinterpret_cast<unsigned short *>(dst), t);
  } else {
    __stcs(reinterpret_cast<unsigned char *>(dst), *reinterpret_cast<const unsigned char *>(src));
  }
}
// Loads are explicit PTX of the width the kernel wants, in the order the solver loop wants them issued (ptxas is still
// free to merge and move them; the order below is what keeps it from parking a scoreboard wait behind a DRAM access,
// see profiles/README.md).
template <int BYTES, bool NC>
__device__ __forceinline__ void ld_chunk(void *dst, const void *src) {
  if constexpr (BYTES == 16) {
    uint4 t;
    if constexpr (NC)
      asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(src));
    else
      asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(src));
    memcpy(dst, &t, 16);
  } else if constexpr (BYTES == 8) {
    uint2 t;
    if constexpr (NC)
      asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(t.x), "=r"(t.y) : "l"(src));
    else
      asm volatile("ld.global.cs.v2.u32 {%0,%1}, [%2];" : "=r"(t.x), "=r"(t.y) : "l"(src));
    memcpy(dst, &t, 8);
  } else {
    unsigned t;
    if constexpr (NC)
      asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(t) : "l"(src));
    else
      asm volatile("ld.global.cs.u32 %0, [%1];" : "=r"(t) : "l"(src));
    memcpy(dst, &t, 4);
  }
}
template <int BYTES>
__device__ __forceinline__ void ldcs_chunk(void *dst, const void *src) {
  ld_chunk<BYTES, false>(dst, src);
}
// Marks values as used at this point of the program (compiler-level only: it emits no instruction).
__device__ __forceinline__ void consume_now(unsigned a) { asm volatile("" ::"r"(a)); }
template <typename T, int N>
__device__ __forceinline__ void store_vec(T *dst, const T *v) {
  constexpr int ROW = N * (int)sizeof(T);
  constexpr int CH = ROW < 16 ? ROW : 16;
#pragma unroll
  for (int i = 0; i < ROW / CH; ++i)
    stcs_chunk<CH>(reinterpret_cast<char *>(dst) + i * CH, reinterpret_cast<const char *>(v) + i * CH);
}
template <typename T, int N>
__device__ __forceinline__ void load_vec(T *v, const T *src) {
  constexpr int ROW = N * (int)sizeof(T);
  constexpr int CH = ROW < 16 ? ROW : 16;
#pragma unroll
  for (int i = 0; i < ROW / CH; ++i)
    ldcs_chunk<CH>(reinterpret_cast<char *>(v) + i * CH, reinterpret_cast<const char *>(src) + i * CH);
}

// ---- the schedule, staged through shared memory -------------------------------------------------------------------------
// Every warp of the grid walks the same per-step records (TmStep: word, step length, first-save record).  Reading them
// with global loads put them on the scoreboard of the action stream's DRAM load (ptxas shares one scoreboard between
// the loads of the loop), so the first use of a record waited for DRAM, and a lane of a merged 16-byte load that the
// step does not use made that wait a write-after-write stall at the top of the step: 22-27 % of the stall samples of
// the headline kernel (profiles/README.md, round 2).  Each warp therefore copies the records of the next 32 steps into
// its own shared-memory buffer with cp.async (LDGSTS: no register, no scoreboard; one record per lane, double
// buffered, one __syncwarp per 32 steps) and reads them back with broadcast LDS.
template <typename T>
constexpr size_t tm_sched_smem_bytes(int threads) {
  return (size_t)(threads / 32) * 2 * TM_CHUNK * sizeof(TmStep<T>);
}
template <typename T>
struct TmSchedReader {
  const TmStep<T> *g;  // global records (padded to whole chunks + one, exenv_api.cu: upload_records)
  TmStep<T> *buf;      // this warp's two chunk buffers
  __device__ __forceinline__ void stage(int chunk) {  // lane l copies the record of step chunk * 32 + l
    constexpr int CH16 = (int)sizeof(TmStep<T>) / 16;
    const unsigned lane = threadIdx.x & 31u;
    const char *src = reinterpret_cast<const char *>(g + (size_t)chunk * TM_CHUNK + lane);
    const unsigned dst = (unsigned)__cvta_generic_to_shared(buf + (chunk & 1) * TM_CHUNK + lane);
#pragma unroll
    for (int c = 0; c < CH16; ++c)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + c * 16), "l"(src + c * 16) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // Call once, with all 32 lanes of the warp, before the first at().  (A warp that straddles the end of the batch keeps
  // its surplus lanes: they shadow the last environment group, see the kernels.)
  __device__ __forceinline__ void init(const TmStep<T> *records, unsigned char *smem_base) {
    g = records;
    buf = reinterpret_cast<TmStep<T> *>(smem_base) + (size_t)(threadIdx.x >> 5) * 2 * TM_CHUNK;
    // a persistent warp starting its next job: the look-ahead copy of the previous one has landed and every lane is done
    // reading the buffers
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    stage(0);
    enter_chunk(0);
  }
  // chunk c has landed (it was requested 32 steps ago); request chunk c + 1 into the other buffer, whose previous
  // content (chunk c - 1) every lane has finished reading
  __device__ __forceinline__ void enter_chunk(int c) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    stage(c + 1);
  }
  // records of steps [32 c, 32 c + 32); chunks must be visited in order (chunk 0 is ready after init())
  __device__ __forceinline__ const TmStep<T> *chunk(int c) {
    if (c) enter_chunk(c);
    return buf + (c & 1) * TM_CHUNK;
  }
};

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

// The same window for a step whose stages read at most two consecutive indices (every step, up to the rare lag of
// the accumulated solver time): one select per stage on the low offset bit.
template <typename T, int A, int W>
struct TmFetchPair {
  const TmFetch<T, A, W> &f;
  __device__ __forceinline__ void stage(uint64_t w, int s, T *u) const {
    const bool hi = (unsigned)(w >> (32 + 2 * s)) & 1u;
#pragma unroll
    for (int c = 0; c < A; ++c) u[c] = hi ? f.uw[1][c] : f.uw[0][c];
  }
  __device__ __forceinline__ void operator()(int idx, T *u) const { f(idx, u); }
};

// Light right-hand sides carry a second copy of the step body for such steps (W = 4 windows only).
template <class Env>
constexpr bool simtm_dual_path() {
  return Env::ID == 0 || Env::ID == 1 || Env::ID == 2;
}

// PhysicalState fields of a saved sample from its staged values (PMSM: eps, i_d, i_q, torque, u_d_buffer, u_q_buffer).
template <class Env, typename T>
__device__ __forceinline__ void staged_to_fields(const T *st, T omega_el, T *sv) {
  if constexpr (Env::IS_PMSM) {
    sv[0] = st[4];
    sv[1] = st[5];
    sv[2] = st[0];
    sv[3] = st[1];
    sv[4] = st[2];
    sv[5] = st[3];
    sv[6] = omega_el;
  } else {
#pragma unroll
    for (int d = 0; d < Env::D; ++d) sv[d] = st[d];
  }
}

// V consecutive environments per thread: their leaves / action rows / observation rows are moved with one vector
// access each, and the per-step bookkeeping (schedule fetch and decode, window slide, pointer updates, loop) is paid
// once for V env-steps.  V > 1 needs B % V == 0 and V * sizeof(T)-aligned buffers (the launcher checks).
// EPI: the fast instantiation with vmap_generate_rew_trunc_term_ahead (core_env.py:618-647) fused for
// control_state == [] (reward is the constant 0 there): truncated for every saved state, reward / terminated for
// states[1:], written as V-wide vectors like the leaves.
// GENERIC (V == 1 only): reference entries in the observation (control_state != []), the fused reward / truncated /
// terminated outputs (vmap_generate_rew_trunc_term_ahead, core_env.py:618-647), per-environment parameters, skipped
// (NULL) trajectory leaves.
// Constants derived from SHARED parameters (Env::precompute, the observation normalisation) are prepared on the host
// with the same IEEE expressions and travel in the kernel-argument constant bank: operands instead of registers (the
// fp64 mass-spring-damper keeps two environments per thread inside 128 registers that way).  Per-environment
// parameters are still prepared per thread.
// Measured per environment (profiles/README.md): PMSM 123 -> 96 registers (+13 %), acrobot fp64 +19 %, fp64
// mass-spring-damper +7 %; pendulum and fluid tank lose 7-9 % (their allocation was already minimal and the constant-bank
// operands lengthen the dependent chains), so they keep the per-thread preparation.
// The bandwidth-bound Euler kernels lose 2-3 % as well: Tsit5 only.
template <class Env, int SOLVER>
constexpr bool simtm_use_host_pre() {
  return SOLVER == EXENV_TSIT5 && (Env::ID == 1 || Env::ID == 3 || Env::ID == 4 || Env::ID == 5);
}
template <class Env, typename T>
struct SimTmHostPre {
  typename Env::Pre pre;
  NormFma<T> nf[Env::D];
  TruncThr<T> thr[Env::D];  // truncation flags of the fused epilogue as threshold compares (exenv_math.cuh)
};
template <class Env, typename T>
static SimTmHostPre<Env, T> make_host_pre(const KParams<T> &kp, bool *thr_ok = nullptr) {
  SimTmHostPre<Env, T> h;
  h.pre = Env::precompute(kp.v);
  const T *pn = kp.v + Env::NSTATIC;
  bool ok = true;
  for (int d = 0; d < Env::D; ++d) {
    h.nf[d] = make_normfma<T>(pn[2 * d], pn[2 * d + 1]);
    h.thr[d].lo_last = h.thr[d].hi_first = (T)NAN;
    if (thr_ok && !Env::IS_PMSM && Env::ID != 2) ok = make_trunc_thresholds<T>(pn[2 * d], pn[2 * d + 1], h.thr[d]) && ok;
  }
  if (thr_ok) *thr_ok = ok;
  return h;
}
// Shared-parameter ranges for which the threshold form of the truncation flags exists (every sane range)
template <class Env, typename T>
static bool simtm_thresholds_ok(const ExenvPlan &plan) {
  if (plan.per_env) return false;
  bool ok = true;
  make_host_pre<Env, T>(make_kparams<T>(plan), &ok);
  return ok;
}

// ---- kernel-argument block from the ABI arguments ---------------------------------------------------------------------
template <class Env, typename T>
static SimTmK<T> simtm_args(const ExenvPlan &plan, const ExenvSimArgs &x, const SimSchedule &sch) {
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
  a.steps = (const TmStep<T> *)sch.d_steprec;
  a.recs = (const SaveRec<T> *)sch.d_theta;
  a.n_control = plan.desc.n_control;
  for (int j = 0; j < EXENV_MAX_CONTROL; ++j) a.control[j] = plan.desc.control[j];
  a.deadtime = plan.deadtime;
  a.pmsm_adv = (T)plan.pmsm_adv;
  a.pmsm_lin_stop = (T)(plan.desc.tau * (double)(x.n_actions - 1));
  a.lut_blob = plan.lut.blob;
  a.lut = make_lut_geom<T>(plan);
  return a;
}


}  // namespace exenv
