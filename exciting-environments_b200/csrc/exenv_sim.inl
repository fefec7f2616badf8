// exenv_sim.inl (included by exenv_sim_f32.cu / exenv_sim_f64.cu with EXENV_REAL / EXENV_SFX) -- vmap_sim_ahead as ONE
// persistent multi-step sm_100a kernel.
//
// Thread e of a CTA owns environment (blockIdx.x * E + e) for the whole trajectory: its ODE state, the Tsit5 FSAL
// derivative and the stage increments stay in registers across all solver steps.  The batch-major API layout
// (actions (B, T, A) in; observations (B, N+1, O) and D trajectory leaves (B, N+1) out -- reference
// core_env.py:571-616) is bridged through shared-memory tiles of `tile_steps` solver steps:
//
//   prefetch     : the action span of tile t+1 is copied global -> shared with cp.async (LDGSTS, 4/8-byte elements,
//                  lanes along time = contiguous row segments in HBM, written transposed as [time][env] with an odd
//                  pitch); it is issued as soon as tile t's integration is finished and lands while tile t's outputs
//                  are being written: no register staging, no exposed latency
//   in-place     : each thread de-normalises its own action column (PMSM: dq->alpha/beta, voltage-hexagon clip at
//                  the advanced angle, dead-time shift -- pmsm_env.py:709-777)
//   compute      : tile_steps solver steps per thread; every save is the solver's dense output evaluated at the
//                  host-computed abscissa, wrapped / clipped, (PMSM: + torque) and staged to smem as [slot][env]
//   store phase  : warps walk environment rows with lanes along time and write each trajectory leaf and the
//                  observation rows as contiguous segments; a lane keeps a fixed observation column, so its
//                  normalisation constants (min, prepared reciprocal of the span) live in registers.  The optional
//                  reward / truncated / terminated outputs (vmap_generate_rew_trunc_term_ahead, core_env.py:618-647),
//                  reference entries and per-environment parameters take a generic (slower) row writer.
//
// All time arithmetic of diffrax.diffeqsolve (accumulated tnext, clip to t1, truncating ZOH index, linspace save grid,
// dense-output abscissae) is environment independent; it is emulated once on the host in the plan dtype
// (exenv_api.cu: build_schedule) and streamed to the kernel as one 64-bit word + dt per step and one theta per save.
// PMSM flux maps are staged once per CTA into shared memory with a TMA bulk copy (cp.async.bulk + mbarrier).
//
// Algorithmic HBM traffic per env-step: S * (A + O + D) bytes.
#include <cstdio>

#include "exenv_kernel_common.cuh"
#include "exenv_sim_core.cuh"

#define EXENV_CAT_(a, b) a##b
#define EXENV_CAT(a, b) EXENV_CAT_(a, b)

namespace exenv {

constexpr int SIM_E = 128;        // environments (threads) per CTA
constexpr int SIM_EP = SIM_E + 1; // smem pitch of a [time][env] tile row (odd: conflict-free both ways)
constexpr int SIM_NW = SIM_E / 32;

template <typename T>
struct SimK {
  long B, n_actions, n_save;
  const T *sin[EXENV_MAX_PHYS];
  const T *actions;
  T *obs;
  T *traj[EXENV_MAX_PHYS];
  T *last[EXENV_MAX_PHYS];
  T *klast[EXENV_MAX_ODE];
  const T *ref[EXENV_MAX_PHYS];
  T *reward;
  uint8_t *trunc;
  uint8_t *term;
  const uint64_t *words;
  const T *dts;
  const SaveRec<T> *recs;
  const SimTile *tiles;
  int n_tiles, act_cap, save_cap, step_cap;
  int n_control;
  int control[EXENV_MAX_CONTROL];
  int deadtime;
  T pmsm_adv, pmsm_lin_stop;  // (deadtime+0.5)*tau ; tau*(T-1) = linspace stop of pmsm_env.py:721
  const void *lut_blob;
  LutGeom<T> lut;
};

// bytes of the per-tile schedule slice in shared memory (records, words, step lengths), 16-byte aligned
template <typename T>
__host__ __device__ constexpr size_t sim_sched_bytes(int step_cap, int save_cap) {
  return ((size_t)(save_cap + 1) * sizeof(SaveRec<T>) + (size_t)(step_cap + 2) * 8 + (size_t)(step_cap + 2) * sizeof(T) + 15) &
         ~(size_t)15;
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_load(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(mbar))
               : "memory");
}

// Element-granular asynchronous copy global -> shared (SASS: LDGSTS).
template <typename T>
__device__ __forceinline__ void cp_async_elem(T *dst_smem, const T *src_gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "n"(sizeof(T))
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Issue the action span of one tile: rows = environments, lanes along the contiguous (time, action) axis.
// PMSM: staged indices address the dead-time shifted sequence; entries below `deadtime` are the initial buffer
// (filled by the owner thread later), the rest map to raw action q - deadtime.
template <class Env, typename T>
__device__ __forceinline__ void prefetch_actions(const SimK<T> &a, const SimTile &ti, T *sm_act, long blk0, int rows,
                                                 int warp, int lane) {
  constexpr int A = Env::A;
  const int dead = Env::IS_PMSM ? a.deadtime : 0;
  const int k0 = (ti.act_lo < dead) ? (dead - ti.act_lo) : 0;
  const int n_el = (ti.act_cnt - k0) * A;  // contiguous scalars per environment row
  const long src_lo = (long)ti.act_lo + k0 - dead;
  const long row_stride = (long)SIM_NW * a.n_actions * A;
  for (int j0 = 0; j0 < n_el; j0 += 32) {  // one pass per 32 scalars of the span (1-3 passes)
    const int j = j0 + lane;
    const bool on = j < n_el;
    const T *src = a.actions + ((blk0 + warp) * a.n_actions + src_lo) * A + j;
    unsigned dst = smem_u32(sm_act + (size_t)(k0 * A + j) * SIM_EP + warp);
    for (int r = warp; r < rows; r += SIM_NW) {
      if (on) asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst), "l"(src), "n"(sizeof(T)) : "memory");
      src += row_stride;
      dst += SIM_NW * sizeof(T);
    }
  }
}

// The tile's slice of the host-built schedule (step words, step lengths, dense-output records; each with one
// look-ahead entry) is copied next to the actions, so the step loop only touches shared memory.
template <typename T>
struct SchedSmem {
  uint64_t *words;
  T *dts;
  SaveRec<T> *recs;
};
template <typename T>
__device__ __forceinline__ void prefetch_schedule(const SimK<T> &a, const SimTile &ti, const SchedSmem<T> &sm, int e) {
  for (int i = e; i <= ti.step_count; i += SIM_E) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(sm.words + i)),
                 "l"(a.words + ti.step_begin + i)
                 : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(sm.dts + i)),
                 "l"(a.dts + ti.step_begin + i), "n"(sizeof(T))
                 : "memory");
  }
  constexpr int CH = sizeof(SaveRec<T>) / 16;  // 16-byte chunks per record
  const char *src = reinterpret_cast<const char *>(a.recs + ti.save_begin);
  char *dst = reinterpret_cast<char *>(sm.recs);
  for (int i = e; i < (ti.save_cnt + 1) * CH; i += SIM_E)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + 16 * i)), "l"(src + 16 * i) : "memory");
}

// Physical-state field `d` of (row r, slot) from the staged values (PMSM: omega_el is constant per environment).
template <class Env, typename T>
__device__ __forceinline__ T staged_field(const SimK<T> &a, const T *sm_out, long blk0, int d, int r, int slot) {
  constexpr int NST = staged_values<Env>();
  if constexpr (Env::IS_PMSM) {
    if (d == 6) return __ldg(a.sin[6] + blk0 + r);
    // staged order: eps, i_d, i_q, torque, u_d_buffer, u_q_buffer
    const int v = (d >= 2) ? d - 2 : d + 4;
    return sm_out[(slot * NST + v) * SIM_EP + r];
  } else {
    return sm_out[(slot * NST + d) * SIM_EP + r];
  }
}

// PMSM observation entry o -> PhysicalState field it reads (pmsm_env.py:905-916)
__device__ __forceinline__ int pmsm_obs_field(int o) {
  return (o == 0) ? 3 : (o == 1) ? 4 : (o == 2) ? 6 : (o == 3) ? 5 : (o == 6) ? 0 : (o == 7) ? 1 : 2;
}

// Fast row writer: control_state == [], no epilogue, shared parameters.
// A warp-wide store covers 32 consecutive scalars of one environment row when the tile holds more than 16 saves;
// shorter tiles pack 2 or 4 rows into one store so that lanes stay busy.  A lane keeps a fixed (slot, observation
// column), so its normalisation constants and smem/global offsets are loop invariants.  Rows are processed four at
// a time (four shared-memory loads in flight before the first store) to keep the LDS latency off the critical path.
template <class Env, typename T>
__device__ __forceinline__ void store_chunk_fast(const SimK<T> &a, const KParams<T> &kp, int save_begin, int cnt,
                                                 const T *sm_out, long blk0, int rows, int warp, int lane);

template <class Env, typename T>
__device__ __forceinline__ void store_chunk_fast(const SimK<T> &a, const KParams<T> &kp, int save_begin, int cnt,
                                                 const T *sm_out, long blk0, int rows, int warp, int lane);

// Long tiles (more than 32 saves): every environment row receives its whole run of `cnt` samples in consecutive
// warp-wide stores (row outer, 32-slot chunks inner), so that HBM sees segments of cnt * 4 bytes (leaves) and
// cnt * O * 4 bytes (observations) instead of isolated 128-byte lines.
template <class Env, typename T>
__device__ __forceinline__ void store_tile_long(const SimK<T> &a, const KParams<T> &kp, const SimTile &ti,
                                                const T *sm_out, long blk0, int rows, int warp, int lane) {
  constexpr int D = Env::D, O = Env::O, NST = staged_values<Env>();
  constexpr int LO = (O == 1) ? 0 : (O == 2) ? 1 : (O == 4) ? 2 : 3;
  const int cnt = ti.save_cnt;
  const T *pn = kp.v + Env::NSTATIC;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    T *leaf = a.traj[d];
    if (!leaf) continue;
    const int v = Env::IS_PMSM ? ((d >= 2) ? d - 2 : d + 4) : d;
    for (int r = warp; r < rows; r += SIM_NW) {
      T *dst = leaf + (blk0 + r) * a.n_save + ti.save_begin;
      const T om = (Env::IS_PMSM && d == 6) ? __ldg(a.sin[6] + blk0 + r) : (T)0;
      for (int k = lane; k < cnt; k += 32)
        dst[k] = (Env::IS_PMSM && d == 6) ? om : sm_out[(k * NST + v) * SIM_EP + r];
    }
  }
  {
    const int o = lane & (O - 1);
    int fld = o;
    if constexpr (Env::IS_PMSM) fld = pmsm_obs_field(o);
    T mn = pn[0];
    T span = pn[1] - pn[0];
#pragma unroll
    for (int i = 1; i < D; ++i)
      if (fld == i) {
        mn = pn[2 * i];
        span = pn[2 * i + 1] - pn[2 * i];
      }
    const InvDiv<T> sdiv = make_invdiv<T>(span);
    const int v = Env::IS_PMSM ? ((fld >= 2) ? fld - 2 : fld + 4) : fld;
    const int n_el = cnt * O;
    for (int r = warp; r < rows; r += SIM_NW) {
      T *dst = a.obs + ((blk0 + r) * a.n_save + ti.save_begin) * O;
      for (int j = lane; j < n_el; j += 32) {
        const int k = j >> LO;
        T x;
        if (Env::IS_PMSM && fld == 6) x = __ldg(a.sin[6] + blk0 + r);
        else x = sm_out[(k * NST + v) * SIM_EP + r];
        T val;
        if constexpr (Env::IS_PMSM) {
          val = (o == 4) ? Mth<T>::cos(x) : (o == 5) ? Mth<T>::sin(x) : norm_inv<true, T>(x, mn, sdiv);
        } else {
          val = norm_inv<true, T>(x, mn, sdiv);
        }
        dst[j] = val;
      }
    }
  }
}

template <class Env, typename T>
__device__ __forceinline__ void store_chunk_fast(const SimK<T> &a, const KParams<T> &kp, int save_begin, int cnt,
                                                 const T *sm_out, long blk0, int rows, int warp, int lane);

template <class Env, typename T>
__device__ __forceinline__ void store_tile_fast(const SimK<T> &a, const KParams<T> &kp, const SimTile &ti,
                                                const T *sm_out, long blk0, int rows, int warp, int lane) {
  constexpr int NST = staged_values<Env>();
  for (int c0 = 0; c0 < ti.save_cnt; c0 += 32) {  // 32 slots per pass: one full 128-byte store per leaf row
    const int cnt = ti.save_cnt - c0 < 32 ? ti.save_cnt - c0 : 32;
    store_chunk_fast<Env, T>(a, kp, ti.save_begin + c0, cnt, sm_out + (size_t)c0 * NST * SIM_EP, blk0, rows, warp, lane);
  }
}

template <class Env, typename T>
__device__ __forceinline__ void store_chunk_fast(const SimK<T> &a, const KParams<T> &kp, int save_begin, int cnt,
                                                 const T *sm_out, long blk0, int rows, int warp, int lane) {
  constexpr int D = Env::D, O = Env::O, NST = staged_values<Env>(), U = 4;
  const T *pn = kp.v + Env::NSTATIC;
  const int lp = cnt > 16 ? 5 : (cnt > 8 ? 4 : 3);  // log2 of the padded slot count P
  // ---- trajectory leaves: P slots per row, 32 / P rows per warp-wide store
  {
    const int k = lane & ((1 << lp) - 1), rs = lane >> lp, R = 32 >> lp;
    const bool on = k < cnt;
    const int rstep = SIM_NW * R;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      T *leaf = a.traj[d];
      if (!leaf) continue;
      if (Env::IS_PMSM && d == 6) {  // omega_el: constant per environment, broadcast over time
        for (int r = warp * R + rs; r < rows; r += rstep)
          if (on) leaf[(blk0 + r) * a.n_save + save_begin + k] = __ldg(a.sin[6] + blk0 + r);
        continue;
      }
      const int v = Env::IS_PMSM ? ((d >= 2) ? d - 2 : d + 4) : d;
      const T *src = sm_out + (k * NST + v) * SIM_EP + warp * R + rs;
      T *dst = leaf + (blk0 + warp * R + rs) * a.n_save + save_begin + k;
      const long stride = (long)rstep * a.n_save;
      for (int r = warp * R + rs; r < rows; r += U * rstep) {
        T x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) x[u] = (on && r + u * rstep < rows) ? src[u * rstep] : (T)0;
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (on && r + u * rstep < rows) __stcs(dst + u * stride, x[u]);
        src += U * rstep;
        dst += U * stride;
      }
    }
  }
  // ---- observation rows: P * O scalars per row (O divides 32: a lane owns one observation column)
  {
    constexpr int LO = (O == 1) ? 0 : (O == 2) ? 1 : (O == 4) ? 2 : 3;
    const int lrow = lp + LO;                   // log2 of the padded row length
    const int llanes = lrow < 5 ? lrow : 5;     // log2 lanes spent on one row
    const int iters = 1 << (lrow - llanes);     // warp-wide stores per row
    const int R = 32 >> llanes;                 // rows per warp-wide store
    const int jl = lane & ((1 << llanes) - 1), rs = lane >> llanes;
    const int o = lane & (O - 1);
    const int rstep = SIM_NW * R;
    int fld = o;
    if constexpr (Env::IS_PMSM) fld = pmsm_obs_field(o);
    T mn = pn[0];
    T span = pn[1] - pn[0];
#pragma unroll
    for (int i = 1; i < D; ++i)
      if (fld == i) {
        mn = pn[2 * i];
        span = pn[2 * i + 1] - pn[2 * i];
      }
    const InvDiv<T> sdiv = make_invdiv<T>(span);
    int v = Env::IS_PMSM ? ((fld >= 2) ? fld - 2 : fld + 4) : fld;
    const bool raw = Env::IS_PMSM && (o == 4 || o == 5);  // staged cos(eps) / sin(eps): copied as they are
    if (raw) v = o + 2;
    const long stride = (long)rstep * a.n_save * O;
    for (int it = 0; it < iters; ++it) {
      const int j = it * 32 + jl, k = j >> LO;
      const bool on = k < cnt;
      const T *src = sm_out + (k * NST + v) * SIM_EP + warp * R + rs;
      T *dst = a.obs + ((blk0 + warp * R + rs) * a.n_save + save_begin) * O + j;
      for (int r = warp * R + rs; r < rows; r += U * rstep) {
        T x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const bool act = on && r + u * rstep < rows;
          if (Env::IS_PMSM && fld == 6) x[u] = act ? __ldg(a.sin[6] + blk0 + r + u * rstep) : (T)0;
          else x[u] = act ? src[u * rstep] : (T)0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (on && r + u * rstep < rows) {
            const T nv = norm_inv<true, T>(x[u], mn, sdiv);
            __stcs(dst + u * stride, raw ? x[u] : nv);
          }
        }
        src += U * rstep;
        dst += U * stride;
      }
    }
  }
}

// Generic row writer: reference entries, per-environment parameters, fused reward / truncated / terminated.
template <class Env, typename T, bool PER_ENV>
__device__ __noinline__ void store_tile_generic(const SimK<T> &a, const KParams<T> &kp, const SimTile &ti,
                                                const T *sm_out, long blk0, int rows, int warp, int lane) {
  constexpr int D = Env::D, O = Env::O, NP = num_params<Env>();
  const int cnt = ti.save_cnt;
  const int OW = O + a.n_control;
  const bool epilogue = a.reward || a.trunc || a.term;
  const int nt = trunc_dim<Env>(a.n_control);
  for (int r = warp; r < rows; r += SIM_NW) {
    const long br = blk0 + r;
    T plr[NP];
    const T *pr = kp.v;
    if constexpr (PER_ENV) {
      load_env_params<Env, T, true>(kp, br, plr);
      pr = plr;
    }
    const T *pn = pr + Env::NSTATIC;
    const long row0 = br * a.n_save + ti.save_begin;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      T *dst = a.traj[d];
      if (dst)
        for (int k = lane; k < cnt; k += 32) dst[row0 + k] = staged_field<Env, T>(a, sm_out, blk0, d, r, k);
    }
    {
      T *dst = a.obs + row0 * OW;
      const int n_el = cnt * OW;
      for (int j = lane; j < n_el; j += 32) {
        const int k = j / OW, o = j - k * OW;
        T val;
        if (o < O) {
          int fld = o;
          if constexpr (Env::IS_PMSM) fld = pmsm_obs_field(o);
          const T fv = staged_field<Env, T>(a, sm_out, blk0, fld, r, k);
          T mn = pn[0], mx = pn[1];
#pragma unroll
          for (int i = 1; i < D; ++i)
            if (fld == i) {
              mn = pn[2 * i];
              mx = pn[2 * i + 1];
            }
          if (Env::IS_PMSM && o == 4) val = Mth<T>::cos(fv);
          else if (Env::IS_PMSM && o == 5) val = Mth<T>::sin(fv);
          else val = norm_mm<T>(fv, mn, mx);
        } else {
          const int fld = a.control[o - O];
          const T rv = __ldg(a.ref[fld] + br);
          T mn = pn[0], mx = pn[1];
#pragma unroll
          for (int i = 1; i < D; ++i)
            if (fld == i) {
              mn = pn[2 * i];
              mx = pn[2 * i + 1];
            }
          val = norm_mm<T>(rv, mn, mx);
        }
        dst[j] = val;
      }
    }
    // fused generate_rew_trunc_term_ahead: truncated for every saved state, reward / terminated for states[1:]
    if (epilogue) {
      for (int k = lane; k < cnt; k += 32) {
        T sv[D], rf[D], ob[O + EXENV_MAX_CONTROL];
#pragma unroll
        for (int d = 0; d < D; ++d) {
          sv[d] = staged_field<Env, T>(a, sm_out, blk0, d, r, k);
          rf[d] = (a.n_control > 0) ? __ldg(a.ref[d] + br) : (T)0;
        }
        observe<Env, T>(sv, pr, ob);
#pragma unroll
        for (int j = 0; j < EXENV_MAX_CONTROL; ++j)
          if (j < a.n_control) ob[O + j] = norm_field<Env, T>(a.control[j], rf, pr);
        const long gs = (long)ti.save_begin + k;  // global save index
        bool t0 = false;
#pragma unroll
        for (int i = 0; i < O + EXENV_MAX_CONTROL; ++i)
          if (i < nt) {
            const bool tv = trunc_entry<Env, T>(i, sv, ob, pr);
            if (i == 0) t0 = tv;
            if (a.trunc) a.trunc[(br * a.n_save + gs) * nt + i] = tv;
          }
        if (gs >= 1) {
          const T rw = a.n_control > 0 ? reward_of<Env, T>(sv, rf, pr, a.n_control, a.control) : (T)0;
          if (a.reward) a.reward[br * (a.n_save - 1) + gs - 1] = rw;
          if (a.term) a.term[br * (a.n_save - 1) + gs - 1] = terminated_of<Env, T>(rw, t0);
        }
      }
    }
  }
}

template <class Env, int SOLVER, typename T, bool PER_ENV>
__global__ void __launch_bounds__(SIM_E)
    sim_kernel(const __grid_constant__ SimK<T> a, const __grid_constant__ KParams<T> kp) {
  constexpr int D = Env::D, A = Env::A, NODE = Env::NODE, NP = num_params<Env>();
  constexpr int NST = staged_values<Env>();
  extern __shared__ __align__(16) unsigned char smem_raw[];

  const int e = threadIdx.x, lane = e & 31, warp = e >> 5;
  const long blk0 = (long)blockIdx.x * SIM_E;
  const long b = blk0 + e;
  const bool valid = b < a.B;
  const int rows = (int)((a.B - blk0) < (long)SIM_E ? (a.B - blk0) : (long)SIM_E);  // valid environments here

  const size_t act_elems = (size_t)a.act_cap * A * SIM_EP;
  SchedSmem<T> sched;
  sched.recs = reinterpret_cast<SaveRec<T> *>(smem_raw);                                    // [save_cap + 1]
  sched.words = reinterpret_cast<uint64_t *>(sched.recs + (a.save_cap + 1));               // [step_cap + 1]
  sched.dts = reinterpret_cast<T *>(sched.words + (a.step_cap + 2));                        // [step_cap + 1]
  T *sm_act = reinterpret_cast<T *>(smem_raw + sim_sched_bytes<T>(a.step_cap, a.save_cap));  // [act_cap * A][EP]
  T *sm_out = sm_act + act_elems;                                                           // [save_cap * NST][EP]

  // first tile's actions and schedule are on their way while the constants below are set up
  prefetch_actions<Env, T>(a, a.tiles[0], sm_act, blk0, rows, warp, lane);
  prefetch_schedule<T>(a, a.tiles[0], sched, e);
  cp_async_commit();

  typename Env::Aux aux{};
  if constexpr (Env::IS_PMSM && Env::SATURATED) {
    // stage the flux maps once per CTA: TMA bulk copy + mbarrier
    size_t off = (sim_sched_bytes<T>(a.step_cap, a.save_cap) + (act_elems + (size_t)a.save_cap * NST * SIM_EP) * sizeof(T) + 15) & ~(size_t)15;
    unsigned char *sm_lut = smem_raw + off;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(sm_lut + a.lut.bytes);
    if (e == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(a.lut.bytes)
                   : "memory");
      tma_bulk_load(sm_lut, a.lut_blob, a.lut.bytes, mbar);
    }
    __syncthreads();
    {
      unsigned done = 0;
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(mbar))
            : "memory");
      }
    }
    aux.lut = lut_view_at<T>(sm_lut, a.lut);
  }

  // ---- per-environment constants and initial state
  T pl[NP];
  const T *p = kp.v;
  if constexpr (PER_ENV) {
    load_env_params<Env, T, true>(kp, valid ? b : 0, pl);
    p = pl;
  }
  const typename Env::Pre pre = Env::precompute(p);
  T s0[D];
#pragma unroll
  for (int d = 0; d < D; ++d) s0[d] = valid ? __ldg(a.sin[d] + b) : (T)0;
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
  long save_idx = 0;
  const bool fast_store = !PER_ENV && a.n_control == 0 && !(a.reward || a.trunc || a.term);

  for (int tile = 0; tile < a.n_tiles; ++tile) {
    const SimTile ti = a.tiles[tile];
    cp_async_wait_all();
    __syncthreads();  // this tile's actions have landed; the previous store phase is done with sm_out

    if (valid) {
      T *col = sm_act + e;  // this environment's column of the [time][env] tile
      // ---- in place: normalised -> physical action ------------------------------------------------------------
      for (int k = 0; k < ti.act_cnt; ++k) {
        T raw[A], u[A];
#pragma unroll
        for (int c = 0; c < A; ++c) raw[c] = col[(k * A + c) * SIM_EP];
        prepare_action<Env, T>((long)ti.act_lo + k, raw, s0, p, pre, a.deadtime, a.n_actions, a.pmsm_adv,
                               a.pmsm_lin_stop, u);
#pragma unroll
        for (int c = 0; c < A; ++c) col[(k * A + c) * SIM_EP] = u[c];
      }

      // ---- compute phase --------------------------------------------------------------------------------------
      const T *col0 = col - (long)ti.act_lo * A * SIM_EP;  // indexable by absolute ZOH index
      T *out_col = sm_out + e - (long)ti.save_begin * NST * SIM_EP;
      auto fetch = [&](int idx, T *u) {
#pragma unroll
        for (int c = 0; c < A; ++c) u[c] = col0[(idx * A + c) * SIM_EP];
      };
      // the schedule entries (word, dt, first save record) of step i+1 are read while step i is integrated
      const SaveRec<T> *srec = sched.recs - ti.save_begin;  // indexable by absolute save index
      uint64_t w_next = sched.words[0];
      T dt_next = sched.dts[0];
      SaveRec<T> rec_next = srec[save_idx];
      if (SOLVER == EXENV_TSIT5 && tile == 0) sim_first_derivative<Env, T>(w_next, y, f, p, pre, aux, fetch);
      for (int i = 0; i < ti.step_count; ++i) {
        const uint64_t w = w_next;
        const T dt = dt_next;
        const SaveRec<T> rec = rec_next;
        w_next = sched.words[i + 1];
        dt_next = sched.dts[i + 1];
        rec_next = srec[save_idx + sched_saves(w)];
        sim_one_step<Env, SOLVER, T>(w, dt, y, f, p, pre, aux, rec, srec, save_idx, fetch,
                                     [&](long sidx, const T *st) {
                                       T *o = out_col + (long)sidx * NST * SIM_EP;
#pragma unroll
                                       for (int i2 = 0; i2 < NST; ++i2) {
                                         last[i2] = st[i2];
                                         o[i2 * SIM_EP] = st[i2];
                                       }
                                     });
      }
    } else {
      save_idx += ti.save_cnt;
    }
    __syncthreads();  // integration of this tile is finished: sm_out is complete, sm_act is free again

    // the next tile's actions travel while this tile's outputs are written
    const bool more = tile + 1 < a.n_tiles;
    if (more) {
      const SimTile tn = a.tiles[tile + 1];
      prefetch_actions<Env, T>(a, tn, sm_act, blk0, rows, warp, lane);
      prefetch_schedule<T>(a, tn, sched, e);
      cp_async_commit();
    }

    // ---- store phase: rows of the batch-major outputs, lanes along time ---------------------------------------
    if (ti.save_cnt > 0) {
      if constexpr (!PER_ENV) {
        if (fast_store)
          store_tile_fast<Env, T>(a, kp, ti, sm_out, blk0, rows, warp, lane);
        else
          store_tile_generic<Env, T, false>(a, kp, ti, sm_out, blk0, rows, warp, lane);
      } else {
        store_tile_generic<Env, T, true>(a, kp, ti, sm_out, blk0, rows, warp, lane);
      }
    }
  }

  // ---- last_state (core_env.py:484-488) and the fresh solver.init derivative at (t1, y_last) --------------------
  if (valid) {
    const SimTile tl = a.tiles[a.n_tiles - 1];
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
      const int k = (int)(a.n_actions - 1 - tl.act_lo);
      T u[A], yl[NODE], fi[NODE];
#pragma unroll
      for (int c = 0; c < A; ++c) u[c] = sm_act[(k * A + c) * SIM_EP + e];
      if constexpr (Env::IS_PMSM) {
        yl[0] = last[1];
        yl[1] = last[2];
        yl[2] = last[0];
      } else {
#pragma unroll
        for (int i = 0; i < NODE; ++i) yl[i] = last[i];
      }
      sim_rhs<Env, T>(yl, u, p, pre, aux, fi);
#pragma unroll
      for (int i = 0; i < NODE; ++i) a.klast[i][b] = fi[i];
    }
  }
}

// ---- launcher ----------------------------------------------------------------------------------------------------
template <class Env, typename T>
static size_t smem_bytes_t(const ExenvPlan &plan, int act_cap, int save_cap, int step_cap) {
  constexpr int NST = staged_values<Env>();
  size_t n = sim_sched_bytes<T>(step_cap, save_cap) + (size_t)(act_cap * Env::A + save_cap * NST) * SIM_EP * sizeof(T);
  if (Env::IS_PMSM && (plan.desc.flags & EXENV_FLAG_SATURATED)) {
    n = (n + 15) & ~(size_t)15;
    n += plan.lut.bytes + 16;
  }
  return n;
}

template <class Env, int SOLVER, typename T, bool PER_ENV>
static int launch_sim_t(const ExenvPlan &plan, const ExenvSimArgs &x, const SimSchedule &sch, cudaStream_t stream) {
  SimK<T> a{};
  a.B = x.batch;
  a.n_actions = x.n_actions;
  a.n_save = sch.n_save;
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
  a.words = sch.d_words;
  a.dts = (const T *)sch.d_dt;
  a.recs = (const SaveRec<T> *)sch.d_theta;
  a.tiles = sch.d_tiles;
  a.n_tiles = sch.n_tiles;
  a.act_cap = sch.act_cap;
  a.save_cap = sch.save_cap;
  a.step_cap = sch.tile_steps;
  a.n_control = plan.desc.n_control;
  for (int j = 0; j < EXENV_MAX_CONTROL; ++j) a.control[j] = plan.desc.control[j];
  a.deadtime = plan.deadtime;
  a.pmsm_adv = (T)plan.pmsm_adv;
  a.pmsm_lin_stop = (T)(plan.desc.tau * (double)(x.n_actions - 1));
  a.lut_blob = plan.lut.blob;
  a.lut = make_lut_geom<T>(plan);
  KParams<T> kp = make_kparams<T>(plan);
  const size_t smem = smem_bytes_t<Env, T>(plan, sch.act_cap, sch.save_cap, sch.tile_steps);
  auto kern = sim_kernel<Env, SOLVER, T, PER_ENV>;
  if (smem > 48 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("sim_ahead needs %zu bytes of shared memory per CTA: %s", smem, cudaGetErrorString(e));
      return EXENV_ERR_CUDA;
    }
  }
  const unsigned grid = (unsigned)((x.batch + SIM_E - 1) / SIM_E);
  kern<<<grid, SIM_E, smem, stream>>>(a, kp);
  g_launch_count++;
  const cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("exenv_sim_ahead launch failed: %s", cudaGetErrorString(e));
    return EXENV_ERR_CUDA;
  }
  return EXENV_OK;
}

template <class Env, typename T>
static int sim_dispatch2(const ExenvPlan &plan, const ExenvSimArgs &x, const SimSchedule &s, cudaStream_t st) {
  const bool tsit = plan.desc.solver == EXENV_TSIT5;
  if constexpr (Env::IS_PMSM) {
    return tsit ? launch_sim_t<Env, EXENV_TSIT5, T, false>(plan, x, s, st)
                : launch_sim_t<Env, EXENV_EULER, T, false>(plan, x, s, st);
  } else {
    if (plan.per_env)
      return tsit ? launch_sim_t<Env, EXENV_TSIT5, T, true>(plan, x, s, st)
                  : launch_sim_t<Env, EXENV_EULER, T, true>(plan, x, s, st);
    return tsit ? launch_sim_t<Env, EXENV_TSIT5, T, false>(plan, x, s, st)
                : launch_sim_t<Env, EXENV_EULER, T, false>(plan, x, s, st);
  }
}

template <typename T>
static int sim_dispatch(const ExenvPlan &plan, const ExenvSimArgs &x, const SimSchedule &s, cudaStream_t st) {
  switch (plan.desc.env) {
    case EXENV_PENDULUM: return sim_dispatch2<Pendulum<T>, T>(plan, x, s, st);
    case EXENV_MASS_SPRING_DAMPER: return sim_dispatch2<MassSpringDamper<T>, T>(plan, x, s, st);
    case EXENV_FLUID_TANK: return sim_dispatch2<FluidTank<T>, T>(plan, x, s, st);
    case EXENV_CART_POLE: return sim_dispatch2<CartPole<T>, T>(plan, x, s, st);
    case EXENV_ACROBOT: return sim_dispatch2<Acrobot<T>, T>(plan, x, s, st);
    case EXENV_PMSM:
      if (plan.desc.flags & EXENV_FLAG_SATURATED) return sim_dispatch2<Pmsm<T, true>, T>(plan, x, s, st);
      return sim_dispatch2<Pmsm<T, false>, T>(plan, x, s, st);
  }
  set_error("unknown env %d", plan.desc.env);
  return EXENV_ERR_INVALID;
}

int EXENV_CAT(launch_sim_, EXENV_SFX)(const ExenvPlan &plan, const ExenvSimArgs &a, const SimSchedule &sch,
                                       cudaStream_t stream) {
  return sim_dispatch<EXENV_REAL>(plan, a, sch, stream);
}

template <typename T>
static size_t smem_dispatch(const ExenvPlan &plan, int ac, int sc, int stc) {
  switch (plan.desc.env) {
    case EXENV_PENDULUM: return smem_bytes_t<Pendulum<T>, T>(plan, ac, sc, stc);
    case EXENV_MASS_SPRING_DAMPER: return smem_bytes_t<MassSpringDamper<T>, T>(plan, ac, sc, stc);
    case EXENV_FLUID_TANK: return smem_bytes_t<FluidTank<T>, T>(plan, ac, sc, stc);
    case EXENV_CART_POLE: return smem_bytes_t<CartPole<T>, T>(plan, ac, sc, stc);
    case EXENV_ACROBOT: return smem_bytes_t<Acrobot<T>, T>(plan, ac, sc, stc);
    default: return smem_bytes_t<Pmsm<T, true>, T>(plan, ac, sc, stc);
  }
}

size_t EXENV_CAT(sim_smem_bytes_, EXENV_SFX)(const ExenvPlan &plan, int step_cap, int act_cap, int save_cap) {
  return smem_dispatch<EXENV_REAL>(plan, act_cap, save_cap, step_cap);
}

}  // namespace exenv
