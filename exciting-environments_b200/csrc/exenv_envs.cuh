// exenv_envs.cuh -- per-environment physics: ODE right-hand sides, post-step wrap/clip, observation, reward,
// truncated/terminated, PMSM inverter voltage limit and flux-map look-up.  Pure scalar code (no memory layout
// knowledge): the kernels in exenv_step.cu / exenv_sim.cu own the data movement.
//
// Parameter vector `p` (same slots as ExenvDesc.p, include/exenv_b200.h):
//   [static params..., (min,max) per PhysicalState field..., (min,max) per Action field...]
// Reference citations are relative to /root/reference/exciting_environments/.
#pragma once

#include "exenv_math.cuh"

namespace exenv {

// ---- PMSM flux maps ----------------------------------------------------------------------------------------
// Device layout prepared at plan creation: grids gx[nx] (i_d), gy[ny] (i_q) and cells[(ix*ny + iy)*6 + q] with
// q = L_dd, L_dq, L_qd, L_qq, Psi_d, Psi_q -- the six RegularGridInterpolator objects of pmsm_env.py:352-361 share
// one index/weight computation here.
template <typename T>
struct LutView {
  const T *gx;
  const T *gy;
  const T *cells;
  int nx, ny;
  T inv_dx, inv_dy;  // 1 / grid spacing: only used for the initial cell guess
  T gx0, gy0;        // g[0] of both grids (kernel-argument copies: no read for the arithmetic cell index)
  T dx, dy;          // grid spacing (coefficient form: cell origins are gx0 + i * dx, exact for the grids it is built for)
  // optional (NULL: plain IEEE division): per grid cell four values {g[i], width d = g[i+1] - g[i], RN(1 / d), 0} -- the
  // cell origin and the operands of the correctly rounded reciprocal division of exenv_math.cuh in one 16-byte read
  const T *wx;
  const T *wy;
  int regular;  // both grids are equidistant up to rounding: the initial guess is at most one cell off
  int cs;       // words per grid node in `cells` (>= 6).  The device blob uses 7: with an even stride the 32 lanes of a
                // warp, each in its own random cell, would reach only 16 of the 32 shared-memory banks per read
};

// Where the sections of the device blob prepared at plan creation live (exenv_api.cu), for building a LutView over
// the global copy (step kernel) or a shared-memory copy (sim_ahead kernels).
template <typename T>
struct LutGeom {
  unsigned bytes, off_gy, off_cells, off_wx, off_wy;
  int nx, ny, regular;
  T inv_dx, inv_dy, gx0, gy0, dx, dy;
};
constexpr int LUT_BLOB_CELL_STRIDE = 7;
// Coefficient form of the flux maps (fp32 trajectory kernels, Pmsm<float, 2>): per grid CELL the bilinear interpolant of
// every map as {c0, c_s, c_t, c_st} (value = c0 + s c_s + t (c_t + s c_st), s / t = position inside the cell), six maps
// = six 16-byte reads instead of 24 four-byte corner reads, 18 FMAs instead of 6 weights + 24 products.  28 words per
// cell (24 used): with seven 16-byte chunks per cell the chunk index 7 c + q walks through all eight bank groups.
constexpr int LUT_COEF_STRIDE = 28;
template <typename T>
EX_HD LutView<T> lut_view_at(const void *base, const LutGeom<T> &g) {
  const char *b = (const char *)base;
  LutView<T> L;
  L.gx = (const T *)b;
  L.gy = (const T *)(b + g.off_gy);
  L.cells = (const T *)(b + g.off_cells);
  L.wx = (const T *)(b + g.off_wx);
  L.wy = (const T *)(b + g.off_wy);
  L.nx = g.nx;
  L.ny = g.ny;
  L.regular = g.regular;
  L.inv_dx = g.inv_dx;
  L.inv_dy = g.inv_dy;
  L.gx0 = g.gx0;
  L.gy0 = g.gy0;
  L.dx = g.dx;
  L.dy = g.dy;
  L.cs = LUT_BLOB_CELL_STRIDE;
  return L;
}
// View over a coefficient-form table (device blob plan.lut.coef / its shared-memory copy): `cells` is the table.
template <typename T>
EX_HD LutView<T> lut_view_coef(const void *base, const LutGeom<T> &g) {
  LutView<T> L{};
  L.cells = (const T *)base;
  L.nx = g.nx;
  L.ny = g.ny;
  L.regular = 1;
  L.inv_dx = g.inv_dx;
  L.inv_dy = g.inv_dy;
  L.gx0 = g.gx0;
  L.gy0 = g.gy0;
  L.dx = g.dx;
  L.dy = g.dy;
  L.cs = LUT_COEF_STRIDE;
  return L;
}
// Host: fills the coefficient table from the six maps (map-major, (nx, ny) each, double).  (nx-1) x (ny-1) cells.
inline void lut_coef_fill(const double *maps, int nx, int ny, float *out) {
  for (int i = 0; i + 1 < nx; ++i)
    for (int j = 0; j + 1 < ny; ++j) {
      float *c = out + ((size_t)i * (ny - 1) + j) * LUT_COEF_STRIDE;
      for (int q = 0; q < 6; ++q) {
        const double *m = maps + (size_t)q * nx * ny;
        const double v00 = m[(size_t)i * ny + j], v01 = m[(size_t)i * ny + j + 1];
        const double v10 = m[(size_t)(i + 1) * ny + j], v11 = m[(size_t)(i + 1) * ny + j + 1];
        c[4 * q + 0] = (float)v00;
        c[4 * q + 1] = (float)(v10 - v00);
        c[4 * q + 2] = (float)(v01 - v00);
        c[4 * q + 3] = (float)((v11 - v10) - (v01 - v00));
      }
      for (int k = 24; k < LUT_COEF_STRIDE; ++k) c[k] = 0.0f;
    }
}
// Host: the coefficient form needs cell origins that the kernel can rebuild exactly: g[i] == fl(g[0] + i * d) in fp32.
inline bool lut_grid_exact_f32(const double *g, int n) {
  if (n < 2) return false;
  const float g0 = (float)g[0], d = (float)(g[1] - g[0]);
  if (!(d > 0.0f) || (double)g0 != g[0] || (double)d != g[1] - g[0]) return false;
  for (int i = 0; i < n; ++i)
    if ((double)(g0 + (float)i * d) != g[i] || (double)((float)i * d) != (double)i * (double)d) return false;
  return true;
}

// searchsorted(g, x, side="left") - 1 clipped to [0, n-2]  (jax RegularGridInterpolator._find_indices)
template <typename T>
EX_HD int lut_cell(const T *g, int n, T inv_d, T x, int regular) {
  T gf = (x - g[0]) * inv_d;
  int i = (gf > (T)0) ? ((gf < (T)(n - 2)) ? (int)gf : n - 2) : 0;
  if (regular) {  // the guess is the cell or a neighbour: two predicated corrections, no divergent loop
    i -= (i > 0 && !(g[i] < x)) ? 1 : 0;
    i += (i < n - 2 && g[i + 1] < x) ? 1 : 0;
    return i;
  }
  while (i > 0 && !(g[i] < x)) --i;
  while (i < n - 2 && g[i + 1] < x) ++i;
  return i;
}

template <typename T, int FIRST, int COUNT>
EX_HD void lut_eval(const LutView<T> &L, T id, T iq, T *out) {
  const int i = lut_cell<T>(L.gx, L.nx, L.inv_dx, id, L.regular);
  const int j = lut_cell<T>(L.gy, L.ny, L.inv_dy, iq, L.regular);
  T yi, yj;
  if (L.wx) {
    InvDiv<T> di, dj;
    di.d = L.wx[4 * i + 1];
    di.r = L.wx[4 * i + 2];
    dj.d = L.wy[4 * j + 1];
    dj.r = L.wy[4 * j + 2];
    yi = div_inv<true, T>(id - L.wx[4 * i], di);
    yj = div_inv<true, T>(iq - L.wy[4 * j], dj);
  } else {
    yi = (id - L.gx[i]) / (L.gx[i + 1] - L.gx[i]);
    yj = (iq - L.gy[j]) / (L.gy[j + 1] - L.gy[j]);
  }
  const T w00 = ((T)1 - yi) * ((T)1 - yj);
  const T w01 = ((T)1 - yi) * yj;
  const T w10 = yi * ((T)1 - yj);
  const T w11 = yi * yj;
  const T *c00 = L.cells + ((long)i * L.ny + j) * L.cs;
  const T *c10 = c00 + (long)L.ny * L.cs;
#pragma unroll
  for (int q = 0; q < COUNT; ++q) {
    T v = c00[FIRST + q] * w00;
    v += c00[L.cs + FIRST + q] * w01;
    v += c10[FIRST + q] * w10;
    v += c10[L.cs + FIRST + q] * w11;
    out[q] = v;
  }
}

// The cell of x on an equidistant grid by arithmetic alone (no reads of the grid): floor((x - g0) / d) clipped to
// [0, n - 2].  At a grid node (and within rounding distance of one) this may name the cell on the other side than
// searchsorted(side="left") - 1 does; the interpolant is continuous there, so its value is the same up to rounding.
template <typename T>
EX_HD int lut_cell_regular(T g0, T inv_d, int n, T x) {
  const T gf = (x - g0) * inv_d;
  return (int)min_(max_(gf, (T)0), (T)(n - 2));  // clamp, then truncate: no branch; NaN -> 0 (fmax / fmin drop a NaN)
}

// The look-up of the instruction-bound fp32 kernels: cell by arithmetic (equidistant grids of the device blob), origin +
// reciprocal width in one read per dimension, the four-corner form of the reference.  No branch, no state: 26
// shared-memory reads.  (A per-thread cache of the current cell's bilinear coefficients was measured and dropped: with
// 32 lanes in 32 different cells some lane leaves its cell in almost every evaluation, so the warp issues the refill
// path anyway, and the 32 extra registers cost a resident CTA: 1.04e10 / 1.24e10 env-steps/s with the cache against
// 1.60e10 / 1.74e10 without, full-range / 2 % voltages, profiles/README.md.)
template <typename T, int FIRST, int COUNT>
EX_HD void lut_eval_direct(const LutView<T> &L, T id, T iq, T *out) {
  if (!(L.wx && L.regular)) {  // host harness / irregular grids: the exact-index path
    lut_eval<T, FIRST, COUNT>(L, id, iq, out);
    return;
  }
  const int i = lut_cell_regular<T>(L.gx0, L.inv_dx, L.nx, id);
  const int j = lut_cell_regular<T>(L.gy0, L.inv_dy, L.ny, iq);
  const T s = (id - L.wx[4 * i]) * L.wx[4 * i + 2], t = (iq - L.wy[4 * j]) * L.wy[4 * j + 2];
  const T s1 = (T)1 - s, t1 = (T)1 - t;
  const T w00 = s1 * t1, w01 = s1 * t, w10 = s * t1, w11 = s * t;
  const T *c00 = L.cells + (i * L.ny + j) * L.cs;
  const T *c10 = c00 + L.ny * L.cs;
#pragma unroll
  for (int q = 0; q < COUNT; ++q) {
    T v = c00[FIRST + q] * w00;
    v = fma_(c00[L.cs + FIRST + q], w01, v);
    v = fma_(c10[FIRST + q], w10, v);
    v = fma_(c10[L.cs + FIRST + q], w11, v);
    out[q] = v;
  }
}
// Coefficient-form look-up (LutView from lut_view_coef).  Cell by arithmetic as above; the position inside the cell from
// the exactly rebuilt cell origin (lut_grid_exact_f32), times the reciprocal spacing; outside the grid the edge cell
// extrapolates linearly, like the four-corner form with weights outside [0, 1] (RegularGridInterpolator, fill_value=None).
template <typename T, int FIRST, int COUNT>
EX_HD void lut_eval_coef(const LutView<T> &L, T id, T iq, T *out) {
  const T fi = min_(max_((id - L.gx0) * L.inv_dx, (T)0), (T)(L.nx - 2));
  const T fj = min_(max_((iq - L.gy0) * L.inv_dy, (T)0), (T)(L.ny - 2));
  const int i = (int)fi, j = (int)fj;  // truncation; NaN -> 0 (fmax / fmin drop a NaN)
  const T s = (id - fma_((T)i, L.dx, L.gx0)) * L.inv_dx;
  const T t = (iq - fma_((T)j, L.dy, L.gy0)) * L.inv_dy;
  const T *c = L.cells + (i * (L.ny - 1) + j) * LUT_COEF_STRIDE + 4 * FIRST;
#if defined(__CUDA_ARCH__)
  // all COUNT reads are issued before the first result is used: one shared-memory latency per look-up.  Left to itself
  // ptxas recycles one or two register quads and pays the latency per map -- a third of the kernel's stall samples
  // (profiles/README.md).  The warp-level barrier is the fence it does not move shared-memory reads across; it orders
  // nothing the kernel needs ordered.  All 32 lanes arrive: the trajectory kernel that instantiates Pmsm<float, 2> keeps
  // every lane of a running warp in its loops (surplus lanes shadow the last environment).
  float4 k[COUNT];
#pragma unroll
  for (int q = 0; q < COUNT; ++q) k[q] = *reinterpret_cast<const float4 *>(c + 4 * q);
  __syncwarp();
#pragma unroll
  for (int q = 0; q < COUNT; ++q) out[q] = fma_(t, fma_(s, (T)k[q].w, (T)k[q].z), fma_(s, (T)k[q].y, (T)k[q].x));
#else
#pragma unroll
  for (int q = 0; q < COUNT; ++q)
    out[q] = fma_(t, fma_(s, c[4 * q + 3], c[4 * q + 2]), fma_(s, c[4 * q + 1], c[4 * q]));
#endif
}
struct NoAux {};

template <typename T>
struct PmsmAux {
  T omega_el;  // constant over a step / a sim_ahead call (pmsm_env.py:546, 639)
  int deadtime;
  LutView<T> lut;
};

// ---- Pendulum (pendulum/pendulum_env.py) -------------------------------------------------------------------
template <typename T>
struct Pendulum {
  static constexpr int ID = 0, D = 2, NODE = 2, A = 1, O = 2, NSTATIC = 3, NTRUNC_BASE = 2;
  static constexpr bool IS_PMSM = false, SATURATED = false;
  using Aux = NoAux;
  struct Pre {
    T lmg;
    InvDiv<T> ml2;
  };
  EX_HD static Pre precompute(const T *p) {
    const T g = p[0], l = p[1], m = p[2];
    Pre q;
    q.lmg = l * m * g;
    q.ml2 = make_invdiv<T>(m * (l * l));
    return q;
  }
  // :144-150  p = g, l, m
  template <bool FAST>
  EX_HD static void rhs(const T *y, const T *u, const T *, const Pre &q, const Aux &, T *dy) {
    dy[0] = y[1];
    if constexpr (FAST) dy[1] = div_mul<T>(u[0] + q.lmg * Mth<T>::sin(y[0]), q.ml2);
    else dy[1] = div_inv<false, T>(u[0] + q.lmg * Mth<T>::sin(y[0]), q.ml2);
  }
  EX_HD static void post(T *y) { y[0] = wrap_pi<T>(y[0]); }  // :188, :241
  template <typename U>
  using rebind = Pendulum<U>;
  EX_HD static bool is_angle(int f) { return f == 0; }
  EX_HD static void default_norm_state(T *z) {  // :265-268
    z[0] = (T)1;
    z[1] = (T)0;
  }
};

// ---- Mass-spring-damper (mass_spring_damper/mass_spring_damper_env.py) --------------------------------------
template <typename T>
struct MassSpringDamper {
  static constexpr int ID = 1, D = 2, NODE = 2, A = 1, O = 2, NSTATIC = 3, NTRUNC_BASE = 2;
  static constexpr bool IS_PMSM = false, SATURATED = false;
  using Aux = NoAux;
  struct Pre {
    InvDiv<T> m;
    T dm, km;  // d / m, k / m
    T a1, a0;  // force / m as one FMA of the normalised action: a1 * z + a0
  };
  EX_HD static Pre precompute(const T *p) {
    Pre q;
    q.m = make_invdiv<T>(p[2]);
    q.dm = p[0] / p[2];
    q.km = p[1] / p[2];
    const T amin = p[NSTATIC + 2 * D], amax = p[NSTATIC + 2 * D + 1];
    q.a1 = (amax - amin) / ((T)2 * p[2]);
    q.a0 = (amax + amin) / ((T)2 * p[2]);
    return q;
  }
  // :142-148  p = d, k, m
  template <bool FAST>
  EX_HD static void rhs(const T *y, const T *u, const T *p, const Pre &q, const Aux &, T *dy) {
    const T d = p[0], k = p[1];
    dy[0] = y[1];
    dy[1] = div_inv<FAST, T>(u[0] - d * y[1] - k * y[0], q.m);
  }
  // The trajectory kernels' form of the same right-hand side (env_sim_prescaled below): the window of zero-order-hold
  // actions holds um = force / m (act_sim: de-normalisation and the division by m folded into one FMA of the
  // normalised action), and dv/dt = um - (d/m) v - (k/m) x is two FMAs with prepared quotients instead of two FMAs and
  // a correctly rounded division.  The acceleration differs from the expression above by a few ulp (the north star's
  // per-step tolerance is 1e-6 relative; tests/test_sim_gpu.py and tests/test_hostcheck.py hold it against the oracle).
  // the same constants for another value type (scalar -> packed: splats that stay broadcast operands)
  template <class SrcPre>
  EX_HD static Pre from_scalar(const SrcPre &s) {
    Pre q;
    q.m.d = T(s.m.d);
    q.m.r = T(s.m.r);
    q.dm = T(s.dm);
    q.km = T(s.km);
    q.a1 = T(s.a1);
    q.a0 = T(s.a0);
    return q;
  }
  EX_HD static T act_sim(T z, const Pre &q) { return fma_(q.a1, z, q.a0); }
  EX_HD static void rhs_sim(const T *y, const T *um, const Pre &q, T *dy) {
    dy[0] = y[1];
    dy[1] = fma_(-q.dm, y[1], fma_(-q.km, y[0], um[0]));
  }
  EX_HD static void post(T *) {}
  EX_HD static bool is_angle(int) { return false; }
  EX_HD static void default_norm_state(T *z) { z[0] = z[1] = (T)0; }  // :261-265
  template <typename U>
  using rebind = MassSpringDamper<U>;
};

// Environments whose right-hand side, post-processing and observation are straight-line +, -, *, fma code can be
// instantiated with the packed type f32x2 (two environments per value, exenv_math.cuh).
template <class Env>
struct env_packable {
  static constexpr bool value = false;
};
template <>
struct env_packable<MassSpringDamper<float>> {
  static constexpr bool value = true;
};
// Environments whose trajectory kernels keep pre-scaled actions in the zero-order-hold window and evaluate
// Env::rhs_sim(y, u_prescaled, pre, dy) instead of Env::rhs<true> (sim_rhs / prepare_action in exenv_sim_core.cuh).
template <class Env>
struct env_sim_prescaled {
  static constexpr bool value = false;
};
template <typename T>
struct env_sim_prescaled<MassSpringDamper<T>> {
  static constexpr bool value = true;
};

// ---- Fluid tank (fluid_tank/fluid_tank_env.py) ---------------------------------------------------------------
template <typename T>
struct FluidTank {
  static constexpr int ID = 2, D = 1, NODE = 1, A = 1, O = 1, NSTATIC = 4, NTRUNC_BASE = 1;
  static constexpr bool IS_PMSM = false, SATURATED = false;
  using Aux = NoAux;
  struct Pre {
    InvDiv<T> Ab;
    T coef, two_g;
  };
  EX_HD static Pre precompute(const T *p) {
    const T Ab = p[0], Ao = p[1], cd = p[2], g = p[3];
    Pre q;
    q.Ab = make_invdiv<T>(Ab);
    q.coef = cd * Ao / Ab;
    q.two_g = (T)2 * g;
    return q;
  }
  // :97-106  p = base_area, orifice_area, c_d, g
  template <bool FAST>
  EX_HD static void rhs(const T *y, const T *u, const T *, const Pre &q, const Aux &, T *dy) {
    const T h = relu_(y[0]);
    if constexpr (FAST) dy[0] = div_mul<T>(u[0], q.Ab) - q.coef * Mth<T>::sqrt(q.two_g * h);
    else dy[0] = div_inv<false, T>(u[0], q.Ab) - q.coef * Mth<T>::sqrt(q.two_g * h);
  }
  EX_HD static void post(T *y) { y[0] = relu_(y[0]); }  // :146, :196
  template <typename U>
  using rebind = FluidTank<U>;
  EX_HD static bool is_angle(int) { return false; }
  EX_HD static void default_norm_state(T *z) { z[0] = (T)0; }  // :221-224
};

// ---- Cart-pole (cart_pole/cart_pole_env.py) ------------------------------------------------------------------
template <typename T>
struct CartPole {
  static constexpr int ID = 3, D = 4, NODE = 4, A = 1, O = 4, NSTATIC = 6, NTRUNC_BASE = 4;
  static constexpr bool IS_PMSM = false, SATURATED = false;
  using Aux = NoAux;
  struct Pre {
    InvDiv<T> M, mpl;
    T k_mup, l43, lmM;  // FAST form: mu_p / (m_p l), 4 l / 3, l m_p / M
  };
  EX_HD static Pre precompute(const T *p) {
    Pre q;
    q.M = make_invdiv<T>(p[4] + p[3]);
    q.mpl = make_invdiv<T>(p[3] * p[2]);
    q.k_mup = p[0] / (p[3] * p[2]);
    q.l43 = p[2] * (T)(4.0 / 3.0);
    q.lmM = p[2] * p[3] / (p[4] + p[3]);
    return q;
  }
  template <class SrcPre>
  EX_HD static Pre from_scalar(const SrcPre &s) {  // scalar -> packed: splats that stay broadcast operands
    Pre q;
    q.M.d = T(s.M.d);
    q.M.r = T(s.M.r);
    q.mpl.d = T(s.mpl.d);
    q.mpl.r = T(s.mpl.r);
    q.k_mup = T(s.k_mup);
    q.l43 = T(s.l43);
    q.lmM = T(s.lmM);
    return q;
  }
  // :159-180  p = mu_p, mu_c, l, m_p, m_c, g ; y = deflection, velocity, theta, omega
  template <bool FAST>
  EX_HD static void rhs(const T *y, const T *u, const T *p, const Pre &q, const Aux &, T *dy) {
    const T mu_p = p[0], mu_c = p[1], l = p[2], m_p = p[3], g = p[5];
    const T v = y[1], th = y[2], om = y[3];
    T s, c;
    Mth<T>::sincos(th, &s, &c);
    const T sg = sign_of<T>(v);
    T d_om, d_v;
    if constexpr (FAST) {  // quotients by parameters as products with the prepared reciprocal, the one state-dependent
                           // division as a Newton-refined reciprocal (exenv_math.cuh: div_mul, rcp_nr)
      // (the denominator l (4/3 - m_p c^2 / M) and the friction quotient with constants prepared in Pre)
      const T ws = (om * om) * s, fr = mu_c * sg;
      d_om = fma_(-q.k_mup, om, fma_(c, div_mul<T>(fma_(-q.mpl.d, ws, fr - u[0]), q.M), g * s)) *
             rcp_nr(fma_(-q.lmM, c * c, q.l43));
      d_v = div_mul<T>(fma_(q.mpl.d, fma_(-d_om, c, ws), u[0] - fr), q.M);
    } else {
      d_om = (g * s + c * div_inv<false, T>(-u[0] - q.mpl.d * (om * om) * s + mu_c * sg, q.M) -
              div_inv<false, T>(mu_p * om, q.mpl)) /
             (l * ((T)(4.0 / 3.0) - div_inv<false, T>(m_p * (c * c), q.M)));
      d_v = div_inv<false, T>(u[0] + q.mpl.d * ((om * om) * s - d_om * c) - mu_c * sg, q.M);
    }
    dy[0] = v;
    dy[1] = d_v;
    dy[2] = om;
    dy[3] = d_om;
  }
  EX_HD static void post(T *y) { y[2] = wrap_pi<T>(y[2]); }  // :229
  template <typename U>
  using rebind = CartPole<U>;
  EX_HD static bool is_angle(int f) { return f == 2; }
  EX_HD static void default_norm_state(T *z) {  // :317-322
    z[0] = z[1] = z[3] = (T)0;
    z[2] = (T)1;
  }
};

// ---- Acrobot (acrobot/acrobot_env.py) ------------------------------------------------------------------------
template <typename T>
struct Acrobot {
  static constexpr int ID = 4, D = 4, NODE = 4, A = 1, O = 4, NSTATIC = 9, NTRUNC_BASE = 4;
  static constexpr bool IS_PMSM = false, SATURATED = false;
  using Aux = NoAux;
  // :171-197  p = g, l_1, l_2, m_1, m_2, l_c1, l_c2, I_1, I_2 ; y = theta_1, theta_2, omega_1, omega_2
  // Parameter-only products of the right-hand side (FAST form): d11 = A0 + 2 c1 cos(theta_2), d12 = A1 + c1 cos(theta_2),
  // d22 = A1, h1 = -c1 sin(theta_2) omega_2 (omega_2 + 2 omega_1), h2 = c1 sin(theta_2) omega_1^2,
  // phi1 = G1 ca + G2 cb, phi2 = G2 cb.
  struct Pre {
    T A0, A1, c1, c1x2, G1, G2;
  };
  EX_HD static Pre precompute(const T *p) {
    const T g = p[0], l1 = p[1], m1 = p[3], m2 = p[4], lc1 = p[5], lc2 = p[6], I1 = p[7], I2 = p[8];
    Pre q;
    q.A0 = m1 * (lc1 * lc1) + m2 * (l1 * l1 + lc2 * lc2) + I1 + I2;
    q.A1 = m2 * (lc2 * lc2) + I2;
    q.c1 = m2 * l1 * lc2;
    q.c1x2 = (T)2 * q.c1;
    q.G1 = (m1 * lc1 + m2 * l1) * g;
    q.G2 = m2 * lc2 * g;
    return q;
  }
  template <class SrcPre>
  EX_HD static Pre from_scalar(const SrcPre &s) {
    Pre q;
    q.A0 = T(s.A0);
    q.A1 = T(s.A1);
    q.c1 = T(s.c1);
    q.c1x2 = T(s.c1x2);
    q.G1 = T(s.G1);
    q.G2 = T(s.G2);
    return q;
  }
  template <typename U>
  using rebind = Acrobot<U>;
  template <bool FAST>
  EX_HD static void rhs(const T *y, const T *u, const T *p, const Pre &q, const Aux &, T *dy) {
    const T th1 = y[0], th2 = y[1], om1 = y[2], om2 = y[3];
    T s2, c2;
    Mth<T>::sincos(th2, &s2, &c2);
    if constexpr (FAST) {
      // * everything that depends on parameters alone prepared once (Pre): ~45 -> ~15 operations for d11 ... phi2;
      // * cos(x + pi/2) = -sin(x): one sincos of theta_1 and the angle-sum identity replace two cosines (the reference's
      //   arguments carry the rounding of x + fl(pi/2), i.e. a few 1e-7 of noise in fp32; the values agree to that level);
      // * three divisions by two state-dependent denominators -> two Newton-refined reciprocals (rcp_nr)
      T s1, c1;
      Mth<T>::sincos(th1, &s1, &c1);
      const T ca = -s1, cb = -(s1 * c2 + c1 * s2);
      const T d11 = fma_(q.c1x2, c2, q.A0), d12 = fma_(q.c1, c2, q.A1), d22 = q.A1;
      const T cs = q.c1 * s2;
      const T h1 = -(cs * om2) * fma_((T)2, om1, om2), h2 = cs * (om1 * om1);
      const T phi2 = q.G2 * cb, phi1 = fma_(q.G1, ca, phi2);
      const T r12 = rcp_nr(d12);
      const T qq = d22 * r12;
      const T d_om1 = rcp_nr(fma_(-qq, d11, d12)) * (fma_(qq, h1 + phi1, u[0]) - h2 - phi2);
      dy[0] = om1;
      dy[1] = om2;
      dy[2] = d_om1;
      dy[3] = -(fma_(d11, d_om1, h1) + phi1) * r12;
      return;
    }
    // the reference's expressions, operation by operation (Euler in fp64: golden-vector replay)
    const T g = p[0], l1 = p[1], m1 = p[3], m2 = p[4], lc1 = p[5], lc2 = p[6], I1 = p[7], I2 = p[8];
    const T d11 = m1 * (lc1 * lc1) + m2 * (l1 * l1 + lc2 * lc2 + (T)2 * l1 * lc2 * c2) + I1 + I2;
    const T d12 = m2 * (lc2 * lc2 + l1 * lc2 * c2) + I2;
    const T d22 = m2 * (lc2 * lc2) + I2;
    const T h1 = -m2 * l1 * lc2 * s2 * (om2 * om2) - (T)2 * m2 * l1 * lc2 * s2 * om1 * om2;
    const T h2 = m2 * l1 * lc2 * s2 * (om1 * om1);
    const T ca = Mth<T>::cos(th1 + Cst<T>::half_pi());
    const T cb = Mth<T>::cos(th1 + th2 + Cst<T>::half_pi());
    const T phi1 = (m1 * lc1 + m2 * l1) * g * ca + m2 * lc2 * g * cb;
    const T phi2 = m2 * lc2 * g * cb;
    const T d_om1 = (T)1 / (d12 - d22 / d12 * d11) * (u[0] + d22 / d12 * (h1 + phi1) - h2 - phi2);
    const T d_om2 = (-d11 * d_om1 - h1 - phi1) / d12;
    dy[0] = om1;
    dy[1] = om2;
    dy[2] = d_om1;
    dy[3] = d_om2;
  }
  EX_HD static void post(T *y) {  // :247-248
    y[0] = wrap_pi<T>(y[0]);
    y[1] = wrap_pi<T>(y[1]);
  }
  EX_HD static bool is_angle(int f) { return f == 0 || f == 1; }
  EX_HD static void default_norm_state(T *z) {  // :338-343
    z[0] = (T)1;
    z[1] = z[2] = z[3] = (T)0;
  }
};

// The transcendental environments: sine / cosine through the packed Cody-Waite + polynomial pair (exenv_math.cuh:
// sincos2), square roots, reciprocal seeds, signs, clamps and the angle wrap per lane, everything else packed.
template <>
struct env_packable<Pendulum<float>> {
  static constexpr bool value = true;
};
template <>
struct env_packable<FluidTank<float>> {
  static constexpr bool value = true;
};
template <>
struct env_packable<CartPole<float>> {
  static constexpr bool value = true;
};
template <>
struct env_packable<Acrobot<float>> {
  static constexpr bool value = true;
};

// ---- PMSM (pmsm/pmsm_env.py) -----------------------------------------------------------------------------------
// PhysicalState fields: 0 u_d_buffer, 1 u_q_buffer, 2 epsilon, 3 i_d, 4 i_q, 5 torque, 6 omega_el (:281-291)
// ODE state y = (i_d, i_q, epsilon) (:555).  p = p, r_s, l_d, l_q, psi_p, u_dc, deadtime.
template <typename T, int SAT>  // 0 linear, 1 saturated (flux-map nodes), 2 saturated, coefficient-form maps (fp32)
struct Pmsm {
  static constexpr int ID = 5, D = 7, NODE = 3, A = 2, O = 8, NSTATIC = 7, NTRUNC_BASE = 1;
  static constexpr bool IS_PMSM = true;
  static constexpr bool SATURATED = SAT != 0;
  static constexpr bool COEF = SAT == 2;
  using Aux = PmsmAux<T>;

  struct Pre {
    InvDiv<T> ld, lq;
  };
  EX_HD static Pre precompute(const T *p) {
    Pre q;
    q.ld = make_invdiv<T>(p[2]);
    q.lq = make_invdiv<T>(p[3]);
    return q;
  }
  template <bool FAST>
  EX_HD static void rhs(const T *y, const T *u, const T *p, const Pre &q_, const Aux &aux, T *dy) {
    const T r_s = p[1];
    const T id = y[0], iq = y[1], w = aux.omega_el;
    if constexpr (!SAT) {  // :509-523 linear_ode
      const T l_d = p[2], l_q = p[3], psi_p = p[4];
      dy[0] = div_inv<FAST, T>(u[0] + w * l_q * iq - r_s * id, q_.ld);
      dy[1] = div_inv<FAST, T>(u[1] - w * (l_d * id + psi_p) - r_s * iq, q_.lq);
    } else if constexpr (COEF) {
      // :487-507 nonlinear_ode with coefficient-form maps.  inv(L) b for b = u - r_s i - omega J Psi by Cramer's rule: the
      // determinant with one FMA (its second product enters exactly), ONE Newton-refined reciprocal, both numerators
      // independent of it -- half the dependent chain of the pivoted elimination below, one MUFU and the pivot vote
      // less (-40 instructions per env-step, +8 %).  The differential-inductance matrices of the flux maps are well
      // conditioned (condition number <= 7 on the shipped motors): against the float64 solve both forms sit at 3e-7 of
      // the terms' scale (tests/test_hostcheck.py::test_flux_map_coefficient_form).
      T q[6];
      lut_eval_coef<T, 0, 6>(aux.lut, id, iq, q);
      const T rhs0 = fma_(w, q[5], fma_(-r_s, id, u[0]));
      const T rhs1 = fma_(-w, q[4], fma_(-r_s, iq, u[1]));
      const T rdet = rcp_nr(fma_(q[0], q[3], -(q[1] * q[2])));
      dy[0] = fma_(rhs0, q[3], -(q[1] * rhs1)) * rdet;
      dy[1] = fma_(q[0], rhs1, -(q[2] * rhs0)) * rdet;
    } else if constexpr (FAST && sizeof(T) == 4) {  // :487-507 nonlinear_ode, instruction-bound fp32 kernels (fp64 keeps
                                                    // the reference's order of operations: its per-step tolerance is 1e-12)
      // Look-up (lut_eval_direct); then, instead of forming inv(L) (LU with partial pivoting + two solves) and three
      // matrix-vector products, ONE right-hand side b = -r_s i + u - omega J Psi is eliminated with the same pivoting
      // (inv(L) b = inv(L) (-r_s i) + inv(L) u + (-inv(L) J Psi) omega up to rounding), its two divisions by pivots as
      // Newton-refined reciprocals (exenv_math.cuh: rcp_nr, ~1 ulp).  The differential inductances are well conditioned
      // (L_dd, L_qq dominate by >10x): the derivative moves by a few ulp, far inside the per-step tolerance.
      T q[6];
      lut_eval_direct<T, 0, 6>(aux.lut, id, iq, q);
      const T rhs0 = fma_(w, q[5], fma_(-r_s, id, u[0]));   // u_d - r_s i_d + omega Psi_q
      const T rhs1 = fma_(-w, q[4], fma_(-r_s, iq, u[1]));  // u_q - r_s i_q - omega Psi_d
      const bool swap = Mth<T>::abs(q[2]) > Mth<T>::abs(q[0]);
      const T b00 = swap ? q[2] : q[0], b01 = swap ? q[3] : q[1];
      const T b10 = swap ? q[0] : q[2], b11 = swap ? q[1] : q[3];
      const T r0 = swap ? rhs1 : rhs0, r1 = swap ? rhs0 : rhs1;
      const T i00 = rcp_nr(b00);
      const T lf = b10 * i00;
      const T x1 = fma_(-lf, r0, r1) * rcp_nr(fma_(-lf, b01, b11));
      dy[0] = fma_(-b01, x1, r0) * i00;
      dy[1] = x1;
    } else {  // :487-507 nonlinear_ode
      T q[6];
      lut_eval<T, 0, 6>(aux.lut, id, iq, q);
      // jnp.linalg.inv of the 2x2 differential-inductance matrix: LU with partial pivoting + two solves
      const bool swap = Mth<T>::abs(q[2]) > Mth<T>::abs(q[0]);
      const T b00 = swap ? q[2] : q[0], b01 = swap ? q[3] : q[1];
      const T b10 = swap ? q[0] : q[2], b11 = swap ? q[1] : q[3];
      const T lf = b10 / b00;
      const T u11 = b11 - lf * b01;
      T inv[2][2];
#pragma unroll
      for (int col = 0; col < 2; ++col) {
        const T e0 = (T)((swap ? 1 : 0) == col), e1 = (T)((swap ? 0 : 1) == col);
        const T y0 = e0, y1 = e1 - lf * y0;
        const T x1 = y1 / u11;
        const T x0 = (y0 - b01 * x1) / b00;
        inv[0][col] = x0;
        inv[1][col] = x1;
      }
      const T jp0 = -q[5], jp1 = q[4];  // J_k @ (Psi_d, Psi_q)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const T d1 = (-inv[r][0] * r_s) * id + (-inv[r][1] * r_s) * iq;
        const T d2 = inv[r][0] * u[0] + inv[r][1] * u[1];
        const T d3 = ((-inv[r][0]) * jp0 + (-inv[r][1]) * jp1) * w;
        dy[r] = d1 + d2 + d3;
      }
    }
    dy[2] = w;
  }
  EX_HD static void post(T *y) { y[2] = wrap_pi<T>(y[2]); }  // :571, :669
  EX_HD static bool is_angle(int) { return false; }

  // :365-381 currents_to_torque / currents_to_torque_saturated
  template <bool FAST = false>
  EX_HD static T torque(T id, T iq, const T *p, const Aux &aux) {
    if constexpr (!SAT) {
      return (T)1.5 * p[0] * (p[4] + (p[2] - p[3]) * id) * iq;
    } else {
      T q[2];
      if constexpr (COEF) lut_eval_coef<T, 4, 2>(aux.lut, id, iq, q);
      else if constexpr (FAST && sizeof(T) == 4) lut_eval_direct<T, 4, 2>(aux.lut, id, iq, q);
      else lut_eval<T, 4, 2>(aux.lut, id, iq, q);
      return (T)1.5 * p[0] * (q[0] * iq - q[1] * id);
    }
  }
};

// :594-616 constraint_denormalization, :82-89 step_eps, :46-72 dq<->alpha/beta, :92-102 apply_hex_constraint.
// ROTATION_MAP is complex64 (:37-43): the rotation constants are float32-rounded in both precisions.
// adv = (deadtime + 0.5) * tau (Python-double product, host side).  a_min/a_max: action normalisation slots.
template <typename T>
EX_HD void pmsm_clip(T a0, T a1, T eps, T omega_el, T u_dc, T adv, const T *an, T *ud_out, T *uq_out) {
  T ud = denorm_mm<T>(a0, an[0], an[1]);
  T uq = denorm_mm<T>(a1, an[2], an[3]);
  const T half_dc = u_dc / (T)2;
  const T sc = (T)1 / half_dc;
  ud = ud * sc;
  uq = uq * sc;
  T ea = eps + adv * omega_el;
  ea = pymod_two_pi<T>(ea);
  if (ea > Cst<T>::pi()) ea = ea + (T)-2 * Cst<T>::pi();
  // t_dq_alpha_beta(-eps) and, further down, t_dq_alpha_beta(eps): one sincos serves both (sin is odd, cos even)
  T sp, cp;
  Mth<T>::sincos(ea, &sp, &cp);
  const T sm = -sp, cm = cp;
  const T ua = cm * ud + sm * uq;
  const T ub = -sm * ud + cm * uq;
  // sector bits i_k = [sin(angle(u) - k * 2 pi / 3) >= 0] (pmsm_env.py:92-96).  sin(angle(u) - phi) has the sign of
  // ub * cos(phi) - ua * sin(phi), so the three bits are half-plane tests on (ua, ub) -- no atan2, no sine.  Only
  // vectors within rounding distance of a sector boundary can land in the neighbouring sector, where the clipped
  // result is the same up to rounding (the limit is continuous across sectors).
  const T r3 = Mth<T>::sqrt((T)3);
  const int i0 = ub >= (T)0;
  const int i1 = ub + r3 * ua <= (T)0;
  const int i2 = r3 * ua - ub >= (T)0;
  const T H = (T)0.5f, S3 = (T)0.8660254037844386f;
  // ROTATION_MAP[i0, i1, i2] (pmsm_env.py:37-43) as selects: neighbouring lanes sit in different sectors, a branch
  // per sector would serialise the warp
  const int code = (i0 << 2) | (i1 << 1) | i2;
  const bool c1 = code == 1, c2 = code == 2, c3 = code == 3, c5 = code == 5, c6 = code == 6;
  const T rr = (c5 || c6) ? H : ((c1 || c2) ? -H : (c3 ? (T)-1 : (T)1));
  const T ri = (c5 || c1) ? S3 : ((c6 || c2) ? -S3 : (T)0);
  T zr = ua * rr - ub * ri, zi = ua * ri + ub * rr;
  const T lim_r = (T)(2.0 / 3.0);
  const T lim_i = (T)(2.0 / 3.0) * r3;
  zr = zr < -lim_r ? -lim_r : (zr > lim_r ? lim_r : zr);
  zi = zi < (T)0 ? (T)0 : (zi > lim_i ? lim_i : zi);
  const T wr = zr * rr - zi * (-ri), wi = zr * (-ri) + zi * rr;
  const T vd = cp * wr + sp * wi;
  const T vq = -sp * wr + cp * wi;
  *ud_out = vd * half_dc;
  *uq_out = vq * half_dc;
}

// ---- counter-based random numbers for the device-side reference generator -----------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11): four 32-bit words per (128-bit counter, 64-bit key).  The GymWrapper reference
// generator (reference gym_wrapper.py:170-192) draws a fresh random initial state per environment whenever a hold
// counter runs out; JAX's threefry bit stream cannot be reproduced without JAX, so parity is statistical: same
// distributions (init_state of each environment), one independent stream per environment (key) and draw (counter).
EX_HD void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t *out) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}
template <typename T>
EX_HD T u01_from_bits(uint32_t x) {  // uniform in [0, 1) with 24 random bits (enough for a reference value)
  return (T)(x >> 8) * (T)(1.0 / 16777216.0);
}

// A random initial physical state as init_state(env_properties, rng) draws it: uniform(-1, 1) per normalised field,
// de-normalised (pendulum_env.py:271-276 and the same lines of the other environments); fluid tank uniform(0, 1)
// (fluid_tank_env.py:227); PMSM: currents uniform in the disc of radius i_max reflected into the permitted box,
// epsilon / omega_el uniform in their ranges, zero buffers, torque from the currents (pmsm_env.py:402-456).
// u[0..7]: uniform [0, 1) numbers.
template <class Env, typename T>
EX_HD void random_physical_state(const T *u, const T *p, const typename Env::Aux &aux, T *s) {
  const T *pn = p + Env::NSTATIC;
  if constexpr (Env::IS_PMSM) {
    const T r = Mth<T>::sqrt(u[0]), phi = Cst<T>::two_pi() * u[1];
    T sp, cp;
    Mth<T>::sincos(phi, &sp, &cp);
    T imax = Mth<T>::abs(pn[6]);
    imax = Mth<T>::abs(pn[7]) > imax ? Mth<T>::abs(pn[7]) : imax;
    imax = Mth<T>::abs(pn[8]) > imax ? Mth<T>::abs(pn[8]) : imax;
    imax = Mth<T>::abs(pn[9]) > imax ? Mth<T>::abs(pn[9]) : imax;
    auto relu = [](T x) { return x > (T)0 ? x : (T)0; };
    const T xd = r * cp * imax, xq = r * sp * imax;
    const T id = xd - (T)2 * relu(xd - pn[7]) + (T)2 * relu(-xd + pn[6]);
    const T iq = xq - (T)2 * relu(xq - pn[9]) + (T)2 * relu(-xq + pn[8]);
    s[0] = (T)0;
    s[1] = (T)0;
    s[2] = denorm_mm<T>((T)2 * u[2] - (T)1, pn[4], pn[5]);
    s[3] = id;
    s[4] = iq;
    s[5] = Env::template torque<false>(id, iq, p, aux);
    s[6] = denorm_mm<T>((T)2 * u[3] - (T)1, pn[12], pn[13]);
  } else {
#pragma unroll
    for (int i = 0; i < Env::D; ++i) {
      const T z = Env::ID == 2 ? u[i] : (T)2 * u[i] - (T)1;
      s[i] = denorm_mm<T>(z, pn[2 * i], pn[2 * i + 1]);
    }
  }
}

// ---- observation / reward / flags ------------------------------------------------------------------------------
// Base observation (without reference entries): e.g. pendulum_env.py:311-329; PMSM pmsm_env.py:898-919 (order
// i_d, i_q, omega_el, torque, cos eps, sin eps, u_d_buffer, u_q_buffer -- not the order of `_obs_description`).
template <class Env, typename T>
EX_HD T obs_entry(int o, const T *s, const T *p) {
  const T *pn = p + Env::NSTATIC;
  if constexpr (Env::IS_PMSM) {
    switch (o) {
      case 0: return norm_mm<T>(s[3], pn[6], pn[7]);
      case 1: return norm_mm<T>(s[4], pn[8], pn[9]);
      case 2: return norm_mm<T>(s[6], pn[12], pn[13]);
      case 3: return norm_mm<T>(s[5], pn[10], pn[11]);
      case 4: return Mth<T>::cos(s[2]);
      case 5: return Mth<T>::sin(s[2]);
      case 6: return norm_mm<T>(s[0], pn[0], pn[1]);
      default: return norm_mm<T>(s[1], pn[2], pn[3]);
    }
  } else {
    T v = s[0], mn = pn[0], mx = pn[1];
#pragma unroll
    for (int i = 1; i < Env::D; ++i)
      if (o == i) {
        v = s[i];
        mn = pn[2 * i];
        mx = pn[2 * i + 1];
      }
    return norm_mm<T>(v, mn, mx);
  }
}

template <class Env, typename T>
EX_HD void observe(const T *s, const T *p, T *obs) {
#pragma unroll
  for (int o = 0; o < Env::O; ++o) {
    if (Env::IS_PMSM && (o == 4 || o == 5)) continue;
    obs[o] = obs_entry<Env, T>(o, s, p);
  }
  if constexpr (Env::IS_PMSM) Mth<T>::sincos(s[2], &obs[5], &obs[4]);
}

// Normalised field `f` of a physical-state-shaped vector (state or reference).
template <class Env, typename T>
EX_HD T norm_field(int f, const T *s, const T *p) {
  const T *pn = p + Env::NSTATIC;
  T v = s[0], mn = pn[0], mx = pn[1];
#pragma unroll
  for (int i = 1; i < Env::D; ++i)
    if (f == i) {
      v = s[i];
      mn = pn[2 * i];
      mx = pn[2 * i + 1];
    }
  return norm_mm<T>(v, mn, mx);
}
template <class Env, typename T>
EX_HD T field_of(int f, const T *s) {
  T v = s[0];
#pragma unroll
  for (int i = 1; i < Env::D; ++i)
    if (f == i) v = s[i];
  return v;
}

// pmsm_env.py:1014-1037 torque_reward_func (i_lim_multiplier = 1, gamma = 0.85), on normalised quantities
template <typename T>
EX_HD T pmsm_torque_reward(T i_d, T i_q, T tq, T tq_ref) {
  const T i_s = Mth<T>::sqrt(i_d * i_d + i_q * i_q);
  const T i_n = (T)1 / (T)1;
  const T i_d_plus = (T)0.2 * i_n;
  const T tol = (T)0.01;
  T rew = (T)0;
  if (i_s > (T)1) rew = (T)-1 * Mth<T>::abs(i_s);
  if (i_s < (T)1.0 && i_s > i_n) rew = (T)0.5 * ((T)1 - (i_s - i_n) / ((T)1 - i_n)) - (T)1;
  if (i_s < i_n && i_d > i_d_plus) rew = (T)-0.5 * ((i_d - i_d_plus) / (i_n - i_d_plus));
  if (i_s < i_n && i_d < i_d_plus && Mth<T>::abs(tq - tq_ref) > tol)
    rew = (T)0.5 * ((T)1 - Mth<T>::abs((tq_ref - tq) / (T)2));
  if (i_s < i_n && i_d < i_d_plus && Mth<T>::abs(tq - tq_ref) < tol) rew = (T)1 - (T)0.5 * i_s;
  return rew * ((T)1 - (T)0.85);
}

// generate_reward (pendulum_env.py:297-309, mass_spring_damper_env.py:295-302, fluid_tank_env.py:252-259,
// cart_pole_env.py:354-366, acrobot_env.py:375-388, pmsm_env.py:985-1008).  control_state = [] -> 0.
template <class Env, typename T>
EX_HD T reward_of(const T *s, const T *ref, const T *p, int n_control, const int *control) {
  T rew = (T)0;
  if constexpr (Env::IS_PMSM) {
    bool has_id = false, has_iq = false, has_tq = false;
    for (int j = 0; j < n_control; ++j) {
      has_id |= control[j] == 3;
      has_iq |= control[j] == 4;
      has_tq |= control[j] == 5;
    }
    const T idn = norm_field<Env, T>(3, s, p), iqn = norm_field<Env, T>(4, s, p);
    if (has_id && has_iq) {
      const T idr = norm_field<Env, T>(3, ref, p), iqr = norm_field<Env, T>(4, ref, p);
      const T mse = (T)0.5 * ((idn - idr) * (idn - idr)) + (T)0.5 * ((iqn - iqr) * (iqn - iqr));
      rew += (T)-1 * (mse * ((T)1 - (T)0.85));
    }
    if (has_tq) rew += pmsm_torque_reward<T>(idn, iqn, norm_field<Env, T>(5, s, p), norm_field<Env, T>(5, ref, p));
  } else {
    for (int j = 0; j < n_control; ++j) {
      const int f = control[j];
      if (Env::is_angle(f)) {
        const T th = field_of<Env, T>(f, s), thr = field_of<Env, T>(f, ref);
        const T a = Mth<T>::sin(th) - Mth<T>::sin(thr);
        const T b = Mth<T>::cos(th) - Mth<T>::cos(thr);
        rew += -(a * a + b * b);
      } else {
        const T e = norm_field<Env, T>(f, s, p) - norm_field<Env, T>(f, ref, p);
        rew += -(e * e);
      }
    }
  }
  return rew;
}

// generate_truncated: |obs| > 1 over the whole observation incl. reference entries (pendulum_env.py:381-385);
// fluid tank: constant [0] (fluid_tank_env.py:323-326); PMSM: [sqrt(i_d_n^2 + i_q_n^2) > 1] (pmsm_env.py:972-979).
template <class Env>
EX_HD constexpr int trunc_dim(int n_control) {
  return (Env::ID == 2 || Env::ID == 5) ? 1 : Env::O + n_control;
}
template <class Env, typename T>
EX_HD bool trunc_entry(int i, const T *s, const T *obs_full, const T *p) {
  if constexpr (Env::ID == 2) {
    return false;
  } else if constexpr (Env::IS_PMSM) {
    // i_s = sqrt(i_d_norm**2 + i_q_norm**2) > 1: un-fused squares and sum (bit-exact flag whatever -fmad says)
    const T idn = norm_field<Env, T>(3, s, p), iqn = norm_field<Env, T>(4, s, p);
    return Mth<T>::sqrt(add_rn(mul_rn(idn, idn), mul_rn(iqn, iqn))) > (T)1;
  } else {
    return Mth<T>::abs(obs_full[i]) > (T)1;
  }
}
// generate_terminated: reward == 0 (pendulum_env.py:387-390); tank False (fluid_tank_env.py:328-331);
// PMSM = truncated (pmsm_env.py:981-983)
template <class Env, typename T>
EX_HD bool terminated_of(T reward, bool trunc0) {
  if constexpr (Env::ID == 2) return false;
  if constexpr (Env::IS_PMSM) return trunc0;
  return reward == (T)0;
}

}  // namespace exenv
