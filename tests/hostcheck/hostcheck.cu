// TEST-ONLY device-math host harness (never loaded by the product package).
//
// The per-environment device functions of the CUDA kernels (csrc/exenv_envs.cuh, exenv_sim_core.cuh,
// exenv_kernel_common.cuh: env_step_one) are __host__ __device__; this file calls their HOST instantiations in plain
// loops so that `pytest -m "not gpu"` can compare the kernels' arithmetic with the oracle on a machine without a GPU.
// It is test infrastructure like oracle/: it shares no code path with `libexenv_b200.so`'s entry points and is built
// into tests/hostcheck/_build/ by tests/conftest.py.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../exciting-environments_b200/csrc/exenv_kernel_common.cuh"
#include "../../exciting-environments_b200/csrc/exenv_sim_core.cuh"

using namespace exenv;

namespace exenv {
void set_error(const char *, ...) {}
std::atomic<uint64_t> g_launch_count{0};
}  // namespace exenv

struct HcLut {
  const void *gx, *gy, *cells;  // plan dtype, cells interleaved [(ix*ny+iy)*6+q]
  int nx, ny;
  double inv_dx, inv_dy;
};

template <typename T>
static LutView<T> view(const HcLut *l) {
  LutView<T> v{};
  if (l && l->cells) {
    v.gx = (const T *)l->gx;
    v.gy = (const T *)l->gy;
    v.cells = (const T *)l->cells;
    v.nx = l->nx;
    v.ny = l->ny;
    v.inv_dx = (T)l->inv_dx;
    v.inv_dy = (T)l->inv_dy;
    v.cs = 6;
  }
  return v;
}

template <class Env, int SOLVER, typename T>
static void step_t(const ExenvDesc *d, const HcLut *lut, long B, const T *const *sin, const T *action, T *const *sout,
                   T *obs, T *const *kout) {
  constexpr int D = Env::D, A = Env::A, NODE = Env::NODE, O = Env::O;
  T p[EXENV_MAX_PARAMS];
  for (long b = 0; b < B; ++b) {
    for (int i = 0; i < EXENV_MAX_PARAMS; ++i) p[i] = d->pp[i] ? ((const T *)d->pp[i])[b] : (T)d->p[i];
    T s[D], fl[NODE], ob[O];
    for (int i = 0; i < D; ++i) s[i] = sin[i][b];
    typename Env::Aux aux{};
    if constexpr (Env::IS_PMSM) aux.lut = view<T>(lut);
    const double dead = Env::IS_PMSM ? d->p[6] : 0.0;
    env_step_one<Env, SOLVER, T>(s, action + b * A, p, aux, (T)d->tau, (T)((dead + 0.5) * d->tau), (int)dead, fl);
    observe<Env, T>(s, p, ob);
    for (int i = 0; i < D; ++i) sout[i][b] = s[i];
    for (int i = 0; i < O; ++i) obs[b * O + i] = ob[i];
    if (kout)
      for (int i = 0; i < NODE; ++i) kout[i][b] = fl[i];
  }
}

template <class Env, int SOLVER, typename T>
static void sim_t(const ExenvDesc *d, const HcLut *lut, long B, long n_actions, long n_save, long n_steps,
                  const uint64_t *words, const double *dts, const double *thetas, const T *const *sin,
                  const T *actions, T *const *traj /* NST leaves (B, n_save) */) {
  constexpr int D = Env::D, A = Env::A, NODE = Env::NODE, NST = staged_values<Env>();
  std::vector<SaveRec<T>> recs((size_t)n_save);
  make_saverecs<T>(SOLVER == EXENV_TSIT5, n_steps, words, dts, thetas, recs);
  const double dead_d = Env::IS_PMSM ? d->p[6] : 0.0;
  const int dead = (int)dead_d;
  for (long b = 0; b < B; ++b) {
    T p[EXENV_MAX_PARAMS];
    for (int i = 0; i < EXENV_MAX_PARAMS; ++i) p[i] = d->pp[i] ? ((const T *)d->pp[i])[b] : (T)d->p[i];
    T s0[D];
    for (int i = 0; i < D; ++i) s0[i] = sin[i][b];
    typename Env::Aux aux{};
    T y[NODE], f[NODE];
    if constexpr (Env::IS_PMSM) {
      aux.lut = view<T>(lut);
      aux.omega_el = s0[6];
      aux.deadtime = dead;
      y[0] = s0[3];
      y[1] = s0[4];
      y[2] = s0[2];
    } else {
      for (int i = 0; i < NODE; ++i) y[i] = s0[i];
    }
    for (int i = 0; i < NODE; ++i) f[i] = 0;
    long save_idx = 0;
    const typename Env::Pre pre = Env::precompute(p);
    const T adv = (T)((dead_d + 0.5) * d->tau), lin_stop = (T)(d->tau * (double)(n_actions - 1));
    auto fetch = [&](int idx, T *u) {
      const long raw = (long)idx - dead;
      const T *src = actions + (b * n_actions + (raw < 0 ? 0 : raw)) * A;
      prepare_action<Env, T>(idx, src, s0, p, pre, dead, n_actions, adv, lin_stop, u);
    };
    if (SOLVER == EXENV_TSIT5) sim_first_derivative<Env, T>(words[0], y, f, p, pre, aux, fetch);
    for (long n = 0; n < n_steps; ++n) {
      sim_one_step<Env, SOLVER, T>(
          words[n], (T)dts[n], y, f, p, pre, aux, recs[(size_t)(save_idx < n_save ? save_idx : n_save - 1)],
          recs.data(), save_idx, fetch,
          [&](long sidx, const T *st) {
            for (int i = 0; i < NST; ++i) traj[i][b * n_save + sidx] = st[i];
          });
    }
  }
}

#define DISPATCH(FUNC, ...)                                                                                     \
  switch (d->env) {                                                                                             \
    case EXENV_PENDULUM: tsit ? FUNC<Pendulum<T>, 1, T>(__VA_ARGS__) : FUNC<Pendulum<T>, 0, T>(__VA_ARGS__); break; \
    case EXENV_MASS_SPRING_DAMPER:                                                                              \
      tsit ? FUNC<MassSpringDamper<T>, 1, T>(__VA_ARGS__) : FUNC<MassSpringDamper<T>, 0, T>(__VA_ARGS__);       \
      break;                                                                                                    \
    case EXENV_FLUID_TANK: tsit ? FUNC<FluidTank<T>, 1, T>(__VA_ARGS__) : FUNC<FluidTank<T>, 0, T>(__VA_ARGS__); break; \
    case EXENV_CART_POLE: tsit ? FUNC<CartPole<T>, 1, T>(__VA_ARGS__) : FUNC<CartPole<T>, 0, T>(__VA_ARGS__); break;   \
    case EXENV_ACROBOT: tsit ? FUNC<Acrobot<T>, 1, T>(__VA_ARGS__) : FUNC<Acrobot<T>, 0, T>(__VA_ARGS__); break;       \
    default:                                                                                                    \
      if (d->flags & EXENV_FLAG_SATURATED)                                                                      \
        tsit ? FUNC<Pmsm<T, true>, 1, T>(__VA_ARGS__) : FUNC<Pmsm<T, true>, 0, T>(__VA_ARGS__);                 \
      else                                                                                                      \
        tsit ? FUNC<Pmsm<T, false>, 1, T>(__VA_ARGS__) : FUNC<Pmsm<T, false>, 0, T>(__VA_ARGS__);               \
  }

template <typename T>
static void step_any(const ExenvDesc *d, const HcLut *lut, long B, const void *const *sin, const void *action,
                     void *const *sout, void *obs, void *const *kout) {
  const bool tsit = d->solver == EXENV_TSIT5;
  DISPATCH(step_t, d, lut, B, (const T *const *)sin, (const T *)action, (T *const *)sout, (T *)obs, (T *const *)kout)
}
template <typename T>
static void sim_any(const ExenvDesc *d, const HcLut *lut, long B, long n_actions, long n_save, long n_steps,
                    const uint64_t *words, const double *dts, const double *thetas, const void *const *sin,
                    const void *actions, void *const *traj) {
  const bool tsit = d->solver == EXENV_TSIT5;
  DISPATCH(sim_t, d, lut, B, n_actions, n_save, n_steps, words, dts, thetas, (const T *const *)sin, (const T *)actions,
           (T *const *)traj)
}

// exenv_math.cuh: the three normalisation forms on arrays -- norm_mm (the reference's expression, IEEE division),
// norm_inv<true> (prepared reciprocal + correction: what the truncation flags are evaluated on) and norm_fma (one FMA:
// what the time-major trajectory kernels return)
template <typename T>
static void norm_forms_t(long n, const T *x, T mn, T mx, T *o_mm, T *o_inv, T *o_fma, T *coef) {
  const InvDiv<T> span = make_invdiv<T>(mx - mn);
  const NormFma<T> nf = make_normfma<T>(mn, mx);
  coef[0] = nf.c1;
  coef[1] = nf.c0;
  for (long i = 0; i < n; ++i) {
    o_mm[i] = norm_mm<T>(x[i], mn, mx);
    o_inv[i] = norm_inv<true, T>(x[i], mn, span);
    o_fma[i] = norm_fma<T>(x[i], nf);
  }
}

extern "C" {
void hc_step(const ExenvDesc *d, const HcLut *lut, long B, const void *const *sin, const void *action,
             void *const *sout, void *obs, void *const *kout) {
  if (d->dtype == EXENV_F64) step_any<double>(d, lut, B, sin, action, sout, obs, kout);
  else step_any<float>(d, lut, B, sin, action, sout, obs, kout);
}
void hc_sim(const ExenvDesc *d, const HcLut *lut, long B, long n_actions, long n_save, long n_steps,
            const uint64_t *words, const double *dts, const double *thetas, const void *const *sin,
            const void *actions, void *const *traj) {
  if (d->dtype == EXENV_F64) sim_any<double>(d, lut, B, n_actions, n_save, n_steps, words, dts, thetas, sin, actions, traj);
  else sim_any<float>(d, lut, B, n_actions, n_save, n_steps, words, dts, thetas, sin, actions, traj);
}
// exenv_math.cuh: fmod_pos (exact C fmod without the iterative reduction) and wrap_pi on arrays
void hc_norm_forms(long n, int is_f64, const void *x, double mn, double mx, void *o_mm, void *o_inv, void *o_fma,
                   void *coef) {
  if (is_f64)
    norm_forms_t<double>(n, (const double *)x, mn, mx, (double *)o_mm, (double *)o_inv, (double *)o_fma, (double *)coef);
  else
    norm_forms_t<float>(n, (const float *)x, (float)mn, (float)mx, (float *)o_mm, (float *)o_inv, (float *)o_fma,
                        (float *)coef);
}
// exenv_math.cuh: the packed sine / cosine pair (two angles per call) and the packed Newton reciprocal
void hc_sincos2(long n, const float *x, float *s, float *c) {
  for (long i = 0; i + 1 < n; i += 2) {
    f32x2 sn, cs;
    sincos2(f32x2(x[i], x[i + 1]), &sn, &cs);
    s[i] = sn.x;
    s[i + 1] = sn.y;
    c[i] = cs.x;
    c[i + 1] = cs.y;
  }
}
// exenv_envs.cuh: the saturated-PMSM look-up in its node form (four corners) and its coefficient form (Pmsm<float, 2>),
// and the right-hand side built on each.  maps: (6, nx, ny) doubles; points (id, iq); out_*: (n, 6); dy_*: (n, 3).
// Returns whether both grids qualify for the coefficient form (lut_grid_exact_f32).
int hc_lut_forms(long n, const double *gx, const double *gy, const double *maps, int nx, int ny, const float *id,
                 const float *iq, const float *u, const float *omega, const float *params, float *out_node,
                 float *out_coef, float *dy_node, float *dy_coef) {
  std::vector<float> gxf(gx, gx + nx), gyf(gy, gy + ny), cells((size_t)nx * ny * 6);
  for (int q = 0; q < 6; ++q)
    for (int i = 0; i < nx; ++i)
      for (int j = 0; j < ny; ++j) cells[((size_t)i * ny + j) * 6 + q] = (float)maps[((size_t)q * nx + i) * ny + j];
  LutView<float> node{};
  node.gx = gxf.data();
  node.gy = gyf.data();
  node.cells = cells.data();
  node.nx = nx;
  node.ny = ny;
  node.inv_dx = (float)(1.0 / (gx[1] - gx[0]));
  node.inv_dy = (float)(1.0 / (gy[1] - gy[0]));
  node.regular = 1;
  node.cs = 6;
  const bool exact = lut_grid_exact_f32(gx, nx) && lut_grid_exact_f32(gy, ny);
  std::vector<float> coef((size_t)(nx - 1) * (ny - 1) * LUT_COEF_STRIDE);
  lut_coef_fill(maps, nx, ny, coef.data());
  LutGeom<float> g{};
  g.nx = nx;
  g.ny = ny;
  g.gx0 = (float)gx[0];
  g.gy0 = (float)gy[0];
  g.dx = (float)(gx[1] - gx[0]);
  g.dy = (float)(gy[1] - gy[0]);
  g.inv_dx = node.inv_dx;
  g.inv_dy = node.inv_dy;
  const LutView<float> cv = lut_view_coef<float>(coef.data(), g);
  for (long k = 0; k < n; ++k) {
    lut_eval<float, 0, 6>(node, id[k], iq[k], out_node + 6 * k);
    lut_eval_coef<float, 0, 6>(cv, id[k], iq[k], out_coef + 6 * k);
    const float y[3] = {id[k], iq[k], 0.0f};
    PmsmAux<float> a1{}, a2{};
    a1.omega_el = a2.omega_el = omega[k];
    a1.lut = node;
    a2.lut = cv;
    Pmsm<float, 1>::rhs<true>(y, u + 2 * k, params, Pmsm<float, 1>::precompute(params), a1, dy_node + 3 * k);
    Pmsm<float, 2>::rhs<true>(y, u + 2 * k, params, Pmsm<float, 2>::precompute(params), a2, dy_coef + 3 * k);
  }
  return exact ? 1 : 0;
}
// exenv_envs.cuh: does a grid qualify for the coefficient form (nodes == fl(g0 + i * d) exactly in float32)?
int hc_lut_grid_exact(const double *g, int n) { return lut_grid_exact_f32(g, n) ? 1 : 0; }
void hc_fmod_two_pi(long n, int is_f64, const void *x, void *fmod_out, void *wrap_out) {
  for (long i = 0; i < n; ++i) {
    if (is_f64) {
      const double y = 2 * 3.141592653589793;
      ((double *)fmod_out)[i] = fmod_pos<double>(((const double *)x)[i], y, 1.0 / y);
      ((double *)wrap_out)[i] = wrap_pi<double>(((const double *)x)[i]);
    } else {
      const float y = (float)(2 * 3.141592653589793);
      ((float *)fmod_out)[i] = fmod_pos<float>(((const float *)x)[i], y, 1.0f / y);
      ((float *)wrap_out)[i] = wrap_pi<float>(((const float *)x)[i]);
    }
  }
}
}
