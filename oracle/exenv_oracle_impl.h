/*
 * ORACLE (test infrastructure, NOT product code).
 *
 * Type-generic body of the CPU restatement of the exciting-environments batched dynamics path.
 * Included twice by exenv_oracle.c (REAL=float, SFX=f32 / REAL=double, SFX=f64).
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/exciting_environments/ unless noted).  Third-party arithmetic that is NOT in
 * /root/reference (diffrax==0.7.0 Euler/Tsit5/diffeqsolve, jax==0.9.0 linspace / remainder /
 * RegularGridInterpolator) is restated from its published algorithm; see oracle/README.md.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this.
 */

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)

/* ------------------------------------------------------------------------------------------------
 * Scalar helpers
 * ---------------------------------------------------------------------------------------------- */

/* utils.py:13-17  MinMaxNormalization.normalize / denormalize */
static inline REAL FN(norm)(REAL x, REAL mn, REAL mx) { return (REAL)2 * (x - mn) / (mx - mn) - (REAL)1; }
static inline REAL FN(denorm)(REAL z, REAL mn, REAL mx) { return (z + (REAL)1) / (REAL)2 * (mx - mn) + mn; }

/* jnp `%` (jax.numpy.remainder): lax.rem (C fmod) + sign correction. SURVEY Appendix A header. */
static inline REAL FN(pymod)(REAL x, REAL y) {
  REAL r = R_FMOD(x, y);
  if (r != (REAL)0 && ((r < (REAL)0) != (y < (REAL)0))) r += y;
  return r;
}

#define R_PI ((REAL)3.141592653589793)
#define R_TWO_PI ((REAL)(2 * 3.141592653589793))
#define R_HALF_PI ((REAL)(3.141592653589793 / 2))

/* pendulum/pendulum_env.py:188  theta = ((theta + pi) % (2 pi)) - pi */
static inline REAL FN(wrap_pi)(REAL th) { return FN(pymod)(th + R_PI, R_TWO_PI) - R_PI; }

static inline REAL FN(sgn)(REAL v) { return (REAL)((v > (REAL)0) - (v < (REAL)0)); }

/* Per-env parameter fetch: shared scalar or (B,) array (core_env.py:253-277 in_axes 0 / None). */
typedef struct {
  REAL p[EXO_MAX_PARAMS];
} FN(Params);

static void FN(load_params)(const ExoDesc *d, long b, FN(Params) * out) {
  for (int i = 0; i < EXO_MAX_PARAMS; ++i) {
    if (d->pp[i])
      out->p[i] = ((const REAL *)d->pp[i])[b];
    else
      out->p[i] = (REAL)d->p[i];
  }
}

/* ------------------------------------------------------------------------------------------------
 * PMSM look-up tables.  pmsm/pmsm_env.py:316-363 builds six RegularGridInterpolator objects over
 * (x = i_d grid, y = i_q grid) with values[x_idx][y_idx]; jax.scipy.interpolate.RegularGridInterpolator
 * (method="linear", bounds_error=False, fill_value=None) is restated here.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const REAL *gx, *gy; /* grids */
  int nx, ny;
  const REAL *maps[6]; /* L_dd, L_dq, L_qd, L_qq, Psi_d, Psi_q; each [nx][ny] */
} FN(Lut);

static inline int FN(find_cell)(const REAL *g, int n, REAL x) {
  /* searchsorted(g, x, side="left") - 1, clipped to [0, n-2] */
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (g[mid] < x)
      lo = mid + 1;
    else
      hi = mid;
  }
  int i = lo - 1;
  if (i < 0) i = 0;
  if (i > n - 2) i = n - 2;
  return i;
}

static void FN(lut_eval)(const FN(Lut) * L, REAL id, REAL iq, int first, int count, REAL *out) {
  int i = FN(find_cell)(L->gx, L->nx, id);
  int j = FN(find_cell)(L->gy, L->ny, iq);
  REAL yi = (id - L->gx[i]) / (L->gx[i + 1] - L->gx[i]);
  REAL yj = (iq - L->gy[j]) / (L->gy[j + 1] - L->gy[j]);
  /* edges in itertools.product order: (i,j), (i,j+1), (i+1,j), (i+1,j+1); weight = (1*wx)*wy */
  REAL w00 = ((REAL)1 * ((REAL)1 - yi)) * ((REAL)1 - yj);
  REAL w01 = ((REAL)1 * ((REAL)1 - yi)) * yj;
  REAL w10 = ((REAL)1 * yi) * ((REAL)1 - yj);
  REAL w11 = ((REAL)1 * yi) * yj;
  for (int q = first; q < first + count; ++q) {
    const REAL *V = L->maps[q];
    REAL v = (REAL)0;
    v += V[(long)i * L->ny + j] * w00;
    v += V[(long)i * L->ny + j + 1] * w01;
    v += V[(long)(i + 1) * L->ny + j] * w10;
    v += V[(long)(i + 1) * L->ny + j + 1] * w11;
    out[q - first] = v;
  }
}

/* ------------------------------------------------------------------------------------------------
 * Right-hand sides.  y / dy have NODE entries, u has A entries, p is the per-env parameter vector
 * laid out as [static params..., (min,max) per physical field..., (min,max) per action field...].
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const ExoDesc *d;
  FN(Params) P;
  FN(Lut) lut;
  REAL omega_el; /* PMSM: constant during a step / sim_ahead (pmsm_env.py:546, 639) */
} FN(Ctx);

static void FN(rhs)(const FN(Ctx) * c, const REAL *y, const REAL *u, REAL *dy) {
  const REAL *p = c->P.p;
  switch (c->d->env) {
  case EXO_PENDULUM: { /* pendulum/pendulum_env.py:144-150; p = g,l,m */
    REAL g = p[0], l = p[1], m = p[2];
    dy[0] = y[1];
    dy[1] = (u[0] + l * m * g * R_SIN(y[0])) / (m * (l * l));
  } break;
  case EXO_MSD: { /* mass_spring_damper/mass_spring_damper_env.py:142-148; p = d,k,m */
    REAL dd = p[0], k = p[1], m = p[2];
    dy[0] = y[1];
    dy[1] = (u[0] - dd * y[1] - k * y[0]) / m;
  } break;
  case EXO_FLUID_TANK: { /* fluid_tank/fluid_tank_env.py:97-106; p = base_area, orifice_area, c_d, g */
    REAL Ab = p[0], Ao = p[1], cd = p[2], g = p[3];
    REAL h = y[0] < (REAL)0 ? (REAL)0 : y[0]; /* jnp.clip(h, 0) */
    dy[0] = u[0] / Ab - cd * Ao / Ab * R_SQRT((REAL)2 * g * h);
  } break;
  case EXO_CART_POLE: { /* cart_pole/cart_pole_env.py:159-180; p = mu_p, mu_c, l, m_p, m_c, g */
    REAL mu_p = p[0], mu_c = p[1], l = p[2], m_p = p[3], m_c = p[4], g = p[5];
    REAL v = y[1], th = y[2], om = y[3];
    REAL s = R_SIN(th), co = R_COS(th);
    REAL d_om = (g * s + co * ((-u[0] - m_p * l * (om * om) * s + mu_c * FN(sgn)(v)) / (m_c + m_p)) -
                 (mu_p * om) / (m_p * l)) /
                (l * ((REAL)(4.0 / 3.0) - (m_p * (co * co)) / (m_c + m_p)));
    REAL d_v = (u[0] + m_p * l * ((om * om) * s - d_om * co) - mu_c * FN(sgn)(v)) / (m_c + m_p);
    dy[0] = v;
    dy[1] = d_v;
    dy[2] = om;
    dy[3] = d_om;
  } break;
  case EXO_ACROBOT: { /* acrobot/acrobot_env.py:171-197; p = g,l_1,l_2,m_1,m_2,l_c1,l_c2,I_1,I_2 */
    REAL g = p[0], l1 = p[1], m1 = p[3], m2 = p[4], lc1 = p[5], lc2 = p[6], I1 = p[7], I2 = p[8];
    REAL th1 = y[0], th2 = y[1], om1 = y[2], om2 = y[3];
    REAL c2 = R_COS(th2), s2 = R_SIN(th2);
    REAL d11 = m1 * (lc1 * lc1) + m2 * (l1 * l1 + lc2 * lc2 + (REAL)2 * l1 * lc2 * c2) + I1 + I2;
    REAL d12 = m2 * (lc2 * lc2 + l1 * lc2 * c2) + I2;
    REAL d22 = m2 * (lc2 * lc2) + I2;
    REAL h1 = -m2 * l1 * lc2 * s2 * (om2 * om2) - (REAL)2 * m2 * l1 * lc2 * s2 * om1 * om2;
    REAL h2 = m2 * l1 * lc2 * s2 * (om1 * om1);
    REAL ca = R_COS(th1 + R_HALF_PI);
    REAL cb = R_COS(th1 + th2 + R_HALF_PI);
    REAL phi1 = (m1 * lc1 + m2 * l1) * g * ca + m2 * lc2 * g * cb;
    REAL phi2 = m2 * lc2 * g * cb;
    REAL d_om1 = (REAL)1 / (d12 - d22 / d12 * d11) * (u[0] + d22 / d12 * (h1 + phi1) - h2 - phi2);
    REAL d_om2 = (-d11 * d_om1 - h1 - phi1) / d12;
    dy[0] = om1;
    dy[1] = om2;
    dy[2] = d_om1;
    dy[3] = d_om2;
  } break;
  case EXO_PMSM: { /* p = p, r_s, l_d, l_q, psi_p, u_dc, deadtime */
    REAL r_s = p[1];
    REAL id = y[0], iq = y[1], w = c->omega_el;
    if (!c->d->saturated) { /* pmsm/pmsm_env.py:509-523 linear_ode */
      REAL l_d = p[2], l_q = p[3], psi_p = p[4];
      dy[0] = (u[0] + w * l_q * iq - r_s * id) / l_d;
      dy[1] = (u[1] - w * (l_d * id + psi_p) - r_s * iq) / l_q;
    } else { /* pmsm/pmsm_env.py:487-507 nonlinear_ode */
      REAL q[6];
      FN(lut_eval)(&c->lut, id, iq, 0, 6, q);
      REAL a00 = q[0], a01 = q[1], a10 = q[2], a11 = q[3];
      /* jnp.linalg.inv == LU (partial pivoting) + two triangular solves against the identity */
      REAL inv[2][2];
      int swap = R_FABS(a10) > R_FABS(a00);
      REAL b00 = swap ? a10 : a00, b01 = swap ? a11 : a01, b10 = swap ? a00 : a10, b11 = swap ? a01 : a11;
      REAL lf = b10 / b00;
      REAL u11 = b11 - lf * b01;
      for (int col = 0; col < 2; ++col) {
        /* right-hand side = P e_col */
        REAL e0 = (REAL)((swap ? 1 : 0) == col), e1 = (REAL)((swap ? 0 : 1) == col);
        REAL y0 = e0, y1 = e1 - lf * y0;
        REAL x1 = y1 / u11;
        REAL x0 = (y0 - b01 * x1) / b00;
        inv[0][col] = x0;
        inv[1][col] = x1;
      }
      REAL psi_d = q[4], psi_q = q[5];
      REAL jp0 = -psi_q, jp1 = psi_d; /* J_k @ psi */
      for (int r = 0; r < 2; ++r) {
        REAL d1 = (-inv[r][0] * r_s) * id + (-inv[r][1] * r_s) * iq;
        REAL d2 = inv[r][0] * u[0] + inv[r][1] * u[1];
        REAL d3 = ((-inv[r][0]) * jp0 + (-inv[r][1]) * jp1) * w;
        dy[r] = d1 + d2 + d3;
      }
    }
    dy[2] = w;
  } break;
  }
}

/* Post-processing after a step / over a saved trajectory:
 * pendulum_env.py:188,241; fluid_tank_env.py:146,196; cart_pole_env.py:229; acrobot_env.py:247-248;
 * pmsm_env.py:571,669. */
static void FN(post)(const ExoDesc *d, REAL *y) {
  switch (d->env) {
  case EXO_PENDULUM: y[0] = FN(wrap_pi)(y[0]); break;
  case EXO_FLUID_TANK: y[0] = y[0] < (REAL)0 ? (REAL)0 : y[0]; break;
  case EXO_CART_POLE: y[2] = FN(wrap_pi)(y[2]); break;
  case EXO_ACROBOT:
    y[0] = FN(wrap_pi)(y[0]);
    y[1] = FN(wrap_pi)(y[1]);
    break;
  case EXO_PMSM: y[2] = FN(wrap_pi)(y[2]); break;
  default: break;
  }
}

/* pmsm_env.py:365-381 torque from currents */
static REAL FN(pmsm_torque)(const FN(Ctx) * c, REAL id, REAL iq) {
  const REAL *p = c->P.p;
  if (!c->d->saturated) return (REAL)1.5 * p[0] * (p[4] + (p[2] - p[3]) * id) * iq;
  REAL q[2];
  FN(lut_eval)(&c->lut, id, iq, 4, 2, q);
  return (REAL)1.5 * p[0] * (q[0] * iq - q[1] * id);
}

/* ------------------------------------------------------------------------------------------------
 * Solvers (diffrax 0.7.0, not vendored in the reference).
 * Tsit5 stage increments are k_i = f_i * dt ("vf_prod"), stage states y0 + sum_j a_ij k_j.
 * ---------------------------------------------------------------------------------------------- */
static const double FN(TS_C)[7] = {0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0};
static const double FN(TS_A)[7][6] = {
    {0},
    {0.161},
    {-0.008480655492356989, 0.335480655492357},
    {2.8971530571054935, -6.359448489975075, 4.3622954328695815},
    {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525},
    {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383},
    {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774}};

/* Tsit5 dense-output weights, diffrax _Tsit5Interpolation.evaluate */
static void FN(tsit5_dense_w)(REAL t, REAL *b) {
  b[0] = (REAL)-1.0530884977290216 * t * (t - (REAL)1.3299890189751412) *
         (t * t - (REAL)1.4364028541716351 * t + (REAL)0.7139816917074209);
  b[1] = (REAL)0.1017 * (t * t) * (t * t - (REAL)2.1966568338249754 * t + (REAL)1.2949852507374631);
  b[2] = (REAL)2.490627285651252793 * (t * t) * (t * t - (REAL)2.38535645472061657 * t + (REAL)1.57803468208092486);
  b[3] = (REAL)-16.54810288924490272 * (t - (REAL)1.21712927295533244) * (t - (REAL)0.61620406037800089) * (t * t);
  b[4] = (REAL)47.37952196281928122 * (t - (REAL)1.203071208372362603) * (t - (REAL)0.658047292653547382) * (t * t);
  b[5] = (REAL)-34.87065786149660974 * (t - (REAL)1.2) * (t - (REAL)0.666666666666666667) * (t * t);
  b[6] = (REAL)2.5 * (t - (REAL)1.0) * (t - (REAL)0.6) * (t * t);
}

/* ZOH action closure: actions[jnp.array(t / action_stepsize, int)] (pendulum_env.py:215-216),
 * index truncated toward zero and clamped by JAX gather semantics. `acts` already de-normalised. */
typedef struct {
  const REAL *acts; /* (n_actions, A) */
  long n_actions;
  int A;
  REAL act_dt;
  int constant; /* step(): lambda t: action */
} FN(Zoh);

static const REAL *FN(zoh)(const FN(Zoh) * z, REAL t) {
  if (z->constant) return z->acts;
  long idx = (long)(t / z->act_dt);
  if (idx < 0) idx = 0;
  if (idx > z->n_actions - 1) idx = z->n_actions - 1;
  return z->acts + idx * z->A;
}

/* One solver step over [t0, t1]. f0 = FSAL derivative at (t0, y0) (Tsit5; evaluated here when
 * have_f0 == 0, as diffrax does on `first_step`).  Outputs y1, f_last (Tsit5: f at (t0+dt, y1)),
 * and the k stages for dense output.  Returns nothing; NODE from the descriptor. */
static void FN(solver_step)(const FN(Ctx) * c, const FN(Zoh) * z, int node, REAL t0, REAL t1, const REAL *y0,
                            int have_f0, const REAL *f0, REAL *y1, REAL *f_last, REAL k[7][4]) {
  REAL dt = t1 - t0;
  REAL f[4];
  if (c->d->solver == EXO_EULER) { /* diffrax.Euler: y1 = y0 + dt * f(t0, y0) */
    FN(rhs)(c, y0, FN(zoh)(z, t0), f);
    for (int i = 0; i < node; ++i) {
      k[0][i] = f[i] * dt;
      y1[i] = y0[i] + k[0][i];
    }
    return;
  }
  REAL ys[4];
  for (int s = 0; s < 7; ++s) {
    if (s == 0) {
      if (have_f0) {
        for (int i = 0; i < node; ++i) f[i] = f0[i];
      } else {
        FN(rhs)(c, y0, FN(zoh)(z, t0), f);
      }
    } else {
      for (int i = 0; i < node; ++i) {
        REAL acc = (REAL)0;
        for (int j = 0; j < s; ++j) acc += (REAL)FN(TS_A)[s][j] * k[j][i];
        ys[i] = y0[i] + acc;
      }
      REAL ti = t0 + (REAL)FN(TS_C)[s] * dt;
      FN(rhs)(c, ys, FN(zoh)(z, ti), f);
    }
    for (int i = 0; i < node; ++i) k[s][i] = f[i] * dt;
    if (s == 6)
      for (int i = 0; i < node; ++i) {
        y1[i] = ys[i]; /* FSAL/SSAL: the solution is the stage-7 state */
        f_last[i] = f[i];
      }
  }
}

/* Dense output at absolute time t in [t0, t1] (diffrax LocalLinearInterpolation / _Tsit5Interpolation,
 * both through linear_rescale). */
static void FN(dense_eval)(const ExoDesc *d, int node, REAL t0, REAL t1, REAL t, const REAL *y0, const REAL *y1,
                           REAL k[7][4], REAL *out) {
  REAL th = (t0 == t1) ? (REAL)0 : (t - t0) / (t1 - t0);
  if (d->solver == EXO_EULER) {
    for (int i = 0; i < node; ++i) out[i] = y0[i] + th * (y1[i] - y0[i]);
    return;
  }
  REAL b[7];
  FN(tsit5_dense_w)(th, b);
  for (int i = 0; i < node; ++i) {
    REAL acc = (REAL)0;
    for (int s = 0; s < 7; ++s) acc += b[s] * k[s][i];
    out[i] = y0[i] + acc;
  }
}

/* ------------------------------------------------------------------------------------------------
 * PMSM inverter voltage limit.  pmsm_env.py:594-616 constraint_denormalization, :82-89 step_eps,
 * :46-72 dq<->alpha/beta, :92-102 apply_hex_constraint, :37-43 ROTATION_MAP (complex64!).
 * `a_norm` = normalised action (2,), eps/omega_el = state used for the angle advance.
 * ---------------------------------------------------------------------------------------------- */
static void FN(pmsm_clip)(const FN(Ctx) * c, const REAL *a_norm, REAL eps, REAL omega_el, REAL *u_out) {
  const ExoDesc *d = c->d;
  const REAL *p = c->P.p;
  const int na0 = d->n_static + 2 * d->n_phys; /* first action-normalisation slot */
  REAL u_dc = p[5];
  REAL ud = FN(denorm)(a_norm[0], p[na0 + 0], p[na0 + 1]);
  REAL uq = FN(denorm)(a_norm[1], p[na0 + 2], p[na0 + 3]);
  REAL sc = (REAL)1 / (u_dc / (REAL)2);
  ud = ud * sc;
  uq = uq * sc;
  /* step_eps(eps, deadtime + 0.5, tau, omega_el): eps += ((deadtime+0.5)*tau) * omega_el  (Python-float product) */
  REAL ea = eps + (REAL)d->pmsm_adv_coeff * omega_el;
  ea = FN(pymod)(ea, R_TWO_PI);
  if (ea > R_PI) ea = ea + (REAL)-2 * R_PI;
  /* dq2albet: q = t_dq_alpha_beta(-eps) */
  REAL cm = R_COS(-ea), sm = R_SIN(-ea);
  REAL ua = cm * ud + sm * uq;
  REAL ub = -sm * ud + cm * uq;
  /* apply_hex_constraint */
  REAL ang = R_ATAN2(ub, ua);
  int i0 = R_SIN(ang - (REAL)(2.0 / 3.0 * 3.141592653589793) * (REAL)0) >= (REAL)0;
  int i1 = R_SIN(ang - (REAL)(2.0 / 3.0 * 3.141592653589793) * (REAL)1) >= (REAL)0;
  int i2 = R_SIN(ang - (REAL)(2.0 / 3.0 * 3.141592653589793) * (REAL)2) >= (REAL)0;
  /* complex64 entries: 0.5f and float32(sqrt(3)/2) = 0.8660253882408142 */
  const REAL H = (REAL)0.5f, S3 = (REAL)0.8660254037844386f;
  REAL rr = (REAL)1, ri = (REAL)0;
  int code = (i0 << 2) | (i1 << 1) | i2;
  switch (code) {
  case 5: rr = H; ri = S3; break;   /* [1,0,1] */
  case 6: rr = H; ri = -S3; break;  /* [1,1,0] */
  case 2: rr = -H; ri = -S3; break; /* [0,1,0] */
  case 3: rr = (REAL)-1; ri = (REAL)0; break; /* [0,1,1] */
  case 1: rr = -H; ri = S3; break;  /* [0,0,1] */
  default: break;
  }
  REAL zr = ua * rr - ub * ri, zi = ua * ri + ub * rr;
  const REAL lim_r = (REAL)(2.0 / 3.0);
  const REAL lim_i = (REAL)(2.0 / 3.0) * R_SQRT((REAL)3);
  zr = zr < -lim_r ? -lim_r : (zr > lim_r ? lim_r : zr);
  zi = zi < (REAL)0 ? (REAL)0 : (zi > lim_i ? lim_i : zi);
  /* multiply by conj(rot) */
  REAL wr = zr * rr - zi * (-ri), wi = zr * (-ri) + zi * rr;
  /* albet2dq: q_inv = t_dq_alpha_beta(eps) */
  REAL cp = R_COS(ea), sp = R_SIN(ea);
  REAL vd = cp * wr + sp * wi;
  REAL vq = -sp * wr + cp * wi;
  u_out[0] = vd * (u_dc / (REAL)2);
  u_out[1] = vq * (u_dc / (REAL)2);
}

/* ------------------------------------------------------------------------------------------------
 * Observation / reward / truncated / terminated.
 * `s` = D physical fields, `ref` = D reference fields (may be NULL -> NaN), obs gets O + n_control.
 * ---------------------------------------------------------------------------------------------- */
static int FN(obs_dim)(const ExoDesc *d) { return d->n_obs + d->n_control; }

static void FN(observe)(const FN(Ctx) * c, const REAL *s, const REAL *ref, REAL *obs) {
  const ExoDesc *d = c->d;
  const REAL *p = c->P.p;
  const int n0 = d->n_static;
#define NRM(field, x) FN(norm)((x), p[n0 + 2 * (field)], p[n0 + 2 * (field) + 1])
  int o = 0;
  if (d->env == EXO_PMSM) { /* pmsm_env.py:898-919; field order u_d_buf,u_q_buf,eps,i_d,i_q,torque,omega_el */
    obs[0] = NRM(3, s[3]);
    obs[1] = NRM(4, s[4]);
    obs[2] = NRM(6, s[6]);
    obs[3] = NRM(5, s[5]);
    obs[4] = R_COS(s[2]);
    obs[5] = R_SIN(s[2]);
    obs[6] = NRM(0, s[0]);
    obs[7] = NRM(1, s[1]);
    o = 8;
  } else { /* e.g. pendulum_env.py:311-329: normalised physical state in field order */
    for (int i = 0; i < d->n_phys; ++i) obs[i] = NRM(i, s[i]);
    o = d->n_phys;
  }
  for (int j = 0; j < d->n_control; ++j) {
    int fld = d->control[j];
    REAL r = ref ? ref[fld] : (REAL)NAN;
    obs[o + j] = NRM(fld, r);
  }
}

static int FN(is_angle_field)(const ExoDesc *d, int fld) {
  if (d->env == EXO_PENDULUM) return fld == 0;
  if (d->env == EXO_CART_POLE) return fld == 2;
  if (d->env == EXO_ACROBOT) return fld == 0 || fld == 1;
  return 0;
}

/* pmsm_env.py:1010-1037 */
static REAL FN(pmsm_torque_reward)(REAL i_d, REAL i_q, REAL tq, REAL tq_ref) {
  REAL i_s = R_SQRT(i_d * i_d + i_q * i_q);
  REAL i_n = (REAL)1 / (REAL)1;
  REAL i_d_plus = (REAL)0.2 * i_n;
  REAL tol = (REAL)0.01;
  REAL rew = (REAL)0;
  if (i_s > (REAL)1) rew = (REAL)-1 * R_FABS(i_s);
  if (i_s < (REAL)1.0 && i_s > i_n) rew = (REAL)0.5 * ((REAL)1 - (i_s - i_n) / ((REAL)1 - i_n)) - (REAL)1;
  if (i_s < i_n && i_d > i_d_plus) rew = (REAL)-0.5 * ((i_d - i_d_plus) / (i_n - i_d_plus));
  if (i_s < i_n && i_d < i_d_plus && R_FABS(tq - tq_ref) > tol)
    rew = (REAL)0.5 * ((REAL)1 - R_FABS((tq_ref - tq) / (REAL)2));
  if (i_s < i_n && i_d < i_d_plus && R_FABS(tq - tq_ref) < tol) rew = (REAL)1 - (REAL)0.5 * i_s;
  return rew * ((REAL)1 - (REAL)0.85);
}

/* generate_reward: pendulum_env.py:297-309, mass_spring_damper_env.py:295-302, fluid_tank_env.py:252-259,
 * cart_pole_env.py:354-366, acrobot_env.py:375-388, pmsm_env.py:985-1008 */
static REAL FN(reward)(const FN(Ctx) * c, const REAL *s, const REAL *ref) {
  const ExoDesc *d = c->d;
  const REAL *p = c->P.p;
  const int n0 = d->n_static;
  REAL rew = (REAL)0;
  if (d->env == EXO_PMSM) {
    int has_id = 0, has_iq = 0, has_tq = 0;
    for (int j = 0; j < d->n_control; ++j) {
      has_id |= d->control[j] == 3;
      has_iq |= d->control[j] == 4;
      has_tq |= d->control[j] == 5;
    }
    REAL idn = NRM(3, s[3]), iqn = NRM(4, s[4]);
    if (has_id && has_iq) {
      REAL idr = NRM(3, ref[3]), iqr = NRM(4, ref[4]);
      REAL mse = (REAL)0.5 * ((idn - idr) * (idn - idr)) + (REAL)0.5 * ((iqn - iqr) * (iqn - iqr));
      rew += (REAL)-1 * (mse * ((REAL)1 - (REAL)0.85));
    }
    if (has_tq) rew += FN(pmsm_torque_reward)(idn, iqn, NRM(5, s[5]), NRM(5, ref[5]));
    return rew;
  }
  for (int j = 0; j < d->n_control; ++j) {
    int fld = d->control[j];
    if (FN(is_angle_field)(d, fld)) {
      REAL a = R_SIN(s[fld]) - R_SIN(ref[fld]);
      REAL b = R_COS(s[fld]) - R_COS(ref[fld]);
      rew += -(a * a + b * b);
    } else {
      REAL e = NRM(fld, s[fld]) - NRM(fld, ref[fld]);
      rew += -(e * e);
    }
  }
  return rew;
}

/* generate_truncated -> writes n_trunc flags; returns n_trunc.
 * simple envs: abs(obs) > 1 over the whole obs vector (pendulum_env.py:381-385);
 * fluid tank: constant [0] (fluid_tank_env.py:323-326); PMSM: [sqrt(i_d_n^2+i_q_n^2) > 1] (pmsm_env.py:972-979) */
static int FN(trunc_dim)(const ExoDesc *d) {
  if (d->env == EXO_FLUID_TANK || d->env == EXO_PMSM) return 1;
  return FN(obs_dim)(d);
}
static void FN(truncated)(const FN(Ctx) * c, const REAL *s, const REAL *obs, uint8_t *out) {
  const ExoDesc *d = c->d;
  const REAL *p = c->P.p;
  const int n0 = d->n_static;
  if (d->env == EXO_FLUID_TANK) {
    out[0] = 0;
  } else if (d->env == EXO_PMSM) {
    REAL idn = NRM(3, s[3]), iqn = NRM(4, s[4]);
    out[0] = R_SQRT(idn * idn + iqn * iqn) > (REAL)1;
  } else {
    int n = FN(obs_dim)(d);
    for (int i = 0; i < n; ++i) out[i] = R_FABS(obs[i]) > (REAL)1;
  }
}
/* generate_terminated: reward == 0 (pendulum_env.py:387-390); tank [False] (fluid_tank_env.py:328-331);
 * PMSM = truncated (pmsm_env.py:981-983) */
static uint8_t FN(terminated)(const ExoDesc *d, REAL reward, uint8_t trunc0) {
  if (d->env == EXO_FLUID_TANK) return 0;
  if (d->env == EXO_PMSM) return trunc0;
  return reward == (REAL)0;
}
#undef NRM

static void FN(ctx_init)(FN(Ctx) * c, const ExoDesc *d, long b) {
  c->d = d;
  FN(load_params)(d, b, &c->P);
  c->omega_el = (REAL)0;
  if (d->env == EXO_PMSM && d->saturated) {
    c->lut.gx = (const REAL *)d->lut_gx;
    c->lut.gy = (const REAL *)d->lut_gy;
    c->lut.nx = d->lut_nx;
    c->lut.ny = d->lut_ny;
    for (int q = 0; q < 6; ++q) c->lut.maps[q] = (const REAL *)d->lut_maps + (long)q * d->lut_nx * d->lut_ny;
  }
}

/* ------------------------------------------------------------------------------------------------
 * vmap_step  (core_env.py:533-569 -> step :393-425 ; PMSM override pmsm_env.py:851-883)
 *
 * state_in / state_out: D pointers to (B,) leaves (may alias).  action: (B, A) normalised.
 * obs: (B, O+n_control).  k_out: NODE pointers to (B,) leaves receiving the Tsit5 FSAL derivative
 * (solver_state output leaf) or NULL.  ref: D pointers to (B,) reference leaves or NULL.
 * reward (B,), truncated (B, n_trunc), terminated (B,) optional (gym_wrapper.py:117-126).
 * Tsit5 `step`: fresh k1 on every call -- parity unpinned.
 * ---------------------------------------------------------------------------------------------- */
int FN(exo_step)(const ExoDesc *d, long B, const REAL *const *state_in, const REAL *action, REAL *const *state_out,
                 REAL *obs, REAL *const *k_out, const REAL *const *ref, REAL *reward, uint8_t *truncated,
                 uint8_t *terminated) {
  const int D = d->n_phys, A = d->n_act, node = d->n_ode, O = FN(obs_dim)(d);
  const int nt = FN(trunc_dim)(d);
  const int na0 = d->n_static + 2 * D;
#pragma omp parallel for schedule(static)
  for (long b = 0; b < B; ++b) {
    FN(Ctx) c;
    FN(ctx_init)(&c, d, b);
    REAL s[8], r[8], u[2], y0[4], y1[4], fl[4], k[7][4];
    for (int i = 0; i < D; ++i) s[i] = state_in[i][b];
    if (ref)
      for (int i = 0; i < D; ++i) r[i] = ref[i][b];
    const REAL *a = action + b * A;
    if (d->env == EXO_PMSM) {
      REAL uc[2];
      FN(pmsm_clip)(&c, a, s[2], s[6], uc);
      if (d->pmsm_deadtime > 0) { /* pmsm_env.py:866-875 */
        u[0] = s[0];
        u[1] = s[1];
        s[0] = uc[0];
        s[1] = uc[1];
      } else {
        u[0] = uc[0];
        u[1] = uc[1];
      }
      c.omega_el = s[6];
      y0[0] = s[3];
      y0[1] = s[4];
      y0[2] = s[2];
    } else {
      for (int i = 0; i < A; ++i) u[i] = FN(denorm)(a[i], c.P.p[na0 + 2 * i], c.P.p[na0 + 2 * i + 1]);
      for (int i = 0; i < node; ++i) y0[i] = s[i];
    }
    FN(Zoh) z = {u, 1, A, (REAL)1, 1};
    /* solver.step(term, 0, tau, y0, args, solver_state, made_jump=False): dt = tau - 0 */
    FN(solver_step)(&c, &z, node, (REAL)0, (REAL)d->tau, y0, 0, NULL, y1, fl, k);
    FN(post)(d, y1);
    if (d->env == EXO_PMSM) {
      s[3] = y1[0];
      s[4] = y1[1];
      s[2] = y1[2];
      s[5] = FN(pmsm_torque)(&c, y1[0], y1[1]);
    } else {
      for (int i = 0; i < node; ++i) s[i] = y1[i];
    }
    REAL ob[16];
    FN(observe)(&c, s, ref ? r : NULL, ob);
    for (int i = 0; i < D; ++i) state_out[i][b] = s[i];
    for (int i = 0; i < O; ++i) obs[b * O + i] = ob[i];
    if (k_out && d->solver == EXO_TSIT5)
      for (int i = 0; i < node; ++i) k_out[i][b] = fl[i];
    if (reward || truncated || terminated) {
      REAL rw = (d->n_control && ref) ? FN(reward)(&c, s, r) : (REAL)0;
      uint8_t tr[16];
      FN(truncated)(&c, s, ob, tr);
      if (reward) reward[b] = rw;
      if (truncated)
        for (int i = 0; i < nt; ++i) truncated[b * nt + i] = tr[i];
      if (terminated) terminated[b] = FN(terminated)(d, rw, tr[0]);
    }
  }
  return 0;
}

/* _clip_to_end of diffrax/_integrate.py (keep_step == True for ConstantStepSize) */
static inline REAL FN(clip_to_end)(REAL tnext, REAL t1) {
  const REAL tol = (sizeof(REAL) == 8) ? (REAL)1e-10 : (REAL)1e-6;
  return (tnext > t1 - tol) ? t1 : tnext;
}

/* ------------------------------------------------------------------------------------------------
 * vmap_sim_ahead  (core_env.py:571-616 -> sim_ahead :427-488 -> _ode_solver_simulate_ahead, e.g.
 * pendulum_env.py:196-259; PMSM: pmsm_env.py:746-801, :709-744, :618-707) on top of a restatement of
 * diffrax.diffeqsolve(ConstantStepSize, SaveAt(ts=linspace)).
 *
 * actions: (B, n_actions, A) normalised.  obs_out: (B, n_save, O').  traj_out: D pointers to (B, n_save).
 * last_out: D pointers to (B,).  k_last: NODE pointers (B,) = solver.init at (t1, y_last) derivative
 * (Tsit5) or NULL.  n_save must equal 1 + int(t1 / obs_dt) (Python double arithmetic; checked).
 * time_rule: 0 -> tnext' = tnext + dt0 (diffrax >= 0.5 ConstantStepSize: controller_state = dt0),
 *            1 -> tnext' = tnext + (tnext - tprev) (older diffrax).  Returns #solver steps or <0 on error.
 * ---------------------------------------------------------------------------------------------- */
long FN(exo_sim_ahead)(const ExoDesc *d, long B, long n_actions, double obs_dt_d, double act_dt_d,
                       const REAL *const *state_in, const REAL *actions, REAL *obs_out, REAL *const *traj_out,
                       REAL *const *last_out, REAL *const *k_last, const REAL *const *ref, long n_save,
                       long max_steps, REAL *reward, uint8_t *truncated, uint8_t *terminated) {
  const int D = d->n_phys, A = d->n_act, node = d->n_ode, O = FN(obs_dim)(d);
  const int nt = FN(trunc_dim)(d);
  const int na0 = d->n_static + 2 * D;
  const double t1_d = act_dt_d * (double)n_actions;       /* t1 = action_stepsize * actions.shape[0] */
  const long N = (long)(t1_d / obs_dt_d);                 /* int(t1 / obs_stepsize) */
  if (n_save != N + 1) return -2;
  const REAL t1 = (REAL)t1_d, dt0 = (REAL)obs_dt_d, act_dt = (REAL)act_dt_d;
  /* jnp.linspace(0, t1, N+1): start*(1-s_i) + stop*s_i with s_i = i/N, endpoint appended exactly */
  REAL *ts = (REAL *)malloc(sizeof(REAL) * (size_t)(N + 1));
  for (long i = 0; i < N; ++i) ts[i] = (REAL)0 * ((REAL)1 - (REAL)i / (REAL)N) + t1 * ((REAL)i / (REAL)N);
  ts[N] = t1;
  long steps_taken = -1;
  int err = 0;
#pragma omp parallel for schedule(static)
  for (long b = 0; b < B; ++b) {
    FN(Ctx) c;
    FN(ctx_init)(&c, d, b);
    REAL s0[8], r[8];
    for (int i = 0; i < D; ++i) s0[i] = state_in[i][b];
    if (ref)
      for (int i = 0; i < D; ++i) r[i] = ref[i][b];
    /* de-normalised (PMSM: clipped + dead-time shifted) action sequence for this env */
    REAL *acts = (REAL *)malloc(sizeof(REAL) * (size_t)(n_actions + 1) * A);
    REAL *buf_traj = NULL; /* PMSM: u_dq buffer leaves over time */
    const REAL *an = actions + b * n_actions * A;
    REAL y[4];
    if (d->env == EXO_PMSM) {
      /* constraint_denormalization_ahead (pmsm_env.py:709-744):
       * eps_k = eps_0 + linspace(0, tau*(L-1), L)[k] * omega_el (np-style linspace in jnp) */
      REAL *clipped = (REAL *)malloc(sizeof(REAL) * (size_t)n_actions * 2);
      const long L = n_actions;
      const REAL stop = (REAL)(d->tau * (double)(L - 1));
      for (long kk = 0; kk < L; ++kk) {
        REAL lin;
        if (L == 1)
          lin = (REAL)0;
        else if (kk == L - 1)
          lin = stop;
        else {
          REAL sfrac = (REAL)kk / (REAL)(L - 1);
          lin = (REAL)0 * ((REAL)1 - sfrac) + stop * sfrac;
        }
        REAL eps_k = s0[2] + lin * s0[6];
        FN(pmsm_clip)(&c, an + kk * 2, eps_k, s0[6], clipped + kk * 2);
      }
      /* actions_dead = vstack([buffer]*deadtime, actions[:L-deadtime])  (pmsm_env.py:765-777) */
      const int dead = d->pmsm_deadtime;
      for (long kk = 0; kk < L; ++kk) {
        if (kk < dead) {
          acts[kk * 2] = s0[0];
          acts[kk * 2 + 1] = s0[1];
        } else {
          acts[kk * 2] = clipped[(kk - dead) * 2];
          acts[kk * 2 + 1] = clipped[(kk - dead) * 2 + 1];
        }
      }
      /* buffers over time: vstack([buffer]*deadtime, actions) ; zeros when deadtime == 0 (pmsm_env.py:785-791) */
      buf_traj = (REAL *)malloc(sizeof(REAL) * (size_t)n_save * 2);
      for (long j = 0; j < n_save; ++j) {
        if (dead == 0) {
          buf_traj[j * 2] = buf_traj[j * 2 + 1] = (REAL)0;
        } else if (j < dead) {
          buf_traj[j * 2] = s0[0];
          buf_traj[j * 2 + 1] = s0[1];
        } else {
          long q = j - dead;
          if (q > L - 1) q = L - 1;
          buf_traj[j * 2] = clipped[q * 2];
          buf_traj[j * 2 + 1] = clipped[q * 2 + 1];
        }
      }
      free(clipped);
      c.omega_el = s0[6];
      y[0] = s0[3];
      y[1] = s0[4];
      y[2] = s0[2];
    } else {
      for (long kk = 0; kk < n_actions; ++kk)
        for (int i = 0; i < A; ++i)
          acts[kk * A + i] = FN(denorm)(an[kk * A + i], c.P.p[na0 + 2 * i], c.P.p[na0 + 2 * i + 1]);
      for (int i = 0; i < node; ++i) y[i] = s0[i];
    }
    FN(Zoh) z = {acts, n_actions, A, act_dt, 0};

    /* ---- diffeqsolve loop ---- */
    REAL tprev = (REAL)0, tnext = FN(clip_to_end)((REAL)0 + dt0, t1);
    REAL f_fsal[4], y1[4], fl[4], k[7][4], ysave[4];
    int have_f = 0;
    long si = 0, nstep = 0;
    REAL *traj = (REAL *)malloc(sizeof(REAL) * (size_t)n_save * node);
    while (tprev < t1) {
      if (nstep >= max_steps) {
        err = 1;
        break;
      }
      FN(solver_step)(&c, &z, node, tprev, tnext, y, have_f, f_fsal, y1, fl, k);
      if (d->solver == EXO_TSIT5) {
        have_f = 1;
        for (int i = 0; i < node; ++i) f_fsal[i] = fl[i];
      }
      /* SaveAt(ts): while ts[i] <= tnext: ys[i] = interpolate(ts[i]) over [tprev, tnext] */
      while (si < n_save && ts[si] <= tnext) {
        FN(dense_eval)(d, node, tprev, tnext, ts[si], y, y1, k, ysave);
        for (int i = 0; i < node; ++i) traj[si * node + i] = ysave[i];
        ++si;
      }
      /* ConstantStepSize.adapt_step_size + bookkeeping */
      REAL ntprev = tnext;
      REAL ntnext = d->time_rule ? tnext + (tnext - tprev) : tnext + dt0;
      if (ntprev > t1) ntprev = t1;
      ntnext = FN(clip_to_end)(ntnext, t1);
      tprev = ntprev;
      tnext = ntnext;
      for (int i = 0; i < node; ++i) y[i] = y1[i];
      ++nstep;
    }
    /* unsaved slots stay +inf in diffrax; cannot happen for a completed solve */
    for (; si < n_save; ++si)
      for (int i = 0; i < node; ++i) traj[si * node + i] = (REAL)INFINITY;
    if (b == 0) steps_taken = nstep;

    /* ---- post-process trajectory, observations, epilogue ---- */
    REAL sfull[8], ob[16], prev_rw = 0;
    (void)prev_rw;
    for (long j = 0; j < n_save; ++j) {
      REAL yy[4];
      for (int i = 0; i < node; ++i) yy[i] = traj[j * node + i];
      FN(post)(d, yy);
      if (d->env == EXO_PMSM) {
        sfull[0] = buf_traj[j * 2];
        sfull[1] = buf_traj[j * 2 + 1];
        sfull[2] = yy[2];
        sfull[3] = yy[0];
        sfull[4] = yy[1];
        sfull[5] = FN(pmsm_torque)(&c, yy[0], yy[1]);
        sfull[6] = s0[6];
      } else {
        for (int i = 0; i < node; ++i) sfull[i] = yy[i];
      }
      FN(observe)(&c, sfull, ref ? r : NULL, ob);
      for (int i = 0; i < D; ++i) traj_out[i][b * n_save + j] = sfull[i];
      for (int i = 0; i < O; ++i) obs_out[(b * n_save + j) * O + i] = ob[i];
      if (j == n_save - 1) {
        for (int i = 0; i < D; ++i) last_out[i][b] = sfull[i];
        if (k_last && d->solver == EXO_TSIT5) {
          /* solver.init(term, t1, t1 + tau, y_last, args): derivative at (t1, y_last) with the ZOH closure */
          REAL yl[4], fi[4];
          for (int i = 0; i < node; ++i) yl[i] = yy[i];
          FN(rhs)(&c, yl, FN(zoh)(&z, t1), fi);
          for (int i = 0; i < node; ++i) k_last[i][b] = fi[i];
        }
      }
      /* generate_rew_trunc_term_ahead (core_env.py:490-531; PMSM pmsm_env.py:803-849):
       * truncated for all n_save states, reward / terminated for states[1:] */
      if (reward || truncated || terminated) {
        uint8_t tr[16];
        FN(truncated)(&c, sfull, ob, tr);
        if (truncated)
          for (int i = 0; i < nt; ++i) truncated[(b * n_save + j) * nt + i] = tr[i];
        if (j >= 1) {
          REAL rw = (d->n_control && ref) ? FN(reward)(&c, sfull, r) : (REAL)0;
          if (reward) reward[b * (n_save - 1) + (j - 1)] = rw;
          if (terminated) terminated[b * (n_save - 1) + (j - 1)] = FN(terminated)(d, rw, tr[0]);
        }
      }
    }
    free(traj);
    free(acts);
    if (buf_traj) free(buf_traj);
  }
  free(ts);
  if (err) return -3; /* diffrax max_steps reached */
  return steps_taken;
}

/* ------------------------------------------------------------------------------------------------
 * vmap_reset state packing (core_env.py:664-687, :361-391; init_state e.g. pendulum_env.py:261-295;
 * PMSM pmsm_env.py:383-485) and generate_state_from_observation (e.g. pendulum_env.py:331-364,
 * pmsm_env.py:921-970).
 * mode 0: default reset.  mode 1: norm_in = (B, D) normalised physical state (random reset after the
 *         host drew it; PMSM rows are already physical, pmsm_env.py:436-456).  mode 2: norm_in = obs (B, O').
 * Outputs: D state leaves, D reference leaves (NaN unless mode 2 with control_state), obs.
 * ---------------------------------------------------------------------------------------------- */
int FN(exo_reset)(const ExoDesc *d, long B, int mode, const REAL *norm_in, REAL *const *state_out,
                  REAL *const *ref_out, REAL *obs) {
  const int D = d->n_phys, O = FN(obs_dim)(d), n0 = d->n_static;
#pragma omp parallel for schedule(static)
  for (long b = 0; b < B; ++b) {
    FN(Ctx) c;
    FN(ctx_init)(&c, d, b);
    const REAL *p = c.P.p;
    REAL s[8], r[8], z[8];
    for (int i = 0; i < D; ++i) r[i] = (REAL)NAN;
    int already_physical = 0;
    if (mode == 0) {
      for (int i = 0; i < D; ++i) z[i] = (REAL)0;
      switch (d->env) {
      case EXO_PENDULUM: z[0] = (REAL)1; break;  /* pendulum_env.py:265-268 */
      case EXO_CART_POLE: z[2] = (REAL)1; break; /* cart_pole_env.py:317-322 */
      case EXO_ACROBOT: z[0] = (REAL)1; break;   /* acrobot_env.py:338-343 */
      case EXO_PMSM:                             /* pmsm_env.py:386-399: physical, no denormalise */
        already_physical = 1;
        z[3] = (p[n0 + 6] + p[n0 + 7]) / (REAL)2;
        z[6] = (p[n0 + 12] + p[n0 + 13]) / (REAL)2;
        break;
      default: break;
      }
    } else if (mode == 1) {
      for (int i = 0; i < D; ++i) z[i] = norm_in[b * D + i];
      already_physical = d->env == EXO_PMSM;
    } else {
      const REAL *ob = norm_in + b * O;
      if (d->env == EXO_PMSM) { /* pmsm_env.py:928-936 */
        z[0] = ob[6];
        z[1] = ob[7];
        z[2] = R_ATAN2(ob[5], ob[4]) / R_PI;
        z[3] = ob[0];
        z[4] = ob[1];
        z[5] = ob[3];
        z[6] = ob[2];
      } else {
        for (int i = 0; i < D; ++i) z[i] = ob[i];
      }
      for (int j = 0; j < d->n_control; ++j) {
        int fld = d->control[j];
        r[fld] = FN(denorm)(ob[d->n_obs + j], p[n0 + 2 * fld], p[n0 + 2 * fld + 1]);
      }
    }
    for (int i = 0; i < D; ++i) s[i] = already_physical ? z[i] : FN(denorm)(z[i], p[n0 + 2 * i], p[n0 + 2 * i + 1]);
    REAL ob2[16];
    FN(observe)(&c, s, r, ob2);
    for (int i = 0; i < D; ++i) {
      state_out[i][b] = s[i];
      if (ref_out) ref_out[i][b] = r[i];
    }
    for (int i = 0; i < O; ++i) obs[b * O + i] = ob2[i];
  }
  return 0;
}

#undef FN
#undef CAT
#undef CAT_
#undef R_PI
#undef R_TWO_PI
#undef R_HALF_PI
