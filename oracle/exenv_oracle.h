/*
 * ORACLE (test infrastructure, NOT product code) -- descriptor shared by the f32 / f64 instantiations.
 * See exenv_oracle_impl.h for the restated algorithm and its reference citations.
 */
#ifndef EXENV_ORACLE_H
#define EXENV_ORACLE_H

#include <stdint.h>

#define EXO_MAX_PARAMS 32
#define EXO_MAX_CONTROL 8

enum { EXO_PENDULUM = 0, EXO_MSD = 1, EXO_FLUID_TANK = 2, EXO_CART_POLE = 3, EXO_ACROBOT = 4, EXO_PMSM = 5 };
enum { EXO_EULER = 0, EXO_TSIT5 = 1 };

typedef struct ExoDesc {
  int32_t env;
  int32_t solver;
  int32_t n_phys;   /* D: PhysicalState fields */
  int32_t n_ode;    /* ODE dimension */
  int32_t n_act;    /* A */
  int32_t n_obs;    /* O without reference entries */
  int32_t n_static; /* number of static params at the front of p[] */
  int32_t n_control;
  int32_t control[EXO_MAX_CONTROL]; /* indices into PhysicalState fields */
  int32_t saturated;                /* PMSM */
  int32_t pmsm_deadtime;            /* PMSM: static_params.deadtime */
  int32_t time_rule;                /* sim_ahead: 0 tnext+dt0, 1 tnext+(tnext-tprev) */
  int32_t pad_;
  double tau;
  double pmsm_adv_coeff; /* (deadtime + 0.5) * tau, Python-double product (pmsm_env.py:599-604) */
  /* parameters: [static..., (min,max) per physical field..., (min,max) per action]; pp[i] != NULL ->
   * per-env (B,) array in the working dtype (core_env.py:253-277) */
  double p[EXO_MAX_PARAMS];
  const void *pp[EXO_MAX_PARAMS];
  /* PMSM saturated LUT: 6 maps [nx][ny] (values[x_idx][y_idx]), grids in the working dtype */
  const void *lut_maps;
  const void *lut_gx;
  const void *lut_gy;
  int32_t lut_nx, lut_ny;
} ExoDesc;

#endif
