/*
 * ORACLE (test infrastructure, NOT product code).
 *
 * CPU restatement of exciting-environments' batched dynamics path (vmap_step / vmap_sim_ahead /
 * vmap_reset packing) in plain C, instantiated for float and double.  Built by oracle/Makefile into
 * oracle/_build/libexenv_oracle.so and driven from Python through oracle/oracle.py (ctypes).
 *
 * PARITY STATUS
 *   pinned   : Euler `step` for all six envs (PMSM linear, dead time 1) against the reference's own golden
 *              vectors tests/envs/<env>/data/{actions,observations}.npy (copied to tests/golden/reference/) -- see
 *              tests/test_oracle_golden.py.
 *   unpinned : Tsit5 (step and sim_ahead), the diffeqsolve loop emulation, saturated-LUT PMSM, fp32 mode,
 *              reward / truncated / terminated values.  JAX / diffrax are not installable here, the reference has
 *              no tests for these; the third-party algorithms are restated from their published form
 *              (diffrax==0.7.0, jax==0.9.0 as pinned by the reference's pyproject.toml:21-31).
 *
 * Compile WITHOUT floating-point contraction (-ffp-contract=off): the golden vectors behave like
 * un-contracted IEEE arithmetic (fluid tank reproduces bit-exactly).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "exenv_oracle.h"

#define REAL double
#define SFX f64
#define R_FMOD fmod
#define R_SIN sin
#define R_COS cos
#define R_SQRT sqrt
#define R_ATAN2 atan2
#define R_FABS fabs
#include "exenv_oracle_impl.h"
#undef REAL
#undef SFX
#undef R_FMOD
#undef R_SIN
#undef R_COS
#undef R_SQRT
#undef R_ATAN2
#undef R_FABS

#define REAL float
#define SFX f32
#define R_FMOD fmodf
#define R_SIN sinf
#define R_COS cosf
#define R_SQRT sqrtf
#define R_ATAN2 atan2f
#define R_FABS fabsf
#include "exenv_oracle_impl.h"

int exo_abi_version(void) { return 1; }
int exo_desc_size(void) { return (int)sizeof(ExoDesc); }
