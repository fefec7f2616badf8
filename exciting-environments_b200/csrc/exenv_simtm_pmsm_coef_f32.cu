// Saturated PMSM, fp32, time-major trajectories (Tsit5 and Euler) with coefficient-form flux maps (Pmsm<float, 2>): the same
// simtm_kernel template as exenv_simtm_f32.cu, instantiated for the persistent one-CTA-per-SM flavour
// (exenv_simtm_common.cuh: simtm_is_coef).  Its own translation unit so that it compiles in parallel with the others.
#define EXENV_REAL float
#define EXENV_SFX f32
#define EXENV_SIMTM_COEF_TU 1
#include "exenv_simtm.inl"
