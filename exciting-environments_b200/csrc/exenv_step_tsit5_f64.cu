#define EXENV_REAL double
#define EXENV_SFX f64
#define EXENV_STEP_SOLVER 1  // EXENV_TSIT5
#include "exenv_step.inl"
