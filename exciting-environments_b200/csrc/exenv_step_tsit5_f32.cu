#define EXENV_REAL float
#define EXENV_SFX f32
#define EXENV_STEP_SOLVER 1  // EXENV_TSIT5
#include "exenv_step.inl"
