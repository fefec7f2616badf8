#define EXENV_REAL float
#define EXENV_SFX f32
#define EXENV_STEP_SOLVER 0  // EXENV_EULER (+ reset and epilogue kernels)
#include "exenv_step.inl"
