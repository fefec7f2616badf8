#define EXENV_REAL double
#define EXENV_SFX f64
#define EXENV_STEP_SOLVER 0  // EXENV_EULER (+ reset and epilogue kernels)
#include "exenv_step.inl"
