#define EXENV_REAL float
#define EXENV_SFX f32
#define EXENV_SIMTM_PACKED 1  // packed fp32x2 arithmetic flavour (exenv_simtm_packed.inl)
#include "exenv_simtm.inl"
