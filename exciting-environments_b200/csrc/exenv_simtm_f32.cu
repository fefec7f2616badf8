#define EXENV_REAL float
#define EXENV_SFX f32
#include "exenv_simtm.inl"
