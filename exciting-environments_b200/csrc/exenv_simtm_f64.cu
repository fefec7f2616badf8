#define EXENV_REAL double
#define EXENV_SFX f64
#include "exenv_simtm.inl"
