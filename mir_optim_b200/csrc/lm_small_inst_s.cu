// single-precision instantiations of the warp-per-problem LM kernel
#define REAL float
#include "lm_small_inst.inl"
