// double-precision instantiations of the warp-per-problem LM kernel
#define REAL double
#include "lm_small_inst.inl"
