// fp64 validation instantiation of the BoatEnv kernels: the reference's operation
// order; this file is compiled with -fmad=false so multiplies and adds stay unfused.
#define REAL double
#define REAL_SUFFIX f64
#include "step_impl.inl"
