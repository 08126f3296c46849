// fp32 production instantiation of the BoatEnv kernels (FMA contraction on).
#define REAL float
#define REAL_SUFFIX f32
#define BOAT_DEFINE_SHARED_LAUNCHERS
#include "step_impl.inl"
