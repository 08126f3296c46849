// step_impl.inl -- included by step_f32.cu (REAL=float) and step_f64.cu (REAL=double):
// instantiates the kernels for one precision and defines its launchers.
#include "aux_kernels.cuh"
#include "launch.h"

namespace boatenv {

#define BOAT_CAT2(a, b) a##b
#define BOAT_CAT(a, b) BOAT_CAT2(a, b)
#define FN(name) BOAT_CAT(name, REAL_SUFFIX)

static inline unsigned grid_for(long long n, int block) { return (unsigned)((n + block - 1) / block); }

cudaError_t FN(launch_step_)(const DevCfg &c, const StepArgs &a, cudaStream_t st) {
    const long long n = a.env_end - a.env_begin;
    if (n <= 0) return cudaSuccess;
    const dim3 grid(grid_for(n, kTile)), block(kTile);
    switch (c.wind_kind) {
    case WIND_NONE: boat_step_kernel<REAL, WIND_NONE><<<grid, block, 0, st>>>(c, a); break;
    case WIND_CONST: boat_step_kernel<REAL, WIND_CONST><<<grid, block, 0, st>>>(c, a); break;
    case WIND_VEL_CURVE: boat_step_kernel<REAL, WIND_VEL_CURVE><<<grid, block, 0, st>>>(c, a); break;
    case WIND_ANGLE_RECT: boat_step_kernel<REAL, WIND_ANGLE_RECT><<<grid, block, 0, st>>>(c, a); break;
    case WIND_BOTH: boat_step_kernel<REAL, WIND_BOTH><<<grid, block, 0, st>>>(c, a); break;
    default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

cudaError_t FN(launch_reset_)(const DevCfg &c, const uint8_t *mask, void *obs_out, cudaStream_t st) {
    boat_reset_kernel<REAL><<<grid_for(c.n_envs, kTile), kTile, 0, st>>>(c, mask, reinterpret_cast<REAL *>(obs_out));
    count_launch();
    return cudaGetLastError();
}

cudaError_t FN(launch_get_field_)(const DevCfg &c, int field, void *out, cudaStream_t st) {
    boat_get_field_kernel<REAL><<<grid_for(c.n_envs, 256), 256, 0, st>>>(c, field, out);
    count_launch();
    return cudaGetLastError();
}

cudaError_t FN(launch_set_field_)(const DevCfg &c, int field, const void *in, cudaStream_t st) {
    boat_set_field_kernel<REAL><<<grid_for(c.n_envs, 256), 256, 0, st>>>(c, field, in);
    count_launch();
    return cudaGetLastError();
}

cudaError_t FN(launch_fill_actions_)(const DevCfg &c, unsigned long long step_counter, double scale, void *out,
                                     cudaStream_t st) {
    boat_fill_actions_kernel<REAL><<<grid_for(c.n_envs, 256), 256, 0, st>>>(c, step_counter, scale,
                                                                           reinterpret_cast<REAL *>(out));
    count_launch();
    return cudaGetLastError();
}

#ifdef BOAT_DEFINE_SHARED_LAUNCHERS
cudaError_t launch_wind_table(const DevCfg &c, long long env, double *wv, double *wa, cudaStream_t st) {
    boat_wind_table_kernel<<<1, 32, 0, st>>>(c, env, wv, wa);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_reduce_counters(const double *counters, double *out, cudaStream_t st) {
    boat_reduce_counters_kernel<<<1, 32, 0, st>>>(counters, out);
    count_launch();
    return cudaGetLastError();
}
#endif

}  // namespace boatenv
