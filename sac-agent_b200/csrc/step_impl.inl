// step_impl.inl -- included by step_f32.cu (REAL=float) and step_f64.cu (REAL=double):
// instantiates the kernels for one precision and defines its launchers.
#include <algorithm>
#include <mutex>
#include <vector>

#include "aux_kernels.cuh"
#include "launch.h"

namespace boatenv {

#define BOAT_CAT2(a, b) a##b
#define BOAT_CAT(a, b) BOAT_CAT2(a, b)
#define FN(name) BOAT_CAT(name, REAL_SUFFIX)

struct OccEntry { int dev, smem, ctas_per_sm, n_sm; };
constexpr int kMaxDevices = 64;

static inline unsigned grid_for(long long n, int block) { return (unsigned)((n + block - 1) / block); }

// Persistent launch: enough CTAs to fill every SM at the kernel's occupancy (queried once per
// instantiation and shared-memory size), never more than there are 8-warp groups of blocks.
template <int WK, bool KMULTI>
static cudaError_t launch_step_wk(const DevCfg &c, const StepArgs &a_in, cudaStream_t st) {
    StepArgs a = a_in;
    auto kern = boat_step_kernel<REAL, WK, KMULTI>;
    constexpr int n_setup = setup_warps<WK, KMULTI>();
    constexpr int threads = kTile + 32 * n_setup;
    const int smem = CtaSmem<REAL>(c.block_bytes, c.ncurves, c.npieces, n_setup, !KMULTI && a.rp.state != nullptr).bytes;
    // Occupancy of this instantiation per (device, shared-memory size): queried once each, kept in a small
    // mutex-protected table (boatenv_step and boatenv_step_store use the same kernel with different layouts,
    // several handles / devices / host threads may alternate).
    int ctas_per_sm = 0, n_sm = 0;
    {
        static std::mutex mu;
        static std::vector<OccEntry> table;
        static int max_smem_set[kMaxDevices] = {0};
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
        std::lock_guard<std::mutex> lock(mu);
        const OccEntry *hit = nullptr;
        for (const OccEntry &o : table)
            if (o.dev == dev && o.smem == smem) { hit = &o; break; }
        if (!hit) {
            if (smem > max_smem_set[dev]) {  // the attribute is a maximum: raise it, never lower it
                e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                if (e != cudaSuccess) return e;
                max_smem_set[dev] = smem;
            }
            OccEntry o = {dev, smem, 0, 0};
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o.ctas_per_sm, kern, threads, smem);
            if (e != cudaSuccess) return e;
            e = cudaDeviceGetAttribute(&o.n_sm, cudaDevAttrMultiProcessorCount, dev);
            if (e != cudaSuccess) return e;
            if (o.ctas_per_sm < 1) return cudaErrorLaunchOutOfResources;
            table.push_back(o);
            hit = &table.back();
        }
        ctas_per_sm = hit->ctas_per_sm;
        n_sm = hit->n_sm;
    }
    const long long nblk = ((a.env_end + 31) >> 5) - (a.env_begin >> 5);
    long long grid = (nblk + kWarpsPerCta - 1) / kWarpsPerCta;
    const long long resident = (long long)n_sm * ctas_per_sm;
    if (grid > resident) grid = resident;
    constexpr bool kCurves = (WK == WIND_VEL_CURVE || WK == WIND_ANGLE_RECT || WK == WIND_BOTH);
    bool deferred = false;
    if (KMULTI && kCurves && a.kq_entries && (a.flags & BOATENV_AUTO_RESET)) {
        // one queue region per CTA, big enough for every env the CTA can touch in this launch
        const long long tiles_per_warp = (nblk + grid * kWarpsPerCta - 1) / (grid * kWarpsPerCta);
        const long long cap = tiles_per_warp * kWarpsPerCta * 32;
        if (grid * cap <= a.kq_total && cap < (1LL << 31)) {
            a.kq_cap = (int)cap;
            deferred = true;
        }
    }
    if (!deferred) a.kq_entries = nullptr;
    kern<<<(unsigned)grid, threads, smem, st>>>(c, a);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess || !deferred) return err;
    const int warps_per_region = (int)std::max(1LL, (long long)n_sm * 64 / grid);  // ~2 waves of resident warps
    const long long setup_warps_total = grid * warps_per_region;
    boat_setup_queue_kernel<REAL><<<(unsigned)((setup_warps_total + kWarpsPerCta - 1) / kWarpsPerCta), kTile, 0, st>>>(
        c, a.kq_entries, a.kq_counts, (int)grid, a.kq_cap, warps_per_region);
    count_launch();
    return cudaGetLastError();
}

template <int WK>
static cudaError_t launch_step_k(const DevCfg &c, const StepArgs &a, cudaStream_t st) {
    return a.ksteps == 1 ? launch_step_wk<WK, false>(c, a, st) : launch_step_wk<WK, true>(c, a, st);
}

cudaError_t FN(launch_step_)(const DevCfg &c, const StepArgs &a, cudaStream_t st) {
    if (a.env_end <= a.env_begin) return cudaSuccess;
    if (a.env_begin & 31) return cudaErrorInvalidValue;  // launches start on a state-block boundary
    cudaError_t e;
    switch (c.wind_kind) {
    case WIND_NONE: e = launch_step_k<WIND_NONE>(c, a, st); break;
    case WIND_CONST: e = launch_step_k<WIND_CONST>(c, a, st); break;
    case WIND_VEL_CURVE: e = launch_step_k<WIND_VEL_CURVE>(c, a, st); break;
    case WIND_ANGLE_RECT: e = launch_step_k<WIND_ANGLE_RECT>(c, a, st); break;
    case WIND_BOTH: e = launch_step_k<WIND_BOTH>(c, a, st); break;
    default: return cudaErrorInvalidValue;
    }
    count_launch();
    return e;
}

cudaError_t FN(launch_reset_)(const DevCfg &c, const uint8_t *mask, void *obs_out, cudaStream_t st) {
    boat_reset_kernel<REAL, true><<<grid_for(c.n_envs, kTile), kTile, 0, st>>>(c, mask, reinterpret_cast<REAL *>(obs_out));
    count_launch();
    if (mask != nullptr && c.ncurves > 0) {   // warps with only a few envs to reset: the cooperative instantiation
        boat_reset_kernel<REAL, false><<<grid_for(c.n_envs, kTile), kTile, 0, st>>>(c, mask, reinterpret_cast<REAL *>(obs_out));
        count_launch();
    }
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

cudaError_t FN(launch_env_state_)(const DevCfg &c, long long env, double *out, cudaStream_t st) {
    boat_env_state_kernel<REAL><<<1, 32, 0, st>>>(c, env, out);
    count_launch();
    return cudaGetLastError();
}

cudaError_t FN(launch_fill_actions_)(const DevCfg &c, unsigned long long step_counter, double scale, void *out,
                                     cudaStream_t st) {
    const long long quads = (c.n_envs + (c.env_id_offset & 3) + 3) / 4;  // global quads touched by this shard
    boat_fill_actions_kernel<REAL><<<grid_for(quads, 256), 256, 0, st>>>(c, step_counter, scale,
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
