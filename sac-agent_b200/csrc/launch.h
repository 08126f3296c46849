// launch.h -- launcher prototypes shared between the per-precision translation units
// (step_f32.cu, step_f64.cu) and the C-ABI front end (abi.cu).
#pragma once
#include "common.cuh"

// Every handle-based entry point runs on the handle's device and restores the caller's current device on return
// (a multi-GPU process keeps torch's current device untouched).
namespace boatenv {
class DeviceGuard {
public:
    explicit DeviceGuard(int device) : prev_(-1), err_(cudaSuccess) {
        err_ = cudaGetDevice(&prev_);
        if (err_ == cudaSuccess && prev_ != device) err_ = cudaSetDevice(device); else if (err_ == cudaSuccess) prev_ = -1;
    }
    ~DeviceGuard() { if (prev_ >= 0) cudaSetDevice(prev_); }
    bool ok() const { return err_ == cudaSuccess; }
    cudaError_t error() const { return err_; }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
private:
    int prev_;
    cudaError_t err_;
};
}  // namespace boatenv
#define GUARD_DEVICE(h)                                   \
    boatenv::DeviceGuard _guard((h)->device);                      \
    if (!_guard.ok()) return (int)_guard.error()


namespace boatenv {

void count_launch();

cudaError_t launch_step_f32(const DevCfg &, const StepArgs &, cudaStream_t);
cudaError_t launch_step_f64(const DevCfg &, const StepArgs &, cudaStream_t);
cudaError_t launch_reset_f32(const DevCfg &, const uint8_t *mask, void *obs_out, cudaStream_t);
cudaError_t launch_reset_f64(const DevCfg &, const uint8_t *mask, void *obs_out, cudaStream_t);
cudaError_t launch_get_field_f32(const DevCfg &, int field, void *out, cudaStream_t);
cudaError_t launch_get_field_f64(const DevCfg &, int field, void *out, cudaStream_t);
cudaError_t launch_set_field_f32(const DevCfg &, int field, const void *in, cudaStream_t);
cudaError_t launch_set_field_f64(const DevCfg &, int field, const void *in, cudaStream_t);
cudaError_t launch_fill_actions_f32(const DevCfg &, unsigned long long step_counter, double scale, void *out,
                                    cudaStream_t);
cudaError_t launch_fill_actions_f64(const DevCfg &, unsigned long long step_counter, double scale, void *out,
                                    cudaStream_t);
cudaError_t launch_env_state_f32(const DevCfg &, long long env, double *out, cudaStream_t);
cudaError_t launch_env_state_f64(const DevCfg &, long long env, double *out, cudaStream_t);
cudaError_t launch_wind_table(const DevCfg &, long long env, double *wv, double *wa, cudaStream_t);
cudaError_t launch_reduce_counters(const double *counters, double *out, cudaStream_t);

}  // namespace boatenv
