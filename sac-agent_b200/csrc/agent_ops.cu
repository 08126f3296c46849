// agent_ops.cu -- the tanh-squashed Gaussian head of the reference's actor
// (networks/networks.py:47-70, ActorNetwork.sample_normal) as ONE kernel forward and ONE kernel
// backward.  In PyTorch the head is ~23 element-wise launches forward and ~35 backward on [B, n_actions]
// tensors -- 80 of the ~190 kernels of a captured SAC update, all launch-latency.
//
//   t = tanh(raw_std);  log_std = -5 + 3.5 (t + 1);  std = exp(log_std)          (:53-56)
//   u = mean + eps * std                     (Normal.rsample / .sample, :60-63; eps ~ N(0, 1) supplied)
//   action = tanh(u) * max_action                                               (:65)
//   log_prob = sum_a [ -(u - mean)^2 / (2 std^2) - log_std - log sqrt(2 pi) - log(1 - action^2 + 1e-6) ]   (:66-68)
//
// Backward is for the reparameterised draw (u a function of mean and std).  The Gaussian term is then
// -eps^2 / 2 exactly: autograd's two paths into it cancel, so its gradient is 0 here.
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/boatenv.h"
#include "launch.h"

namespace {

constexpr float kLogStdMin = -5.0f, kLogStdHalfSpan = 3.5f;  // 0.5 * (LOG_STD_MAX - LOG_STD_MIN)
constexpr float kHalfLog2Pi = 0.91893853320467274178f;
constexpr float kReparamNoise = 1e-6f;

__global__ void __launch_bounds__(256) gaussian_head_fwd_kernel(const float *__restrict__ mean,
                                                                const float *__restrict__ raw_std,
                                                                const float *__restrict__ eps,
                                                                const float *__restrict__ max_action, long long rows,
                                                                int n_actions, float *__restrict__ action_out,
                                                                float *__restrict__ log_prob_out) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= rows) return;
    float lp = 0.0f;
    for (int a = 0; a < n_actions; ++a) {
        const long long i = b * n_actions + a;
        const float m = mean[i];
        const float log_std = kLogStdMin + kLogStdHalfSpan * (tanhf(raw_std[i]) + 1.0f);
        const float sd = expf(log_std);
        const float u = m + eps[i] * sd;
        const float act = tanhf(u) * max_action[a];
        const float d = u - m;
        lp += -(d * d) / (2.0f * sd * sd) - log_std - kHalfLog2Pi - logf(1.0f - act * act + kReparamNoise);
        action_out[i] = act;
    }
    log_prob_out[b] = lp;
}

__global__ void __launch_bounds__(256) gaussian_head_bwd_kernel(
    const float *__restrict__ mean, const float *__restrict__ raw_std, const float *__restrict__ eps,
    const float *__restrict__ max_action, const float *__restrict__ grad_action /* may be null */,
    const float *__restrict__ grad_log_prob /* [rows], may be null */, long long rows, int n_actions,
    float *__restrict__ grad_mean, float *__restrict__ grad_raw_std) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= rows) return;
    const float glp = grad_log_prob ? grad_log_prob[b] : 0.0f;
    for (int a = 0; a < n_actions; ++a) {
        const long long i = b * n_actions + a;
        const float t = tanhf(raw_std[i]);
        const float dls = kLogStdHalfSpan * (1.0f - t * t);          // d log_std / d raw_std
        const float sd = expf(kLogStdMin + kLogStdHalfSpan * (t + 1.0f));
        const float e = eps[i];
        const float th = tanhf(mean[i] + e * sd);
        const float ma = max_action[a];
        const float act = th * ma;
        const float dact = ma * (1.0f - th * th);                      // d action / d u
        const float ga = grad_action ? grad_action[i] : 0.0f;
        // d/du of -log(1 - action^2 + 1e-6) is 2 action dact / (1 - action^2 + 1e-6)
        const float gu = ga * dact + glp * (2.0f * act * dact) / (1.0f - act * act + kReparamNoise);
        grad_mean[i] = gu;
        grad_raw_std[i] = gu * e * sd * dls - glp * dls;              // through std (u) and through -log_std
    }
}

}  // namespace

extern "C" {

int boatagent_gaussian_head_forward(const float *mean, const float *raw_std, const float *eps, const float *max_action,
                                    int64_t rows, int32_t n_actions, float *action_out, float *log_prob_out,
                                    void *stream) {
    if (!mean || !raw_std || !eps || !max_action || !action_out || !log_prob_out || rows <= 0 || n_actions <= 0)
        return BOATENV_EINVAL;
    gaussian_head_fwd_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        mean, raw_std, eps, max_action, rows, n_actions, action_out, log_prob_out);
    boatenv::count_launch();
    return (int)cudaGetLastError();
}

int boatagent_gaussian_head_backward(const float *mean, const float *raw_std, const float *eps, const float *max_action,
                                     const float *grad_action, const float *grad_log_prob, int64_t rows,
                                     int32_t n_actions, float *grad_mean_out, float *grad_raw_std_out, void *stream) {
    if (!mean || !raw_std || !eps || !max_action || !grad_mean_out || !grad_raw_std_out || rows <= 0 || n_actions <= 0)
        return BOATENV_EINVAL;
    gaussian_head_bwd_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        mean, raw_std, eps, max_action, grad_action, grad_log_prob, rows, n_actions, grad_mean_out, grad_raw_std_out);
    boatenv::count_launch();
    return (int)cudaGetLastError();
}

}  // extern "C"

// ---------------------------------------------------------------------------------
// Adam (torch.optim.Adam defaults, networks.py:31,88,121) for every network of the agent plus the
// Polyak average of the target value network (continuous_agent.py:66-80) in ONE launch.
// torch's multi-tensor Adam hands each CTA a 65536-element chunk: the agent's ~350 k parameters then
// run on a dozen CTAs (~70 us per optimiser, twice per update, plus two more launches for the Polyak
// average).  Here a CTA takes 2048 elements; the slot table travels as a kernel argument, so the launch
// captures into a CUDA graph as it is.
// ---------------------------------------------------------------------------------
namespace {

constexpr int kAdamChunk = 2048, kAdamThreads = 256;
struct AdamTable {
    boatagent_adam_slot slot[BOATAGENT_ADAM_MAX_SLOTS];
    int n_slots;
    float beta1, beta2, eps, tau;
};

__global__ void __launch_bounds__(kAdamThreads) adam_polyak_kernel(const __grid_constant__ AdamTable tab,
                                                                   long long *__restrict__ state /* [step, ticket] */) {
    __shared__ int s_slot;
    __shared__ long long s_first;
    __shared__ float s_bc1, s_bc2_sqrt;
    if (threadIdx.x == 0) {
        long long chunk = blockIdx.x;
        int k = 0;
        for (; k < tab.n_slots; ++k) {
            const long long nchunks = (tab.slot[k].numel + kAdamChunk - 1) / kAdamChunk;
            if (chunk < nchunks) break;
            chunk -= nchunks;
        }
        s_slot = k;
        s_first = chunk * kAdamChunk;
        const double t = (double)(*reinterpret_cast<volatile long long *>(state) + 1);  // this step's number
        s_bc1 = (float)(1.0 - pow((double)tab.beta1, t));
        s_bc2_sqrt = (float)sqrt(1.0 - pow((double)tab.beta2, t));
    }
    __syncthreads();
    if (s_slot < tab.n_slots) {
        const boatagent_adam_slot &sl = tab.slot[s_slot];
        const float step_size = sl.lr / s_bc1;
        const long long end = min(s_first + (long long)kAdamChunk, (long long)sl.numel);
        for (long long i = s_first + threadIdx.x; i < end; i += kAdamThreads) {
            const float g = sl.grad[i];
            const float m = sl.exp_avg[i] + (g - sl.exp_avg[i]) * (1.0f - tab.beta1);   // lerp_(grad, 1 - beta1)
            const float v = sl.exp_avg_sq[i] * tab.beta2 + (1.0f - tab.beta2) * g * g;  // mul_(beta2).addcmul_(g, g, 1 - beta2)
            const float denom = sqrtf(v) / s_bc2_sqrt + tab.eps;
            const float p = sl.param[i] - step_size * (m / denom);                       // addcdiv_(m, denom, -step_size)
            sl.exp_avg[i] = m;
            sl.exp_avg_sq[i] = v;
            sl.param[i] = p;
            if (sl.target) sl.target[i] = tab.tau * p + (1.0f - tab.tau) * sl.target[i];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // the last CTA to finish publishes the new step count
        __threadfence();
        const unsigned long long done = atomicAdd(reinterpret_cast<unsigned long long *>(state + 1), 1ULL) + 1ULL;
        if (done == gridDim.x) {
            state[1] = 0;
            state[0] += 1;
        }
    }
}

}  // namespace

extern "C" int boatagent_adam_polyak_step(const boatagent_adam_slot *slots_host, int32_t n_slots, float beta1, float beta2,
                                          float eps, float tau, int64_t *state_dev, void *stream) {
    if (!slots_host || n_slots <= 0 || n_slots > BOATAGENT_ADAM_MAX_SLOTS || !state_dev) return BOATENV_EINVAL;
    AdamTable tab;
    long long chunks = 0;
    for (int k = 0; k < n_slots; ++k) {
        const boatagent_adam_slot &s = slots_host[k];
        if (!s.param || !s.grad || !s.exp_avg || !s.exp_avg_sq || s.numel <= 0) return BOATENV_EINVAL;
        tab.slot[k] = s;
        chunks += (s.numel + kAdamChunk - 1) / kAdamChunk;
    }
    tab.n_slots = n_slots;
    tab.beta1 = beta1;
    tab.beta2 = beta2;
    tab.eps = eps;
    tab.tau = tau;
    adam_polyak_kernel<<<(unsigned)chunks, kAdamThreads, 0, (cudaStream_t)stream>>>(tab, (long long *)state_dev);
    boatenv::count_launch();
    return (int)cudaGetLastError();
}
