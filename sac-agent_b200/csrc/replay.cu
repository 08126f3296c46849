// replay.cu -- device-resident ReplayBuffer (agent/buffer.py:3-35): ring store,
// uniform-with-replacement sample-gather, and the fused env.step + agent.remember
// launch (main.py:81-88).  Rows are kept exactly as the reference keeps them --
// five parallel arrays state[size][obs], new_state[size][obs], action[size][na],
// reward[size], terminal[size] -- so a sampled batch is five dense tensors the
// learner can consume without a transpose.
#include <cstring>
#include <new>

#include "common.cuh"
#include "launch.h"

using namespace boatenv;

#define CUDA_TRY(expr)                                  \
    do {                                                \
        cudaError_t _e = (expr);                        \
        if (_e != cudaSuccess) return (int)_e;          \
    } while (0)

namespace boatenv {
DevCfg *handle_cfg(boatenv_t h);
int handle_precision(boatenv_t h);
int handle_device(boatenv_t h);
bool handle_was_reset(boatenv_t h);
int handle_next_parity(boatenv_t h);
}  // namespace boatenv

struct boatreplay_handle {
    void *state, *new_state, *action, *reward;
    uint8_t *terminal;
    long long mem_size, mem_cntr;
    int obs_dim, n_actions, precision, device;
    size_t esize;
};

namespace {

// store_transition (buffer.py:13-22) for n rows: flat element e of the [n][width] input
// goes to ring slot (cntr + e / width) % size.  Consecutive threads write consecutive
// addresses except at the single wrap point.
template <typename T>
__global__ void __launch_bounds__(256) replay_store_kernel(
    T *__restrict__ state, T *__restrict__ new_state, T *__restrict__ action, T *__restrict__ reward,
    uint8_t *__restrict__ terminal, const T *__restrict__ s, const T *__restrict__ a, const T *__restrict__ r,
    const T *__restrict__ s2, const uint8_t *__restrict__ done, long long n, long long cntr, long long size,
    int obs_dim, int n_actions) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    // rows older than the last `size` are overwritten inside this very call: skip them
    const long long first = n > size ? n - size : 0;
    for (long long e = tid + first * obs_dim; e < n * obs_dim; e += stride) {
        const long long row = e / obs_dim;
        const int q = (int)(e - row * obs_dim);
        const long long slot = (cntr + row) % size;
        state[slot * obs_dim + q] = __ldcs(s + e);
        new_state[slot * obs_dim + q] = __ldcs(s2 + e);
    }
    for (long long e = tid + first * n_actions; e < n * n_actions; e += stride) {
        const long long row = e / n_actions;
        const int q = (int)(e - row * n_actions);
        action[((cntr + row) % size) * n_actions + q] = __ldcs(a + e);
    }
    for (long long row = tid + first; row < n; row += stride) {
        const long long slot = (cntr + row) % size;
        reward[slot] = __ldcs(r + row);
        terminal[slot] = done[row] ? 1 : 0;  // np.zeros(..., bool) storage (buffer.py:11,20)
    }
}

// np.random.choice(max_mem, batch) stand-in (buffer.py:27): draw b = word (b & 3) of
// Philox(counter = (b >> 2, sample counter), key = seed), reduced by multiply-shift.
__device__ __forceinline__ long long replay_index(unsigned long long seed, unsigned long long counter, long long b,
                                                  long long max_mem) {
    const unsigned long long blk = (unsigned long long)b >> 2;
    const Philox4 r = philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)counter,
                                    kStreamReplay | (uint32_t)((counter >> 32) & 0x0fffffffu), (uint32_t)seed,
                                    (uint32_t)(seed >> 32));
    const uint32_t w = philox_word(r, (int)(b & 3));
    return (long long)__umul64hi((unsigned long long)w << 32, (unsigned long long)max_mem);
}

// sample_buffer's gather (buffer.py:29-33).  One thread per output element of the two
// [batch][obs_dim] tensors; the first `batch` threads also move action/reward/done.
// USE_IDX: indices supplied by the caller; otherwise drawn in-kernel (and optionally
// exported through idx_out).
template <typename T, bool USE_IDX>
__global__ void __launch_bounds__(256) replay_gather_kernel(
    const T *__restrict__ state, const T *__restrict__ new_state, const T *__restrict__ action,
    const T *__restrict__ reward, const uint8_t *__restrict__ terminal, const long long *__restrict__ idx_in,
    long long *__restrict__ idx_out, T *__restrict__ s_out, T *__restrict__ a_out, T *__restrict__ r_out,
    T *__restrict__ s2_out, uint8_t *__restrict__ d_out, long long batch, long long max_mem, long long size,
    unsigned long long seed, unsigned long long counter, int obs_dim, int n_actions) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = tid; e < batch * obs_dim; e += stride) {
        const long long b = e / obs_dim;
        const int q = (int)(e - b * obs_dim);
        long long j = USE_IDX ? idx_in[b] : replay_index(seed, counter, b, max_mem);
        if (USE_IDX) { if (j < 0) j += size; }  // numpy negative indexing
        s_out[e] = __ldg(state + j * obs_dim + q);
        s2_out[e] = __ldg(new_state + j * obs_dim + q);
    }
    for (long long b = tid; b < batch; b += stride) {
        long long j = USE_IDX ? idx_in[b] : replay_index(seed, counter, b, max_mem);
        if (USE_IDX) { if (j < 0) j += size; }
        for (int q = 0; q < n_actions; ++q) a_out[b * n_actions + q] = __ldg(action + j * n_actions + q);
        r_out[b] = __ldg(reward + j);
        d_out[b] = terminal[j];
        if (!USE_IDX && idx_out) idx_out[b] = j;
    }
}

template <typename T>
cudaError_t launch_store(boatreplay_t r, long long n, const void *s, const void *a, const void *rew, const void *s2,
                         const uint8_t *done, cudaStream_t st) {
    const long long work = n * r->obs_dim;
    long long blocks = (work + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    replay_store_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(
        (T *)r->state, (T *)r->new_state, (T *)r->action, (T *)r->reward, r->terminal, (const T *)s, (const T *)a,
        (const T *)rew, (const T *)s2, done, n, r->mem_cntr, r->mem_size, r->obs_dim, r->n_actions);
    count_launch();
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_gather(boatreplay_t r, long long batch, const long long *idx_in, long long *idx_out,
                          unsigned long long seed, unsigned long long counter, void *s_out, void *a_out, void *r_out,
                          void *s2_out, uint8_t *d_out, cudaStream_t st) {
    const long long max_mem = r->mem_cntr < r->mem_size ? r->mem_cntr : r->mem_size;
    const long long work = batch * r->obs_dim;
    long long blocks = (work + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    if (idx_in)
        replay_gather_kernel<T, true><<<(unsigned)blocks, 256, 0, st>>>(
            (const T *)r->state, (const T *)r->new_state, (const T *)r->action, (const T *)r->reward, r->terminal,
            idx_in, nullptr, (T *)s_out, (T *)a_out, (T *)r_out, (T *)s2_out, d_out, batch, max_mem, r->mem_size, seed,
            counter, r->obs_dim, r->n_actions);
    else
        replay_gather_kernel<T, false><<<(unsigned)blocks, 256, 0, st>>>(
            (const T *)r->state, (const T *)r->new_state, (const T *)r->action, (const T *)r->reward, r->terminal,
            nullptr, idx_out, (T *)s_out, (T *)a_out, (T *)r_out, (T *)s2_out, d_out, batch, max_mem, r->mem_size, seed,
            counter, r->obs_dim, r->n_actions);
    count_launch();
    return cudaGetLastError();
}

}  // namespace

extern "C" {

int boatreplay_create(int64_t max_size, int32_t obs_dim, int32_t n_actions, int precision, int device,
                      boatreplay_t *out) {
    if (!out || max_size <= 0 || obs_dim <= 0 || n_actions <= 0) return BOATENV_EINVAL;
    if (precision != 32 && precision != 64) return BOATENV_EINVAL;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return BOATENV_ENODEVICE;
    CUDA_TRY(cudaSetDevice(device));
    boatreplay_handle *r = new (std::nothrow) boatreplay_handle();
    if (!r) return BOATENV_EINVAL;
    std::memset(r, 0, sizeof(*r));
    r->mem_size = max_size;
    r->obs_dim = obs_dim;
    r->n_actions = n_actions;
    r->precision = precision;
    r->device = device;
    r->esize = precision == 32 ? 4 : 8;
    const size_t n = (size_t)max_size;
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void **p, size_t bytes) {  // np.zeros(...)  buffer.py:7-11
        if (e == cudaSuccess) e = cudaMalloc(p, bytes);
        if (e == cudaSuccess) e = cudaMemset(*p, 0, bytes);
    };
    alloc(&r->state, n * obs_dim * r->esize);
    alloc(&r->new_state, n * obs_dim * r->esize);
    alloc(&r->action, n * n_actions * r->esize);
    alloc(&r->reward, n * r->esize);
    alloc((void **)&r->terminal, n);
    if (e != cudaSuccess) {
        boatreplay_destroy(r);
        return (int)e;
    }
    *out = r;
    return BOATENV_OK;
}

int boatreplay_destroy(boatreplay_t r) {
    if (!r) return BOATENV_EINVAL;
    cudaSetDevice(r->device);
    cudaFree(r->state);
    cudaFree(r->new_state);
    cudaFree(r->action);
    cudaFree(r->reward);
    cudaFree(r->terminal);
    delete r;
    return BOATENV_OK;
}

int boatreplay_store(boatreplay_t r, int64_t n, const void *s, const void *a, const void *rew, const void *s2,
                     const uint8_t *done, void *stream) {
    if (!r || n < 0 || !s || !a || !rew || !s2 || !done) return BOATENV_EINVAL;
    if (n == 0) return BOATENV_OK;
    CUDA_TRY(cudaSetDevice(r->device));
    CUDA_TRY(r->precision == 32 ? launch_store<float>(r, n, s, a, rew, s2, done, (cudaStream_t)stream)
                                : launch_store<double>(r, n, s, a, rew, s2, done, (cudaStream_t)stream));
    r->mem_cntr += n;  // buffer.py:22
    return BOATENV_OK;
}

int boatreplay_sample(boatreplay_t r, int64_t batch, uint64_t seed, uint64_t counter, void *s_out, void *a_out,
                      void *r_out, void *s2_out, uint8_t *done_out, int64_t *idx_out, void *stream) {
    if (!r || batch <= 0 || !s_out || !a_out || !r_out || !s2_out || !done_out) return BOATENV_EINVAL;
    if (r->mem_cntr <= 0) return BOATENV_ESTATE;  // np.random.choice(0, n) raises ValueError
    CUDA_TRY(cudaSetDevice(r->device));
    CUDA_TRY(r->precision == 32
                 ? launch_gather<float>(r, batch, nullptr, (long long *)idx_out, seed, counter, s_out, a_out, r_out,
                                        s2_out, done_out, (cudaStream_t)stream)
                 : launch_gather<double>(r, batch, nullptr, (long long *)idx_out, seed, counter, s_out, a_out, r_out,
                                         s2_out, done_out, (cudaStream_t)stream));
    return BOATENV_OK;
}

int boatreplay_gather(boatreplay_t r, int64_t batch, const int64_t *idx, void *s_out, void *a_out, void *r_out,
                      void *s2_out, uint8_t *done_out, void *stream) {
    if (!r || batch <= 0 || !idx || !s_out || !a_out || !r_out || !s2_out || !done_out) return BOATENV_EINVAL;
    CUDA_TRY(cudaSetDevice(r->device));
    CUDA_TRY(r->precision == 32 ? launch_gather<float>(r, batch, (const long long *)idx, nullptr, 0, 0, s_out, a_out,
                                                       r_out, s2_out, done_out, (cudaStream_t)stream)
                                : launch_gather<double>(r, batch, (const long long *)idx, nullptr, 0, 0, s_out, a_out,
                                                        r_out, s2_out, done_out, (cudaStream_t)stream));
    return BOATENV_OK;
}

int64_t boatreplay_mem_cntr(boatreplay_t r) { return r ? r->mem_cntr : BOATENV_EINVAL; }
int64_t boatreplay_mem_size(boatreplay_t r) { return r ? r->mem_size : BOATENV_EINVAL; }

int boatenv_step_store(boatenv_t h, boatreplay_t r, const void *actions, void *obs_inout, void *reward_out,
                       uint8_t *done_out, uint8_t *term_out, int done_flag_mode, uint32_t flags, void *stream) {
    if (!h || !r || !actions || !obs_inout || !reward_out || !done_out) return BOATENV_EINVAL;
    if (!handle_was_reset(h)) return BOATENV_ESTATE;
    const DevCfg &c = *handle_cfg(h);
    if (handle_precision(h) != r->precision || handle_device(h) != r->device || r->obs_dim != kObsDim ||
        r->n_actions != 1)
        return BOATENV_EINVAL;
    if (c.n_envs > r->mem_size) return BOATENV_EUNSUPPORTED;  // a step must not lap the ring
    if ((reinterpret_cast<uintptr_t>(obs_inout) & 15u) != 0) return BOATENV_EALIGN;
    CUDA_TRY(cudaSetDevice(r->device));
    StepArgs a;
    std::memset(&a, 0, sizeof(a));
    a.env_begin = 0;
    a.env_end = c.n_envs;
    a.actions = actions;
    a.ksteps = 1;
    a.obs_out = obs_inout;
    a.obs_in = obs_inout;
    a.reward_out = reward_out;
    a.done_out = done_out;
    a.term_out = term_out;
    a.flags = flags;
    a.rp.state = r->state;
    a.rp.new_state = r->new_state;
    a.rp.action = r->action;
    a.rp.reward = r->reward;
    a.rp.terminal = r->terminal;
    a.rp.mem_size = r->mem_size;
    a.rp.base_slot = r->mem_cntr % r->mem_size;
    a.rp.done_flag_mode = done_flag_mode;
    a.reverse = boatenv::handle_next_parity(h);
    CUDA_TRY(r->precision == 32 ? launch_step_f32(c, a, (cudaStream_t)stream)
                                : launch_step_f64(c, a, (cudaStream_t)stream));
    r->mem_cntr += c.n_envs;
    return BOATENV_OK;
}

}  // extern "C"
