// policy_mlp.cu -- the actor's forward pass for acting (ActorNetwork.forward + sample_normal without
// gradients, networks/networks.py:38-70; choose_action, agent/continuous_agent.py:57-61) over N envs as ONE
// persistent kernel on the 5th-generation tensor cores (tcgen05, accumulators in TMEM).
//
//   obs[N][obs_dim] -> relu(fc1) -> relu(fc2) -> (mean, std heads) -> tanh-squashed draw -> action[N][A]
//
// The three dense layers are 256 wide.  Through cuBLAS every layer writes its [N][256] activations to HBM
// and the next one reads them back (1 GB per layer for a million envs); here a CTA keeps a 128-env tile
// on chip from the observations to the actions:
//   * all weights (bf16, 144 kB) live in shared memory for the life of the CTA, in the K-major
//     no-swizzle core-matrix layout the UMMA descriptors address directly;
//   * the two hidden layers are tcgen05.mma (M = 128 envs, N = 256, K steps of 16) into TMEM; eight epilogue
//     warps read the accumulators back (tcgen05.ld, one TMEM lane = one env, half a row per thread): after
//     layer 1 they add the bias, apply ReLU, round to bf16 and write the tile as the A operand of layer 2;
//     after layer 2 the fp32 activations go straight from registers into the dot products of the two heads
//     (fp32 weights), the draw and the squash -- the heads never leave the registers;
//   * the next tile's observations are fetched while the current tile computes;
//   * HBM traffic is the input and the output: 4 * obs_dim + 4 * A bytes per env.
// Inputs are rounded to bf16 (fp32 accumulation): acting only, the learner never sees this kernel.
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/boatenv.h"
#include "common.cuh"
#include "launch.h"

namespace {

constexpr int kHidden = 256, kTileM = 128, kK1 = 16, kN3 = 16;
// 8 epilogue warps: warp w works on TMEM lanes 32 (w % 4) .. + 31 (the hardware's lane window of a warp), i.e. on
// env row 32 (w % 4) + lane, and on the accumulator columns 128 (w / 4) .. + 127 of that row; warp 8 issues the MMAs.
constexpr int kEpiThreads = 256, kThreads = kEpiThreads + 32, kMmaWarp = kEpiThreads / 32;
// shared-memory map (bytes).  Operand layout: element (row, k) of an R-row operand sits at
// (k / 8) * (R * 16) + row * 16 + (k % 8) * 2  -- 8x8 core matrices, K-adjacent cores R*16 bytes apart (LBO),
// row groups 128 bytes apart (SBO).
constexpr int kOffW1 = 0;                                   // [256 rows][16]
constexpr int kOffW2 = kOffW1 + kHidden * kK1 * 2;          // [256 rows][256]
constexpr int kOffW3 = kOffW2 + kHidden * kHidden * 2;      // fp32 [16 rows][256], row-major: the heads stay in fp32
constexpr int kOffB1 = kOffW3 + kN3 * kHidden * 4;          // fp32[256]
constexpr int kOffB2 = kOffB1 + kHidden * 4;
constexpr int kOffB3 = kOffB2 + kHidden * 4;                // fp32[16]
constexpr int kBlobBytes = kOffB3 + kN3 * 4;                // what the host packs (BOATAGENT_POLICY_BLOB_BYTES)
constexpr int kOffA0 = (kBlobBytes + 127) / 128 * 128;      // [128 rows][16]
constexpr int kOffA1 = kOffA0 + kTileM * kK1 * 2;           // [128 rows][256]
constexpr int kOffPart = kOffA1;                            // fp32 [128 rows][16]: head partial sums of the upper column half
                                                            // (aliases A1: layer 2 has finished reading it by then)
constexpr int kOffBar = kOffA1 + kTileM * kHidden * 2;
constexpr int kSmemBytes = kOffBar + 16;
static_assert(kBlobBytes == BOATAGENT_POLICY_BLOB_BYTES, "header and kernel disagree on the weight blob");
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// UMMA shared-memory descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor: start >> 4 at [0,14), LBO >> 4 at
// [16,30), SBO >> 4 at [32,46), version 1 at [46,48), layout type 0 at [61,64)).
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// Instruction descriptor of tcgen05.mma.kind::f16 (cute::UMMA::InstrDescriptor): D = f32, A = B = bf16, both K-major.
__host__ __device__ constexpr uint32_t umma_idesc(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {   // arrives on `bar` when every MMA issued so far is done
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n\t"
        "D_%=:\n\t"
        "}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
// 32 consecutive fp32 accumulator columns of this thread's TMEM lane: asynchronous load (the registers are valid after
// tmem_ld_wait()), so the next chunk can be in flight while the current one is processed
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
// The registers are tied to the wait as in/out operands: nothing that reads them can be scheduled above it.
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                   "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&p);
}

// Layer-1 epilogue for the thread's half row (columns col0 .. col0 + 127): accumulator -> + bias -> ReLU -> bf16 ->
// A operand tile of layer 2.
__device__ __forceinline__ void hidden_epilogue(uint32_t tmem_row, const float *bias, unsigned char *a1, int row, int col0) {
    uint32_t ra[32], rb[32];
    tmem_ld32_issue(tmem_row + (uint32_t)col0, ra);
    auto chunk = [&](const uint32_t (&r)[32], int c) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {   // four 8-column cores
            const float4 b0 = *reinterpret_cast<const float4 *>(bias + c * 32 + q * 8);
            const float4 b1 = *reinterpret_cast<const float4 *>(bias + c * 32 + q * 8 + 4);
            uint4 out;
            out.x = pack_bf16(fmaxf(__uint_as_float(r[q * 8 + 0]) + b0.x, 0.f), fmaxf(__uint_as_float(r[q * 8 + 1]) + b0.y, 0.f));
            out.y = pack_bf16(fmaxf(__uint_as_float(r[q * 8 + 2]) + b0.z, 0.f), fmaxf(__uint_as_float(r[q * 8 + 3]) + b0.w, 0.f));
            out.z = pack_bf16(fmaxf(__uint_as_float(r[q * 8 + 4]) + b1.x, 0.f), fmaxf(__uint_as_float(r[q * 8 + 5]) + b1.y, 0.f));
            out.w = pack_bf16(fmaxf(__uint_as_float(r[q * 8 + 6]) + b1.z, 0.f), fmaxf(__uint_as_float(r[q * 8 + 7]) + b1.w, 0.f));
            *reinterpret_cast<uint4 *>(a1 + (c * 4 + q) * (kTileM * 16) + row * 16) = out;
        }
    };
    const int c0 = col0 / 32;
    tmem_ld_wait(ra); tmem_ld32_issue(tmem_row + (uint32_t)(col0 + 32), rb); chunk(ra, c0);
    tmem_ld_wait(rb); tmem_ld32_issue(tmem_row + (uint32_t)(col0 + 64), ra); chunk(rb, c0 + 1);
    tmem_ld_wait(ra); tmem_ld32_issue(tmem_row + (uint32_t)(col0 + 96), rb); chunk(ra, c0 + 2);
    tmem_ld_wait(rb); chunk(rb, c0 + 3);
}

// Layer-2 epilogue fused with the two heads: h2 = relu(acc + bias) stays in fp32 registers and goes straight into
// the 2 * n_actions dot products with the fp32 head weights (shared memory, broadcast reads).  acc[j] += partial sums
// over this thread's 128 columns.
template <int NH>
__device__ __forceinline__ void heads_epilogue(uint32_t tmem_row, const float *bias, const float *w3, int col0, float (&acc)[NH]) {
#pragma unroll
    for (int j = 0; j < NH; ++j) acc[j] = 0.f;
    uint32_t ra[32], rb[32];
    tmem_ld32_issue(tmem_row + (uint32_t)col0, ra);
    auto chunk = [&](const uint32_t (&r)[32], int c) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 b = *reinterpret_cast<const float4 *>(bias + c * 32 + 4 * i);
            v[4 * i + 0] = fmaxf(__uint_as_float(r[4 * i + 0]) + b.x, 0.f);
            v[4 * i + 1] = fmaxf(__uint_as_float(r[4 * i + 1]) + b.y, 0.f);
            v[4 * i + 2] = fmaxf(__uint_as_float(r[4 * i + 2]) + b.z, 0.f);
            v[4 * i + 3] = fmaxf(__uint_as_float(r[4 * i + 3]) + b.w, 0.f);
        }
#pragma unroll
        for (int j = 0; j < NH; ++j) {
            const float4 *w = reinterpret_cast<const float4 *>(w3 + j * kHidden + c * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 ww = w[i];
                acc[j] = fmaf(v[4 * i + 0], ww.x, acc[j]);
                acc[j] = fmaf(v[4 * i + 1], ww.y, acc[j]);
                acc[j] = fmaf(v[4 * i + 2], ww.z, acc[j]);
                acc[j] = fmaf(v[4 * i + 3], ww.w, acc[j]);
            }
        }
    };
    const int c0 = col0 / 32;
    tmem_ld_wait(ra); tmem_ld32_issue(tmem_row + (uint32_t)(col0 + 32), rb); chunk(ra, c0);
    tmem_ld_wait(rb); tmem_ld32_issue(tmem_row + (uint32_t)(col0 + 64), ra); chunk(rb, c0 + 1);
    tmem_ld_wait(ra); tmem_ld32_issue(tmem_row + (uint32_t)(col0 + 96), rb); chunk(ra, c0 + 2);
    tmem_ld_wait(rb); chunk(rb, c0 + 3);
}

template <int NA>   // NA = n_actions (1, 2, 4 or 8 instantiated)
__global__ void __launch_bounds__(kThreads, 1)
policy_mlp_kernel(const unsigned char *__restrict__ blob, const float *__restrict__ obs, const float *__restrict__ eps,
                  const float *__restrict__ max_action, unsigned long long seed, unsigned long long step, long long n,
                  int obs_dim, float *__restrict__ action_out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t tmem_base_slot;
    constexpr int NH = 2 * NA;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = (warp & 3) * 32 + lane;      // env row of the tile this epilogue thread works on
    const int col0 = (warp >> 2) * 128;          // its half of the accumulator columns (epilogue warps only)
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + kOffBar);

    // ---- one-time setup: weights into shared memory, barrier, TMEM ----
    for (int i = tid; i < kBlobBytes / 16; i += kThreads)
        reinterpret_cast<uint4 *>(smem)[i] = __ldg(reinterpret_cast<const uint4 *>(blob) + i);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar + 1)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {   // the MMA warp owns the 512 TMEM columns (two 128 x 256 fp32 accumulators)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr(&tmem_base_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    proxy_fence();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_slot;
    const uint32_t tmem_row = tmem + ((uint32_t)(warp & 3) * 32u << 16);
    const uint32_t sbase = smem_addr(smem);
    const float *b1 = reinterpret_cast<const float *>(smem + kOffB1), *b2 = reinterpret_cast<const float *>(smem + kOffB2),
                *b3 = reinterpret_cast<const float *>(smem + kOffB3), *w3 = reinterpret_cast<const float *>(smem + kOffW3);
    unsigned char *a0 = smem + kOffA0, *a1 = smem + kOffA1;
    float *part = reinterpret_cast<float *>(smem + kOffPart);
    uint32_t parity = 0;

    const long long n_tiles = (n + kTileM - 1) / kTileM;
    uint64_t *bar1 = bar, *bar2 = bar + 1;   // layer-1 / layer-2 accumulator ready
    // Software pipeline over the CTA's tiles: layer 1 of tile t + 1 is issued right behind layer 2 of tile t, so it
    // (and the staging of its observations) runs under the layer-2 epilogue of tile t; the observation rows are
    // fetched from HBM one more tile ahead (warps 0-3: one row each).
    float o[kK1];
    auto fetch_obs = [&](long long t) {
        const long long e = t * kTileM + row;
#pragma unroll
        for (int q = 0; q < kK1; ++q) o[q] = (q < obs_dim && t < n_tiles && e < n) ? __ldg(obs + e * obs_dim + q) : 0.f;
    };
    auto stage_obs = [&]() {   // registers -> bf16 A0 [128][16]
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            uint4 out;
            out.x = pack_bf16(o[j * 8 + 0], o[j * 8 + 1]);
            out.y = pack_bf16(o[j * 8 + 2], o[j * 8 + 3]);
            out.z = pack_bf16(o[j * 8 + 4], o[j * 8 + 5]);
            out.w = pack_bf16(o[j * 8 + 6], o[j * 8 + 7]);
            *reinterpret_cast<uint4 *>(a0 + j * (kTileM * 16) + row * 16) = out;
        }
    };
    auto issue_layer1 = [&]() {   // D1[128][256] = A0 . W1^T (one K step)
        umma_bf16(tmem, umma_desc(sbase + kOffA0, kTileM * 16, 128), umma_desc(sbase + kOffW1, kHidden * 16, 128),
                  umma_idesc(kTileM, kHidden), 0u);
        umma_commit(bar1);
    };
    if (warp < 4) {
        fetch_obs(blockIdx.x);
        stage_obs();
        proxy_fence();
        fetch_obs((long long)blockIdx.x + gridDim.x);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp && lane == 0) {
        tc_fence_after();
        issue_layer1();
    }
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long env = tile * kTileM + row;
        const bool more = tile + gridDim.x < n_tiles;
        if (warp < kMmaWarp) {
            bar_wait(bar1, parity);          // layer 1 of this tile is in TMEM (and has finished reading A0)
            tc_fence_after();
            hidden_epilogue(tmem_row, b1, a1, row, col0);
            if (warp < 4 && more) {          // next tile's observations -> A0, the tile after that -> registers
                stage_obs();
                fetch_obs(tile + 2 * (long long)gridDim.x);
            }
            proxy_fence();
        }
        tc_fence_before();
        __syncthreads();
        // ---- layer 2: D2[128][256] = A1 . W2^T (16 K steps), then layer 1 of the next tile behind it ----
        if (warp == kMmaWarp && lane == 0) {
            tc_fence_after();
#pragma unroll 1
            for (int k = 0; k < kHidden / 16; ++k)
                umma_bf16(tmem + 256u, umma_desc(sbase + kOffA1 + k * 2 * (kTileM * 16), kTileM * 16, 128),
                          umma_desc(sbase + kOffW2 + k * 2 * (kHidden * 16), kHidden * 16, 128),
                          umma_idesc(kTileM, kHidden), k > 0 ? 1u : 0u);
            umma_commit(bar2);
            if (more) issue_layer1();
        }
        if (warp < kMmaWarp) {
            bar_wait(bar2, parity);
            tc_fence_after();
            // ---- layer-2 epilogue + heads (fp32): rows of w3 are the mean heads, then the std heads ----
            float acc[NH];
            heads_epilogue<NH>(tmem_row + 256u, b2, w3, col0, acc);
            if (col0 != 0) {
#pragma unroll
                for (int j = 0; j < NH; ++j) part[row * kN3 + j] = acc[j];
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");   // the eight epilogue warps only
            if (col0 == 0 && env < n) {
#pragma unroll
                for (int a = 0; a < NA; ++a) {   // sample_normal, networks.py:47-65
                    const float mean = acc[a] + part[row * kN3 + a] + b3[a];
                    const float raw = acc[NA + a] + part[row * kN3 + NA + a] + b3[NA + a];
                    const float log_std = -5.0f + 3.5f * (tanhf(raw) + 1.0f);
                    float e;
                    if (eps) {
                        e = __ldg(eps + env * NA + a);
                    } else {   // Box-Muller on Philox(seed; env, step, action)
                        const boatenv::Philox4 r = boatenv::philox4x32_10(
                            (uint32_t)env, (uint32_t)((unsigned long long)env >> 32), (uint32_t)step,
                            0xD0000000u | ((uint32_t)a << 16) | (uint32_t)((step >> 32) & 0xffffu), (uint32_t)seed,
                            (uint32_t)(seed >> 32));
                        const float u1 = ((float)(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
                        const float u2 = ((float)(r.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
                        e = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
                    }
                    action_out[env * NA + a] = tanhf(mean + e * expf(log_std)) * __ldg(max_action + a);
                }
            }
        }
        parity ^= 1u;
        tc_fence_before();
        __syncthreads();   // D2, A1 and the partial sums are free for the next tile
        tc_fence_after();
    }
    if (warp == kMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

}  // namespace

extern "C" int boatagent_policy_act(const void *weight_blob, const float *obs, const float *eps, const float *max_action,
                                    uint64_t seed, uint64_t step, int64_t n, int32_t obs_dim, int32_t n_actions,
                                    float *action_out, void *stream) {
    if (!weight_blob || !obs || !max_action || !action_out || n <= 0) return BOATENV_EINVAL;
    if (obs_dim < 1 || obs_dim > kK1) return BOATENV_EUNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(weight_blob) & 15u) != 0) return BOATENV_EALIGN;
    using kern_t = void (*)(const unsigned char *, const float *, const float *, const float *, unsigned long long,
                            unsigned long long, long long, int, float *);
    kern_t kern = nullptr;
    switch (n_actions) {   // the head count is a template parameter (its dot products live in registers)
    case 1: kern = policy_mlp_kernel<1>; break;
    case 2: kern = policy_mlp_kernel<2>; break;
    case 4: kern = policy_mlp_kernel<4>; break;
    case 8: kern = policy_mlp_kernel<8>; break;
    default: return BOATENV_EUNSUPPORTED;
    }
    static int n_sm = 0, configured_dev = -1;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (configured_dev != dev) {
        for (kern_t k : {(kern_t)policy_mlp_kernel<1>, (kern_t)policy_mlp_kernel<2>, (kern_t)policy_mlp_kernel<4>,
                         (kern_t)policy_mlp_kernel<8>}) {
            e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
            if (e != cudaSuccess) return (int)e;
        }
        e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return (int)e;
        configured_dev = dev;
    }
    const long long tiles = (n + kTileM - 1) / kTileM;
    const unsigned grid = (unsigned)(tiles < n_sm ? tiles : n_sm);
    kern<<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>((const unsigned char *)weight_blob, obs, eps, max_action, seed,
                                                              step, n, obs_dim, action_out);
    boatenv::count_launch();
    return (int)cudaGetLastError();
}
