// wind_setup.cuh -- warp-cooperative episode setup (the "slow path" of the step kernel).
//
// The reference rebuilds two 10000-sample tables per episode (wind.py:69-90: a cubic
// not-a-knot spline through `fixed_points` random knots, sampled at linspace(0,L,L),
// min/max renormalised iff a sample leaves [0,1]) -- 160 kB per env, which cannot be
// batched to millions of envs.  Here an env carries only the 4 cubic coefficients of
// the spline piece it is currently in (renormalisation and the experiment's scale folded
// in).  They are re-derived whenever an env starts an episode or walks into the next
// piece, by the whole warp working on that one env:
//   lane k                 : knot k from Philox(seed, env, episode)          (wind.py:78)
//   lane (curve,piece,m)   : coefficient m of a piece = basis row . knots    (wind.py:82-85)
//   lane (curve,piece)     : that piece's extremal SAMPLES (the samples next to the roots
//                            of the derivative and at the piece ends); a warp min/max
//                            then equals np.min/np.max over all L samples     (wind.py:87-89)
// All of it in fp64, in both precision modes, so fp32 and fp64 envs take the same
// renormalisation branch.
#pragma once
#include "common.cuh"

namespace boatenv {

// Per-warp scratch (doubles): [0..7] the folded result a[4], b[4]; [8..39] the knots
// [curve][16]; [40..] the coefficient table [curve][piece][4].
constexpr int kKnotOff = 8, kCoefOff = 8 + 2 * kMaxKnots;
constexpr int kScratchDoubles = kCoefOff + 2 * (kMaxKnots - 1) * 4;  // upper bound (static allocations)
__host__ __device__ constexpr int scratch_doubles(int ncurves, int npieces) { return kCoefOff + ncurves * npieces * 4; }

// Piece index and local coordinate numerator of wind sample `index` (0 <= index < L):
// x_index / h = index * (fp-1) / (L-1) exactly (x_index = index * L/(L-1), h = L/(fp-1)),
// so piece j = floor(index * npieces / Lm1) (capped) and s = r / Lm1 with r the remainder.
// The division is a multiply-high by a host-derived magic number (exact on this range).
__device__ __forceinline__ void piece_of(const DevCfg &c, int index, int &j, int &r) {
    const uint32_t num = (uint32_t)index * (uint32_t)c.npieces;
    uint32_t q = __umulhi(num, c.magic_m) >> c.magic_s;
    q = min(q, (uint32_t)(c.npieces - 1));
    j = (int)q;
    r = (int)(num - q * (uint32_t)c.Lm1);
}

// ---------------------------------------------------------------------------------
// The arithmetic of one episode setup, shared by the two work mappings below (warp-cooperative: one env per
// warp, used inside the step kernels where requests are sparse; lane-parallel: one env per lane, used by the reset
// kernel and the episode-end queue kernel where they are dense).  Both mappings call exactly these functions with
// the same operands, so their results are bit-identical.
// ---------------------------------------------------------------------------------

// coefficient m of piece j in the local coordinate s: c_m = sum_k basis[(j*4+m)*fp + k] * u_k  (two accumulators)
__device__ __forceinline__ double coef_dot(const double *__restrict__ brow, const double *u, int fp) {
    double acc0 = 0.0, acc1 = 0.0;
    if (fp == 8) {  // the reference's default (original_config.yaml:63): fully unrolled
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
            acc0 = fma(__ldg(brow + k), u[k], acc0);
            acc1 = fma(__ldg(brow + k + 1), u[k + 1], acc1);
        }
    } else {
        int k = 0;
        for (; k + 1 < fp; k += 2) {
            acc0 = fma(__ldg(brow + k), u[k], acc0);
            acc1 = fma(__ldg(brow + k + 1), u[k + 1], acc1);
        }
        if (k < fp) acc0 = fma(__ldg(brow + k), u[k], acc0);
    }
    return acc0 + acc1;
}

// Extremal SAMPLES of spline piece `sub` with local coefficients k0..k3 (np.min / np.max of wind.py:87-89 restricted
// to the samples piece_of() assigns to this piece): folded into mn / mx.
// The discrete extremes of the curve sit at the first / last sample of a piece or next to a root of the
// derivative.  Only the samples of THIS piece are looked at (evaluated with the same piece and local coordinate
// the step kernel will use): a neighbour on the far side of a piece boundary is that piece's own first / last
// sample.  Roots are LOCATED in fp32 -- a 4-sample window absorbs the location error -- candidates are EVALUATED
// in fp64.
__device__ __forceinline__ void piece_extremes(const DevCfg &c, int sub, double k0, double k1, double k2, double k3,
                                               double &mn, double &mx) {
    const int np = c.npieces;
    const int first = __ldg(c.piece_bounds + 2 * sub), last = __ldg(c.piece_bounds + 2 * sub + 1);
    const int r_base = sub * c.Lm1;
    const bool has_samples = first <= last;  // a table shorter than the spline can leave a piece without samples
    auto consider = [&](int index) {
        if (!has_samples) return;
        index = max(first, min(index, last));
        const double sl = (double)(index * np - r_base) * c.inv_Lm1;
        const double v = fma(fma(fma(k3, sl, k2), sl, k1), sl, k0);
        mn = (v < mn) ? v : mn;  // no NaNs here: plain compares instead of fmin / fmax
        mx = (v > mx) ? v : mx;
    };
    consider(first);
    consider(last);
    const float A = 3.0f * (float)k3, B = 2.0f * (float)k2, C0 = (float)k1;
    const float nanf_ = __int_as_float(0x7fc00000);
    float s1 = nanf_, s2 = nanf_;
    if (fabsf(A) > 1e-6f * (fabsf(B) + fabsf(C0))) {
        const float disc = fmaf(B, B, -4.0f * A * C0);
        if (disc >= 0.0f) {
            const float q = -0.5f * (B + copysignf(sqrtf(disc), B));
            s1 = __fdividef(q, A);
            if (q != 0.0f) s2 = __fdividef(C0, q);
        }
    } else if (B != 0.0f) {
        s1 = __fdividef(-C0, B);
    }
    const float per_piece = c.per_piece, margin = 3.0f * c.inv_per_piece;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
        const float sr = which ? s2 : s1;
        if (sr > -margin && sr < 1.0f + margin) {
            const int f = (int)floorf(((float)sub + sr) * per_piece);
            consider(f - 1);
            consider(f);
            consider(f + 1);
            consider(f + 2);
        }
    }
}

// Renormalisation (wind.py:87-89: iff a sample leaves [0,1]) and the experiment's scale folded into coefficient m of
// curve `curve`: exp 4 / 6 velocity = curve * max_velocity (wind.py:49,62); exp 5: the rect threshold acts on the
// curve itself; second curve of exp 6: curve * pi * 2 (wind.py:63).
__device__ __forceinline__ double fold_coef(const DevCfg &c, int curve, int m, double cfm, double lo, double hi) {
    double off = 0.0, inv = 1.0;
    if (lo < 0.0 || hi > 1.0) { off = lo; inv = 1.0 / (hi - lo); }
    const double scale = curve ? 3.14159265358979323846 * 2.0
                               : ((c.wind_kind == WIND_ANGLE_RECT) ? 1.0 : c.p.max_velocity);
    return ((m == 0) ? (cfm - off) : cfm) * inv * scale;
}

// Called by all 32 lanes with warp-uniform arguments.  On return scratch[m] (m = 0..3) holds
// the folded coefficients of the first drawn curve's piece containing sample `index_next`
// (exp 4/6: velocity, exp 5: the rect source) and scratch[4 + m] those of the second drawn
// curve (exp 6: angle).  The caller must __syncwarp() before the next call.
// Lane map: half-warp h = lane >> 4 works on curve h, sub-lane k = lane & 15 on knot / piece k.
static __device__ __noinline__ void wind_setup_warp(const DevCfg &c, long long env_local, uint32_t episode,
                                                    int index_next, double *scratch) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int fp = c.fp, np = c.npieces, nc = c.ncurves;
    const int cv = lane >> 4, sub = lane & 15;
    double *kn = scratch + kKnotOff, *coef = scratch + kCoefOff;

    // --- knots (wind.py:78): lane (cv, sub) draws knot sub of curve cv -----------------
    if (cv < nc && sub < fp)
        kn[cv * kMaxKnots + sub] = c.ovr_knots ? c.ovr_knots[env_local * 2 * fp + cv * fp + sub]
                                               : episode_knot(c.seed, c.env_id_offset + env_local, episode, cv * fp + sub);
    __syncwarp();

    // --- piece coefficients in the local coordinate s: c_m = sum_k basis[j][m][k] u_k ---
    const int total = nc * np * 4;
    for (int base = 0; base < total; base += 32) {
        const int t = base + lane;
        if (t < total) {
            const int curve = (t >= np * 4) ? 1 : 0;
            const int rem = t - curve * np * 4;              // piece * 4 + m
            coef[t] = coef_dot(c.basis + (size_t)rem * fp, kn + curve * kMaxKnots, fp);
        }
    }
    __syncwarp();

    // --- extremal samples of every piece (np.min / np.max of wind.py:87-89): lane (cv, sub) owns piece sub of curve cv ---
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    double mn = inf, mx = -inf;
    if (cv < nc && sub < np) {
        const double *cf = coef + (cv * np + sub) * 4;
        piece_extremes(c, sub, cf[0], cf[1], cf[2], cf[3], mn, mx);
    }
    // segmented reduction: each half-warp reduces its own curve
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        const double on = __shfl_xor_sync(FULL, mn, o), ox = __shfl_xor_sync(FULL, mx, o);
        mn = (on < mn) ? on : mn;
        mx = (ox > mx) ? ox : mx;
    }
    // lanes 0..15 hold curve A's extremes, 16..31 curve B's; lanes 0..7 do the fold
    const double mnB = __shfl_sync(FULL, mn, 16), mxB = __shfl_sync(FULL, mx, 16);

    // --- fold renormalisation (wind.py:87-89) and the experiment's scale: lanes 0..7 --
    int jn, rn;
    piece_of(c, min(index_next, c.L - 1), jn, rn);
    if (lane < 4 * nc) {
        const int curve = lane >> 2, m = lane & 3;
        const double cfm = coef[(curve * np + jn) * 4 + m];
        scratch[lane] = fold_coef(c, curve, m, cfm, curve ? mnB : mn, curve ? mxB : mx);
    } else if (lane < 8) {
        scratch[lane] = 0.0;
    }
    __syncwarp();
}

// Lane-parallel mapping: ONE ENV PER LANE, no cooperation, no shared memory -- for the kernels whose setup work is
// dense (reset of every env, the compacted episode-end queue of the K > 1 step kernels).  ~30 x fewer warp
// instructions per env than the cooperative mapping (which keeps 14-16 of 32 lanes busy through five
// __syncwarp-separated phases).  out[0..3]: folded coefficients of the first drawn curve's piece containing sample
// index_next, out[4..7]: the second curve's (zeros when the experiment draws fewer curves).  FP = fixed_points as a
// compile-time constant (8: the reference's default, everything in registers) or 0 (any value up to kMaxKnots).
template <int FP>
__device__ __forceinline__ void wind_setup_lane(const DevCfg &c, long long env_local, uint32_t episode, int index_next,
                                                double (&out)[8]) {
    const int fp = FP ? FP : c.fp, np = fp - 1, nc = c.ncurves;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    int jn, rn;
    piece_of(c, min(index_next, c.L - 1), jn, rn);
#pragma unroll
    for (int q = 0; q < 8; ++q) out[q] = 0.0;
    const long long genv = c.env_id_offset + env_local;
#pragma unroll
    for (int curve = 0; curve < 2; ++curve) {
        if (curve >= nc) break;
        double u[FP ? FP : kMaxKnots];
        if (c.ovr_knots) {
#pragma unroll
            for (int k = 0; k < (FP ? FP : kMaxKnots); ++k)
                if (k < fp) u[k] = c.ovr_knots[env_local * 2 * fp + curve * fp + k];
        } else if (FP == 8) {  // knots 8*curve .. 8*curve+7 = Philox blocks 2*curve+1, 2*curve+2 (episode_knot's layout)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const Philox4 r = philox4x32_10((uint32_t)genv, (uint32_t)((unsigned long long)genv >> 32), episode,
                                                kStreamEpisode | (uint32_t)(1 + 2 * curve + b), (uint32_t)c.seed,
                                                (uint32_t)(c.seed >> 32));
                u[4 * b + 0] = knot_from_word(r.x);
                u[4 * b + 1] = knot_from_word(r.y);
                u[4 * b + 2] = knot_from_word(r.z);
                u[4 * b + 3] = knot_from_word(r.w);
            }
        } else {
            for (int k = 0; k < fp; ++k) u[k] = episode_knot(c.seed, genv, episode, curve * fp + k);
        }
        double mn = inf, mx = -inf, sel[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
        for (int j = 0; j < np; ++j) {
            double cf[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) cf[m] = coef_dot(c.basis + (size_t)(j * 4 + m) * fp, u, fp);
            piece_extremes(c, j, cf[0], cf[1], cf[2], cf[3], mn, mx);
            if (j == jn) {
#pragma unroll
                for (int m = 0; m < 4; ++m) sel[m] = cf[m];
            }
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) out[curve * 4 + m] = fold_coef(c, curve, m, sel[m], mn, mx);
    }
}

__device__ __forceinline__ void wind_setup_lane_any(const DevCfg &c, long long env_local, uint32_t episode, int index_next,
                                                    double (&out)[8]) {
    if (c.fp == 8) wind_setup_lane<8>(c, env_local, episode, index_next, out);
    else wind_setup_lane<0>(c, env_local, episode, index_next, out);
}

// np.random.randint draw of an episode (boat_env.py:147-150): only experiment 2 uses it.
__device__ __forceinline__ int episode_start_y(const DevCfg &c, long long env_local, uint32_t episode) {
    if (c.experiment != 2) return 0;  // boat_env.py:166-170: v_y_integrator initial_value = 0 otherwise
    return c.ovr_s_y ? c.ovr_s_y[env_local] : episode_s_y_start(c.seed, c.env_id_offset + env_local, episode, c.s_y_half);
}

}  // namespace boatenv
