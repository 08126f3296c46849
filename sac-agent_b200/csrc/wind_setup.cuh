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

// Per-warp scratch: [0..7] the folded result a[4], b[4]; [8..] the coefficient table [curve][piece][4].
constexpr int kScratchDoubles = 8 + 2 * (kMaxKnots - 1) * 4;  // upper bound (static allocations)
__host__ __device__ constexpr int scratch_doubles(int ncurves, int npieces) { return 8 + ncurves * npieces * 4; }

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

__device__ __forceinline__ double eval_sample(const DevCfg &c, const double *coef, int curve, int index) {
    index = max(0, min(index, c.L - 1));
    int j, r;
    piece_of(c, index, j, r);
    const double s = (double)r * c.inv_Lm1;
    const double *cf = coef + (curve * c.npieces + j) * 4;
    return fma(fma(fma(cf[3], s, cf[2]), s, cf[1]), s, cf[0]);
}

// Called by all 32 lanes with warp-uniform arguments.  On return scratch[m] (m = 0..3) holds
// the folded coefficients of the first drawn curve's piece containing sample `index_next`
// (exp 4/6: velocity, exp 5: the rect source) and scratch[4 + m] those of the second drawn
// curve (exp 6: angle).  The caller must __syncwarp() before the next call.
static __device__ __noinline__ void wind_setup_warp(const DevCfg &c, long long env_local, uint32_t episode,
                                                    int index_next, double *scratch) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int fp = c.fp, np = c.npieces, nc = c.ncurves;
    const long long genv = c.env_id_offset + env_local;
    double *coef = scratch + 8;

    // --- knots: one per lane -----------------------------------------------------
    double u = 0.0;
    if (lane < nc * fp)
        u = c.ovr_knots ? c.ovr_knots[env_local * 2 * fp + lane] : episode_knot(c.seed, genv, episode, lane);

    // --- piece coefficients in the local coordinate s: c_m = sum_k basis[j][m][k] u_k --
    const int total = nc * np * 4;
    for (int base = 0; base < total; base += 32) {
        const int t = base + lane;
        const bool valid = t < total;
        const int tt = valid ? t : 0;
        const int curve = (tt >= np * 4) ? 1 : 0;
        const int rem = tt - curve * np * 4;
        const double *row = c.basis + (size_t)rem * fp;  // rem = piece * 4 + m
        double acc = 0.0;
        for (int k = 0; k < fp; ++k) {
            const double uk = __shfl_sync(FULL, u, curve * fp + k);
            acc = fma(__ldg(row + k), uk, acc);
        }
        if (valid) coef[t] = acc;
    }
    __syncwarp();

    // --- extremal samples of every piece -------------------------------------------
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    double mn = inf, mx = -inf;
    const int my_curve = (lane >= np) ? 1 : 0;
    if (lane < nc * np) {
        const int j = lane - my_curve * np;
        const double *cf = coef + (my_curve * np + j) * 4;
        const double c1 = cf[1], c2 = cf[2], c3 = cf[3];
        auto consider = [&](int index) {
            const double v = eval_sample(c, coef, my_curve, index);
            mn = fmin(mn, v);
            mx = fmax(mx, v);
        };
        // first / last sample that falls into this piece
        consider((j * c.Lm1 + np - 1) / np);
        consider(((j + 1) * c.Lm1) / np);
        // roots of the derivative c1 + 2 c2 s + 3 c3 s^2
        const double A = 3.0 * c3, B = 2.0 * c2, C0 = c1;
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        double s1 = nan, s2 = nan;
        if (fabs(A) > 1e-14 * (fabs(B) + fabs(C0))) {
            const double disc = B * B - 4.0 * A * C0;
            if (disc >= 0.0) {
                const double q = -0.5 * (B + copysign(sqrt(disc), B));
                s1 = q / A;
                if (q != 0.0) s2 = C0 / q;
            }
        } else if (B != 0.0) {
            s1 = -C0 / B;
        }
        const double per_piece = (double)c.Lm1 / (double)np;  // samples per piece
        const double margin = 2.0 / per_piece;
        if (s1 > -margin && s1 < 1.0 + margin) {
            const int f = (int)floor(((double)j + s1) * per_piece);
            consider(f);
            consider(f + 1);
        }
        if (s2 > -margin && s2 < 1.0 + margin) {
            const int f = (int)floor(((double)j + s2) * per_piece);
            consider(f);
            consider(f + 1);
        }
    }
    const bool inA = lane < np, inB = (lane >= np) && (lane < 2 * np) && nc == 2;
    const double mnA = warp_min(inA ? mn : inf), mxA = warp_max(inA ? mx : -inf);
    double mnB = 0.0, mxB = 1.0;
    if (nc == 2) {
        mnB = warp_min(inB ? mn : inf);
        mxB = warp_max(inB ? mx : -inf);
    }

    // --- fold renormalisation (wind.py:87-89) and the experiment's scale: lanes 0..7 --
    int jn, rn;
    piece_of(c, min(index_next, c.L - 1), jn, rn);
    if (lane < 4 * nc) {
        const int curve = lane >> 2, m = lane & 3;
        const double cf = coef[(curve * np + jn) * 4 + m];
        const double lo = curve ? mnB : mnA, hi = curve ? mxB : mxA;
        double off = 0.0, inv = 1.0;
        if (lo < 0.0 || hi > 1.0) { off = lo; inv = 1.0 / (hi - lo); }
        // exp 4 / 6: curve * max_velocity (wind.py:49,62); exp 5: the rect threshold acts on the curve
        // itself; second curve of exp 6: curve * pi * 2 (wind.py:63)
        const double scale = curve ? 3.14159265358979323846 * 2.0
                                   : ((c.wind_kind == WIND_ANGLE_RECT) ? 1.0 : c.p.max_velocity);
        scratch[lane] = ((m == 0) ? (cf - off) : cf) * inv * scale;
    } else if (lane < 8) {
        scratch[lane] = 0.0;
    }
    __syncwarp();
}

// np.random.randint draw of an episode (boat_env.py:147-150): only experiment 2 uses it.
__device__ __forceinline__ int episode_start_y(const DevCfg &c, long long env_local, uint32_t episode) {
    if (c.experiment != 2) return 0;  // boat_env.py:166-170: v_y_integrator initial_value = 0 otherwise
    return c.ovr_s_y ? c.ovr_s_y[env_local] : episode_s_y_start(c.seed, c.env_id_offset + env_local, episode, c.s_y_half);
}

}  // namespace boatenv
