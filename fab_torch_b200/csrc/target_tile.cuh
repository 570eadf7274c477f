// Target log-densities and their closed-form gradients, one warp per particle.
//   many-well : fab/target_distributions/many_well.py:81-90 + double_well.py:44-58
//   GMM       : fab/target_distributions/gmm.py:44-66  (MixtureSameFamily of diagonal MVNs,
//               log_prob < -1e4 -> -inf)
// `x` is a row of `d` floats (shared or global); `g` (nullable) receives d log p / d x.
#pragma once
#include "common.cuh"

// One coordinate of the many-well energy and of d log p / d x.  Shared by the warp-per-row form
// below and the streaming kernel in misc_kernels.cuh so both contract to the same FMAs.
__device__ __forceinline__ void manywell_elem(const fab_target_desc& t, float v, int j, float& ej,
                                              float& gj) {
    if ((j & 1) == 0) {     // first coordinate of the pair: a x + b x^2 + c x^4
        const float v2 = v * v;
        ej = t.a * v + t.b * v2 + t.c * (v2 * v2);
        gj = -(t.a + 2.f * t.b * v + 4.f * t.c * (v2 * v));
    } else {                // second coordinate: x^2 / 2
        ej = 0.5f * v * v;
        gj = -v;
    }
}

__device__ __forceinline__ float manywell_row(const fab_target_desc& t, const float* x, float* g,
                                              int d, int lane) {
    float e = 0.f;
    for (int j = lane; j < d; j += 32) {
        float ej, gj;
        manywell_elem(t, x[j], j, ej, gj);
        e += ej;
        if (g) g[j] = gj;
    }
    e = warp_sum(e);
    return -e - t.log_norm;
}

__device__ __forceinline__ float gmm_row(const fab_target_desc& t, const float* x, float* g, int d,
                                         int lane) {
    // pass 1: online logsumexp over the lane's components
    float m = -CUDART_INF_F, s = 0.f;
    for (int k = lane; k < t.n_mixes; k += 32) {
        float q = 0.f, ldet = 0.f;
        for (int j = 0; j < d; ++j) {
            const float sc = __ldg(t.d_scales + (size_t)k * d + j);
            const float u = (x[j] - __ldg(t.d_locs + (size_t)k * d + j)) / sc;
            q += u * u;
            ldet += logf(sc);
        }
        const float c = __ldg(t.d_log_weights + k) - 0.5f * ((float)d * 1.8378770664093453f + q) - ldet;
        if (c > m) { s = s * expf(m - c) + 1.f; m = c; }
        else if (c > -CUDART_INF_F) s += expf(c - m);
        else if (c != c) { s = c; m = c; }           // propagate NaN
    }
    const float M = warp_max(m);
    float part = (m > -CUDART_INF_F) ? s * expf(m - M) : 0.f;
    if (m != m) part = m;
    const float S = warp_sum(part);
    float lp = M + logf(S);
    if (g) {
        for (int j = 0; j < d; ++j) {
            float acc = 0.f;
            for (int k = lane; k < t.n_mixes; k += 32) {
                float q = 0.f, ldet = 0.f;
                for (int jj = 0; jj < d; ++jj) {
                    const float sc = __ldg(t.d_scales + (size_t)k * d + jj);
                    const float u = (x[jj] - __ldg(t.d_locs + (size_t)k * d + jj)) / sc;
                    q += u * u;
                    ldet += logf(sc);
                }
                const float c = __ldg(t.d_log_weights + k) -
                                0.5f * ((float)d * 1.8378770664093453f + q) - ldet;
                const float sc = __ldg(t.d_scales + (size_t)k * d + j);
                const float r = expf(c - lp);
                acc += r * (-(x[j] - __ldg(t.d_locs + (size_t)k * d + j)) / (sc * sc));
            }
            acc = warp_sum(acc);
            if (lane == 0) g[j] = acc;
        }
    }
    if (t.mask_below_1e4 && lp < -1e4f) lp = -CUDART_INF_F;
    return lp - t.log_norm;
}

// ALDP surrogate (include/fab_b200.h): per-coordinate harmonic / periodic terms + one torsion
// coupling; closed-form gradient.
__device__ __forceinline__ float aldp_row(const fab_target_desc& t, const float* x, float* g, int d,
                                          int lane) {
    float e = 0.f;
    const int ia = (int)t.b, ib = (int)t.c;
    for (int j = lane; j < d; j += 32) {
        const float v = x[j];
        const float p0 = __ldg(t.d_scales + j), p1 = __ldg(t.d_locs + j), n = __ldg(t.d_log_weights + j);
        float ej, gj;
        if (n == 0.f) {
            const float u = v - p1;
            ej = 0.5f * p0 * u * u;
            gj = -p0 * u;
        } else {
            float sn, cs;
            sincosf(n * v - p1, &sn, &cs);
            ej = p0 * (1.f - cs);
            gj = -p0 * n * sn;
        }
        if (j == ia || j == ib) {
            float sn, cs;
            sincosf(x[ia] - x[ib], &sn, &cs);
            if (j == ia) { ej += t.a * (1.f - cs); gj -= t.a * sn; }
            else gj += t.a * sn;
        }
        e += ej;
        if (g) g[j] = gj;
    }
    e = warp_sum(e);
    return -e - t.log_norm;
}

// Evaluate the target for the T rows xs[T][ld]; lp_out[T] and gp[T][ld] (nullable) in shared.
__device__ __forceinline__ void target_tile(const fab_target_desc& t, const float* xs, int ld,
                                            int d, int T, float* lp_out, float* gp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int p = warp; p < T; p += FAB_NWARPS) {
        float lp;
        if (t.kind == FAB_TARGET_MANYWELL)
            lp = manywell_row(t, xs + p * ld, gp ? gp + p * ld : nullptr, d, lane);
        else if (t.kind == FAB_TARGET_ALDP_SURROGATE)
            lp = aldp_row(t, xs + p * ld, gp ? gp + p * ld : nullptr, d, lane);
        else
            lp = gmm_row(t, xs + p * ld, gp ? gp + p * ld : nullptr, d, lane);
        if (lane == 0) lp_out[p] = lp;
    }
}
