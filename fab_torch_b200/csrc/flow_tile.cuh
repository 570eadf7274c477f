// RealNVP evaluation for the T particles of one CTA, entirely in shared memory.
//
//   flow_inverse  : x -> z, log q(x)             normflows NormalizingFlow.log_prob  (K3)
//   flow_backward : d log q / d x                analytic reverse sweep, replaces
//                                                torch.autograd.grad in base.py:50-56 (K4)
//   flow_sample   : eps -> x, log q(x)           normflows NormalizingFlow.sample     (K1,K2)
//
// Layer semantics follow SURVEY Appendix B / oracle/realnvp.py (the restated normflows):
//   sampling, layer k:  [v1,v2] = u;  (shift,scale) = MLP(v1);  y2 = v2*exp(scale)+shift;
//                       u' = [v1,y2] @ Wmix^-1;  log q -= sum(scale) - sum(log_S)
//   inverse,  layer k:  v = u @ Wmix; (shift,scale) = MLP(v1);  y2 = (v2-shift)*exp(-scale);
//                       u' = [v1,y2];           log q += sum(log_S) - sum(scale)
#pragma once
#include "tile_gemm.cuh"

struct TileBufs {
    float *zs, *vs, *z1b, *par, *h1, *h2, *red, *sy2, *ses, *ld;
    uint32_t *m1, *m2;
};

__device__ __forceinline__ TileBufs tile_bufs(const TileLayout& L, float* smem) {
    TileBufs b;
    b.zs = smem + L.o_zs;   b.vs = smem + L.o_vs;   b.z1b = smem + L.o_z1b;
    b.par = smem + L.o_par; b.h1 = smem + L.o_h1;   b.h2 = smem + L.o_h2;
    b.red = smem + L.o_red; b.sy2 = smem + L.o_sy2; b.ses = smem + L.o_ses;
    b.ld = smem + L.o_ld;
    b.m1 = reinterpret_cast<uint32_t*>(smem + L.o_m1);
    b.m2 = reinterpret_cast<uint32_t*>(smem + L.o_m2);
    return b;
}

// Zero the padding-sensitive buffers once per kernel (pads must stay exactly 0 because the packed
// operands multiply them by 0 and 0*inf would poison a row).
__device__ __forceinline__ void tile_zero_pads(const TileLayout& L, const TileBufs& b) {
    for (int i = threadIdx.x; i < L.T * L.DP; i += FAB_NT) { b.zs[i] = 0.f; b.vs[i] = 0.f; }
    for (int i = threadIdx.x; i < L.T * L.D1P; i += FAB_NT) b.z1b[i] = 0.f;
    for (int i = threadIdx.x; i < L.T * L.P2; i += FAB_NT) b.par[i] = 0.f;
    for (int i = threadIdx.x; i < L.T * L.WP; i += FAB_NT) { b.h1[i] = 0.f; b.h2[i] = 0.f; }
}

// hidden-layer epilogue: h = relu(sum + bias) (FWD) or h = mask ? sum : 0 (BWD), one 32-column
// word of one particle per warp iteration so the ReLU mask is a single ballot.
template <int T, bool FWD, bool SAVE>
__device__ __forceinline__ void hidden_epilogue(const TileLayout& L, const float* red, int KS,
                                                const float* __restrict__ bias, float* h,
                                                uint32_t* mask) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int w = warp; w < T * L.WW; w += FAB_NWARPS) {
        const int p = w / L.WW;
        const int n = (w - p * L.WW) * 32 + lane;
        const bool in = n < L.WP;
        float v = 0.f;
        if (in) v = red_sum<T>(red, KS, L.WP, p, n);
        if (FWD) {
            if (in) v += __ldg(bias + n);
            const bool pos = in && (v > 0.f);
            if (SAVE) {
                const uint32_t bits = __ballot_sync(FAB_FULL, pos);
                if (lane == 0) mask[w] = bits;
            }
            if (in) h[p * L.WP + n] = pos ? v : 0.f;
        } else {
            const uint32_t bits = mask[w];
            if (in) h[p * L.WP + n] = ((bits >> lane) & 1u) ? v : 0.f;
        }
    }
}

// The conditioner MLP on z1b -> red holds the (shift|scale) partial sums (NP = P2).
template <int T, bool SAVE>
__device__ __forceinline__ GemmSplit conditioner_forward(const TileLayout& L, const TileBufs& b,
                                                         const float* __restrict__ lay,
                                                         const fab_flow_desc& f, int k) {
    GemmSplit g = tile_gemm<T>(b.z1b, L.D1P, L.D1P / 4,
                               reinterpret_cast<const float4*>(lay + f.o_w1), L.WP, b.red,
                               L.red_floats);
    __syncthreads();
    hidden_epilogue<T, true, SAVE>(L, b.red, g.KS, lay + f.o_b1, b.h1,
                                   SAVE ? b.m1 + (size_t)k * T * L.WW : nullptr);
    __syncthreads();
    g = tile_gemm<T>(b.h1, L.WP, L.WP / 4, reinterpret_cast<const float4*>(lay + f.o_w2), L.WP,
                     b.red, L.red_floats);
    __syncthreads();
    hidden_epilogue<T, true, SAVE>(L, b.red, g.KS, lay + f.o_b2, b.h2,
                                   SAVE ? b.m2 + (size_t)k * T * L.WW : nullptr);
    __syncthreads();
    g = tile_gemm<T>(b.h2, L.WP, L.WP / 4, reinterpret_cast<const float4*>(lay + f.o_w3), L.P2,
                     b.red, L.red_floats);
    __syncthreads();
    return g;
}

// per-particle  ld[p] += add - sum_j par[p][d2 + j]   (par holds [shift | scale] after coupling)
template <int T>
__device__ __forceinline__ void logdet_accumulate(const TileLayout& L, const TileBufs& b,
                                                  float add, float sign) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int p = warp; p < T; p += FAB_NWARPS) {
        float s = 0.f;
        for (int j = lane; j < L.d2; j += 32) s += b.par[p * L.P2 + L.d2 + j];
        s = warp_sum(s);
        if (lane == 0) b.ld[p] += add + sign * s;
    }
}

// x in b.zs  ->  z in b.zs, log q in lq_out[p] (shared).  With SAVE the per-layer
// (y2, exp(-scale), ReLU masks) needed by flow_backward are kept and b.vs is set to
// d log N(z) / dz.
template <int T, bool SAVE>
__device__ void flow_inverse(const TileLayout& L, const TileBufs& b, const fab_flow_desc& f,
                             const float* __restrict__ blob, float* lq_out) {
    for (int p = threadIdx.x; p < T; p += FAB_NT) b.ld[p] = 0.f;
    __syncthreads();
    for (int k = L.K - 1; k >= 0; --k) {
        const float* lay = blob + f.off_layers + (size_t)k * f.layer_stride;
        // v = z @ Wmix
        GemmSplit g = tile_gemm<T>(b.zs, L.DP, L.DP / 4,
                                   reinterpret_cast<const float4*>(lay + f.o_mix), L.DP, b.red,
                                   L.red_floats);
        __syncthreads();
        for (int i = threadIdx.x; i < T * L.DP; i += FAB_NT) {
            const int p = i / L.DP, n = i - p * L.DP;
            if (n < L.d) {
                const float v = red_sum<T>(b.red, g.KS, L.DP, p, n);
                b.vs[i] = v;
                if (n < L.d1) { b.z1b[p * L.D1P + n] = v; b.zs[i] = v; }
            }
        }
        __syncthreads();
        g = conditioner_forward<T, SAVE>(L, b, lay, f, k);
        // coupling inverse: y2 = (v2 - shift) * exp(-scale)
        for (int i = threadIdx.x; i < T * L.d2; i += FAB_NT) {
            const int p = i / L.d2, j = i - p * L.d2;
            const float shift = red_sum<T>(b.red, g.KS, L.P2, p, j) + __ldg(lay + f.o_b3 + j);
            const float scale = red_sum<T>(b.red, g.KS, L.P2, p, L.d2 + j) +
                                __ldg(lay + f.o_b3 + L.d2 + j);
            const float es = expf(-scale);
            const float y2 = (b.vs[p * L.DP + L.d1 + j] - shift) * es;
            b.zs[p * L.DP + L.d1 + j] = y2;
            b.par[p * L.P2 + L.d2 + j] = scale;
            if (SAVE) {
                b.sy2[((size_t)k * T + p) * L.d2 + j] = y2;
                b.ses[((size_t)k * T + p) * L.d2 + j] = es;
            }
        }
        __syncthreads();
        logdet_accumulate<T>(L, b, __ldg(lay + f.o_logs), -1.f);
        // (the next phase that touches par/ld is at least one barrier away)
    }
    __syncthreads();
    // base Gaussian: log N(z; loc, exp(log_scale)) and its z-gradient
    {
        const float* loc = blob + f.off_base_loc;
        const float* lsc = blob + f.off_base_log_scale;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int p = warp; p < T; p += FAB_NWARPS) {
            float s = 0.f;
            for (int j = lane; j < L.d; j += 32) {
                const float ls = __ldg(lsc + j);
                const float inv = expf(-ls);
                const float u = (b.zs[p * L.DP + j] - __ldg(loc + j)) * inv;
                s += ls + 0.5f * u * u;
                if (SAVE) b.vs[p * L.DP + j] = -u * inv;
            }
            s = warp_sum(s);
            if (lane == 0)
                lq_out[p] = b.ld[p] + (-0.5f * (float)L.d * 1.8378770664093453f - s);
        }
    }
    __syncthreads();
}

// b.vs holds d log q / d z on entry (set by flow_inverse<SAVE=true>) and d log q / d x on exit.
template <int T>
__device__ void flow_backward(const TileLayout& L, const TileBufs& b, const fab_flow_desc& f,
                              const float* __restrict__ blob) {
    float* gs = b.vs;
    for (int k = 0; k < L.K; ++k) {
        const float* lay = blob + f.off_layers + (size_t)k * f.layer_stride;
        // coupling backward: gv2 = g2*es, gshift = -gv2, gscale = -g2*y2 - 1
        for (int i = threadIdx.x; i < T * L.d2; i += FAB_NT) {
            const int p = i / L.d2, j = i - p * L.d2;
            const float es = b.ses[((size_t)k * T + p) * L.d2 + j];
            const float y2 = b.sy2[((size_t)k * T + p) * L.d2 + j];
            const float g2 = gs[p * L.DP + L.d1 + j];
            const float gv2 = g2 * es;
            b.par[p * L.P2 + j] = -gv2;
            b.par[p * L.P2 + L.d2 + j] = -g2 * y2 - 1.0f;
            gs[p * L.DP + L.d1 + j] = gv2;
        }
        __syncthreads();
        // gh2 = (gparam @ W3) * m2
        GemmSplit g = tile_gemm<T>(b.par, L.P2, L.P2 / 4,
                                   reinterpret_cast<const float4*>(lay + f.o_w3t), L.WP, b.red,
                                   L.red_floats);
        __syncthreads();
        hidden_epilogue<T, false, false>(L, b.red, g.KS, nullptr, b.h2,
                                         b.m2 + (size_t)k * T * L.WW);
        __syncthreads();
        // gh1 = (gh2 @ W2) * m1
        g = tile_gemm<T>(b.h2, L.WP, L.WP / 4, reinterpret_cast<const float4*>(lay + f.o_w2t),
                         L.WP, b.red, L.red_floats);
        __syncthreads();
        hidden_epilogue<T, false, false>(L, b.red, g.KS, nullptr, b.h1,
                                         b.m1 + (size_t)k * T * L.WW);
        __syncthreads();
        // gv1 = g1 + gh1 @ W1
        g = tile_gemm<T>(b.h1, L.WP, L.WP / 4, reinterpret_cast<const float4*>(lay + f.o_w1t),
                         L.D1P, b.red, L.red_floats);
        __syncthreads();
        for (int i = threadIdx.x; i < T * L.d1; i += FAB_NT) {
            const int p = i / L.d1, n = i - p * L.d1;
            gs[p * L.DP + n] += red_sum<T>(b.red, g.KS, L.D1P, p, n);
        }
        __syncthreads();
        // g_u = gv @ Wmix^T
        g = tile_gemm<T>(gs, L.DP, L.DP / 4, reinterpret_cast<const float4*>(lay + f.o_mix_t),
                         L.DP, b.red, L.red_floats);
        __syncthreads();
        for (int i = threadIdx.x; i < T * L.DP; i += FAB_NT) {
            const int p = i / L.DP, n = i - p * L.DP;
            if (n < L.d) gs[i] = red_sum<T>(b.red, g.KS, L.DP, p, n);
        }
        __syncthreads();
    }
    // par is used as a zero-padded GEMM operand only inside this function and as scratch in
    // flow_inverse/flow_sample (columns < 2*d2), so its pad columns are still 0.
}

// eps in b.zs -> x in b.zs, forward-pass log q in lq_out[p].
template <int T>
__device__ void flow_sample(const TileLayout& L, const TileBufs& b, const fab_flow_desc& f,
                            const float* __restrict__ blob, float* lq_out) {
    {   // base: z = loc + exp(log_scale)*eps ; log p0 = -d/2 log 2pi - sum(log_scale + eps^2/2)
        const float* loc = blob + f.off_base_loc;
        const float* lsc = blob + f.off_base_log_scale;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int p = warp; p < T; p += FAB_NWARPS) {
            float s = 0.f;
            for (int j = lane; j < L.d; j += 32) {
                const float ls = __ldg(lsc + j);
                const float e = b.zs[p * L.DP + j];
                s += ls + 0.5f * e * e;
                b.zs[p * L.DP + j] = __ldg(loc + j) + expf(ls) * e;
            }
            s = warp_sum(s);
            if (lane == 0) b.ld[p] = -0.5f * (float)L.d * 1.8378770664093453f - s;
        }
    }
    __syncthreads();
    for (int k = 0; k < L.K; ++k) {
        const float* lay = blob + f.off_layers + (size_t)k * f.layer_stride;
        for (int i = threadIdx.x; i < T * L.d1; i += FAB_NT) {
            const int p = i / L.d1, n = i - p * L.d1;
            b.z1b[p * L.D1P + n] = b.zs[p * L.DP + n];
        }
        __syncthreads();
        GemmSplit g = conditioner_forward<T, false>(L, b, lay, f, k);
        for (int i = threadIdx.x; i < T * L.d2; i += FAB_NT) {
            const int p = i / L.d2, j = i - p * L.d2;
            const float shift = red_sum<T>(b.red, g.KS, L.P2, p, j) + __ldg(lay + f.o_b3 + j);
            const float scale = red_sum<T>(b.red, g.KS, L.P2, p, L.d2 + j) +
                                __ldg(lay + f.o_b3 + L.d2 + j);
            b.zs[p * L.DP + L.d1 + j] = b.zs[p * L.DP + L.d1 + j] * expf(scale) + shift;
            b.par[p * L.P2 + L.d2 + j] = scale;
        }
        __syncthreads();
        // log q -= sum(scale);  log q -= (-sum log_S)
        logdet_accumulate<T>(L, b, __ldg(lay + f.o_logs), -1.f);
        // u' = [v1,y2] @ Wmix^-1
        g = tile_gemm<T>(b.zs, L.DP, L.DP / 4, reinterpret_cast<const float4*>(lay + f.o_mix_inv),
                         L.DP, b.red, L.red_floats);
        __syncthreads();
        for (int i = threadIdx.x; i < T * L.DP; i += FAB_NT) {
            const int p = i / L.DP, n = i - p * L.DP;
            if (n < L.d) b.zs[i] = red_sum<T>(b.red, g.KS, L.DP, p, n);
        }
        __syncthreads();
    }
    for (int p = threadIdx.x; p < T; p += FAB_NT) lq_out[p] = b.ld[p];
    __syncthreads();
}
