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
// Per layer and direction this is three GEMMs (include/fab_b200.h: o_mw1/o_w2/o_w3 and
// o_w3t/o_w2t/o_w1mt): the mixing matrix is merged into the neighbouring MLP GEMM and biases ride
// along as an extra K row against a constant-one activation row.
//
// zs, vs, z1b, par, h1, h2 are GEMM operands in the k-major / particle-fastest layout
// (tile_gemm.cuh): element (p, n) of any of them sits at float index kidx<T>(p, n) = n*TP + p.
#pragma once
#include "tile_gemm.cuh"

struct TileBufs {
    float *zs, *vs, *z1b, *par, *h1, *h2, *red, *sy2, *ses, *ld, *scl;
    float *loc, *lsc, *inv, *logs;          // staged constants
    uint32_t *m1, *m2;
};

__device__ __forceinline__ TileBufs tile_bufs(const TileLayout& L, float* smem) {
    TileBufs b;
    b.zs = smem + L.o_zs;   b.vs = smem + L.o_vs;   b.z1b = smem + L.o_z1b;
    b.par = smem + L.o_par; b.h1 = smem + L.o_h1;   b.h2 = smem + L.o_h2;
    b.red = smem + L.o_red; b.sy2 = smem + L.o_sy2; b.ses = smem + L.o_ses;
    b.ld = smem + L.o_ld;   b.scl = smem + L.o_scl;
    b.loc = smem + L.o_const; b.lsc = b.loc + L.DP; b.inv = b.lsc + L.DP; b.logs = b.inv + L.DP;
    b.m1 = reinterpret_cast<uint32_t*>(smem + L.o_m1);
    b.m2 = reinterpret_cast<uint32_t*>(smem + L.o_m2);
    return b;
}

// rows [row, row+4) of an operand buffer := (1,0,0,0) pattern: the constant activation that
// multiplies the bias row of an operand
template <int T>
__device__ __forceinline__ void set_one_rows(float* buf, int row) {
    constexpr int TP = TileDims<T>::TP;
    for (int i = threadIdx.x; i < 4 * TP; i += FAB_NT) buf[(size_t)row * TP + i] = i < TP ? 1.f : 0.f;
}

// Once per kernel: zero the operand buffers (pad rows/columns must stay exactly 0 because the
// packed weights multiply them by 0 and 0*inf would poison a row), set the bias rows, and stage
// the small per-flow constants (base loc / log_scale, per-layer sum(log_S)) in shared memory so
// that no serial code path waits on an L2 round trip.
template <int T>
__device__ __forceinline__ void tile_init(const TileLayout& L, const TileBufs& b,
                                          const fab_flow_desc& f, const float* __restrict__ blob) {
    constexpr int TP = TileDims<T>::TP;
    for (int i = threadIdx.x; i < TP * (L.DP + 4); i += FAB_NT) b.zs[i] = 0.f;
    for (int i = threadIdx.x; i < TP * L.DP; i += FAB_NT) b.vs[i] = 0.f;
    for (int i = threadIdx.x; i < TP * (L.D1P + 4); i += FAB_NT) b.z1b[i] = 0.f;
    for (int i = threadIdx.x; i < TP * L.P2; i += FAB_NT) b.par[i] = 0.f;
    for (int i = threadIdx.x; i < TP * (L.WP + (L.DP > 4 ? L.DP : 4)); i += FAB_NT) b.h1[i] = 0.f;
    for (int i = threadIdx.x; i < TP * (L.WP + 4); i += FAB_NT) b.h2[i] = 0.f;
    for (int j = threadIdx.x; j < L.DP; j += FAB_NT) {
        const float loc = j < L.d ? __ldg(blob + f.off_base_loc + j) : 0.f;
        const float ls = j < L.d ? __ldg(blob + f.off_base_log_scale + j) : 0.f;
        b.loc[j] = loc; b.lsc[j] = ls; b.inv[j] = expf(-ls);
    }
    for (int k = threadIdx.x; k < L.K; k += FAB_NT)
        b.logs[k] = __ldg(blob + f.off_layers + (size_t)k * f.layer_stride + f.o_logs);
    __syncthreads();
    set_one_rows<T>(b.zs, L.DP);
    set_one_rows<T>(b.z1b, L.D1P);
    set_one_rows<T>(b.h1, L.WP);
    set_one_rows<T>(b.h2, L.WP);
    __syncthreads();
}

// hidden-layer epilogue; one quad (4 consecutive particles of one column) per lane and step:
// FWD: h = relu(sum)  (bias already inside the GEMM), ReLU masks = 4 ballots per 32 quads;
// BWD: h = mask ? sum : 0.   `col0` = first column of this block inside the GEMM output.
template <int T, bool FWD, bool SAVE>
__device__ __forceinline__ void hidden_epilogue(const TileLayout& L, const float* red, int KS,
                                                int NP, int col0, float* h, uint32_t* mask) {
    constexpr int TP = TileDims<T>::TP;
    // work item q = (column quad n4, particle p), particle fastest: the KS partial rows are read
    // with one LDS.128 each (conflict-free: rows are NP+4 floats apart) and the four results go
    // to h[4*n4+i][p] (consecutive lanes -> consecutive floats).
    const int lane = threadIdx.x & 31;
    const int nq = (L.WP >> 2) * T;
    const int NPs = NP + 4;
    const int stride = T * NPs;
    for (int q0 = (threadIdx.x & ~31); q0 < nq; q0 += FAB_NT) {
        const int q = q0 + lane;
        const bool in = q < nq;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        int n4 = 0, p = 0;
        if (in) {
            n4 = q / T;
            p = q - n4 * T;
            const float* r = red + p * NPs + col0 + (n4 << 2);
            v = *reinterpret_cast<const float4*>(r);
#pragma unroll 4
            for (int ks = 1; ks < KS; ++ks) {
                r += stride;
                const float4 u = *reinterpret_cast<const float4*>(r);
                v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
            }
        }
        uint32_t* mw = mask + ((q0 >> 5) << 2);
        float* hp = h + (size_t)(n4 << 2) * TP + p;
        if (FWD) {
            const bool px = in && v.x > 0.f, py = in && v.y > 0.f, pz = in && v.z > 0.f,
                       pw = in && v.w > 0.f;
            if (SAVE) {
                const uint32_t bx = __ballot_sync(FAB_FULL, px), by = __ballot_sync(FAB_FULL, py),
                               bz = __ballot_sync(FAB_FULL, pz), bw = __ballot_sync(FAB_FULL, pw);
                if (lane == 0) *reinterpret_cast<uint4*>(mw) = make_uint4(bx, by, bz, bw);
            }
            if (in) {
                hp[0] = px ? v.x : 0.f; hp[TP] = py ? v.y : 0.f;
                hp[2 * TP] = pz ? v.z : 0.f; hp[3 * TP] = pw ? v.w : 0.f;
            }
        } else {
            const uint4 bits = *reinterpret_cast<const uint4*>(mw);
            if (in) {
                hp[0] = (bits.x >> lane) & 1u ? v.x : 0.f; hp[TP] = (bits.y >> lane) & 1u ? v.y : 0.f;
                hp[2 * TP] = (bits.z >> lane) & 1u ? v.z : 0.f; hp[3 * TP] = (bits.w >> lane) & 1u ? v.w : 0.f;
            }
        }
    }
}

// MLP stages 2 and 3 (h1 -> h2 -> [shift|scale] partial sums in red, NP = P2); returns KS.
// `next_*` describe the GEMM that follows the coupling step (prefetched behind the last barrier).
template <int T, bool SAVE>
__device__ __forceinline__ int mlp_tail(const TileLayout& L, const TileBufs& b,
                                        const float* __restrict__ lay, const fab_flow_desc& f,
                                        int k, const float* next_wp, int next_K4, int next_NP) {
    int KS = tile_gemm<T>(b.h1, L.WP / 4 + 1, reinterpret_cast<const float4*>(lay + f.o_w2), L.WP,
                          b.red, L.red_floats);
    tile_gemm_prefetch<T>(L.WP / 4 + 1, reinterpret_cast<const float4*>(lay + f.o_w3), L.P2,
                          L.red_floats);
    __syncthreads();
    hidden_epilogue<T, true, SAVE>(L, b.red, KS, L.WP, 0, b.h2,
                                   SAVE ? b.m2 + (size_t)k * L.MW : nullptr);
    __syncthreads();
    KS = tile_gemm<T>(b.h2, L.WP / 4 + 1, reinterpret_cast<const float4*>(lay + f.o_w3), L.P2,
                      b.red, L.red_floats);
    if (next_wp)
        tile_gemm_prefetch<T>(next_K4, reinterpret_cast<const float4*>(next_wp), next_NP,
                              L.red_floats);
    __syncthreads();
    return KS;
}

// per-particle  ld[p] += add - sum_j scl[p][j]
template <int T>
__device__ __forceinline__ void logdet_accumulate(const TileLayout& L, const TileBufs& b, float add) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int p = warp; p < T; p += FAB_NWARPS) {
        float s = 0.f;
        for (int j = lane; j < L.d2; j += 32) s += b.scl[p * L.d2 + j];
        s = warp_sum(s);
        if (lane == 0) b.ld[p] += add - s;
    }
}

// x in b.zs  ->  z in b.zs, log q in lq_out[p] (shared).  With SAVE the per-layer
// (y2, exp(-scale), ReLU masks) needed by flow_backward are kept and b.vs is set to
// d log N(z) / dz.
template <int T, bool SAVE>
__device__ void flow_inverse(const TileLayout& L, const TileBufs& b, const fab_flow_desc& f,
                             const float* __restrict__ blob, float* lq_out) {
    constexpr int TP = TileDims<T>::TP;
    for (int p = threadIdx.x; p < T; p += FAB_NT) b.ld[p] = 0.f;
    const int NP1 = L.DP + L.WP;
    for (int k = L.K - 1; k >= 0; --k) {
        const float* lay = blob + f.off_layers + (size_t)k * f.layer_stride;
        // [v | h1pre] = [z | 1] @ [Wmix | Wmix[:, :d1] W1^T ; 0 | b1]
        int KS = tile_gemm<T>(b.zs, L.DP / 4 + 1, reinterpret_cast<const float4*>(lay + f.o_mw1),
                              NP1, b.red, L.red_floats);
        tile_gemm_prefetch<T>(L.WP / 4 + 1, reinterpret_cast<const float4*>(lay + f.o_w2), L.WP,
                              L.red_floats);
        __syncthreads();
        for (int e = threadIdx.x; e < TP * L.d; e += FAB_NT) {
            int p, n;
            kdecode<T>(e, p, n);
            if (p < T) {
                const float v = red_sum<T>(b.red, KS, NP1, p, n);
                b.vs[e] = v;
                if (n < L.d1) b.zs[e] = v;
            }
        }
        hidden_epilogue<T, true, SAVE>(L, b.red, KS, NP1, L.DP, b.h1,
                                       SAVE ? b.m1 + (size_t)k * L.MW : nullptr);
        if (SAVE) set_one_rows<T>(b.h1, L.WP);        // flow_backward parks [gv] in these rows
        __syncthreads();
        KS = mlp_tail<T, SAVE>(L, b, lay, f, k, k > 0 ? lay - f.layer_stride + f.o_mw1 : nullptr,
                               L.DP / 4 + 1, NP1);
        // coupling inverse: y2 = (v2 - shift) * exp(-scale)
        for (int i = threadIdx.x; i < T * L.d2; i += FAB_NT) {
            const int j = i / T, p = i - j * T;
            const float shift = red_sum<T>(b.red, KS, L.P2, p, j);
            const float scale = red_sum<T>(b.red, KS, L.P2, p, L.d2 + j);
            const float es = expf(-scale);
            const int e = kidx<T>(p, L.d1 + j);
            const float y2 = (b.vs[e] - shift) * es;
            b.zs[e] = y2;
            b.scl[p * L.d2 + j] = scale;
            if (SAVE) {
                b.sy2[((size_t)k * T + p) * L.d2 + j] = y2;
                b.ses[((size_t)k * T + p) * L.d2 + j] = es;
            }
        }
        __syncthreads();
        logdet_accumulate<T>(L, b, b.logs[k]);
        // (scl / ld are next touched after at least one more barrier)
    }
    __syncthreads();
    // base Gaussian: log N(z; loc, exp(log_scale)) and its z-gradient
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int p = warp; p < T; p += FAB_NWARPS) {
            float s = 0.f;
            for (int j = lane; j < L.d; j += 32) {
                const float inv = b.inv[j];
                const int e = kidx<T>(p, j);
                const float u = (b.zs[e] - b.loc[j]) * inv;
                s += b.lsc[j] + 0.5f * u * u;
                if (SAVE) b.vs[e] = -u * inv;
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
    constexpr int TP = TileDims<T>::TP;
    float* gs = b.vs;
    float* gv = b.h1 + (size_t)L.WP * TP;               // [gv] rows behind gh1
    for (int k = 0; k < L.K; ++k) {
        const float* lay = blob + f.off_layers + (size_t)k * f.layer_stride;
        // coupling backward: gv2 = g2*es, gshift = -gv2, gscale = -g2*y2 - 1
        for (int i = threadIdx.x; i < T * L.d2; i += FAB_NT) {
            const int j = i / T, p = i - j * T;
            const float es = b.ses[((size_t)k * T + p) * L.d2 + j];
            const float y2 = b.sy2[((size_t)k * T + p) * L.d2 + j];
            const int e = kidx<T>(p, L.d1 + j);
            const float g2 = gs[e];
            const float gv2 = g2 * es;
            b.par[kidx<T>(p, j)] = -gv2;
            b.par[kidx<T>(p, L.d2 + j)] = -g2 * y2 - 1.0f;
            gv[e] = gv2;
        }
        for (int e = threadIdx.x; e < TP * L.DP; e += FAB_NT) {
            int p, n;
            kdecode<T>(e, p, n);
            if (n < L.d1) gv[e] = p < T ? gs[e] : 0.f;
            else if (n >= L.d) gv[e] = 0.f;
        }
        __syncthreads();
        // gh2 = (gparam @ W3) * m2
        int KS = tile_gemm<T>(b.par, L.P2 / 4, reinterpret_cast<const float4*>(lay + f.o_w3t), L.WP,
                              b.red, L.red_floats);
        tile_gemm_prefetch<T>(L.WP / 4, reinterpret_cast<const float4*>(lay + f.o_w2t), L.WP,
                              L.red_floats);
        __syncthreads();
        hidden_epilogue<T, false, false>(L, b.red, KS, L.WP, 0, b.h2, b.m2 + (size_t)k * L.MW);
        __syncthreads();
        // gh1 = (gh2 @ W2) * m1
        KS = tile_gemm<T>(b.h2, L.WP / 4, reinterpret_cast<const float4*>(lay + f.o_w2t), L.WP,
                          b.red, L.red_floats);
        tile_gemm_prefetch<T>((L.WP + L.DP) / 4, reinterpret_cast<const float4*>(lay + f.o_w1mt),
                              L.DP, L.red_floats);
        __syncthreads();
        hidden_epilogue<T, false, false>(L, b.red, KS, L.WP, 0, b.h1, b.m1 + (size_t)k * L.MW);
        __syncthreads();
        // g_u = [gh1 | gv] @ [W1 Wmix[:, :d1]^T ; Wmix^T]
        KS = tile_gemm<T>(b.h1, (L.WP + L.DP) / 4, reinterpret_cast<const float4*>(lay + f.o_w1mt),
                          L.DP, b.red, L.red_floats);
        if (k + 1 < L.K)
            tile_gemm_prefetch<T>(L.P2 / 4,
                                  reinterpret_cast<const float4*>(lay + f.layer_stride + f.o_w3t),
                                  L.WP, L.red_floats);
        __syncthreads();
        for (int e = threadIdx.x; e < TP * L.d; e += FAB_NT) {
            int p, n;
            kdecode<T>(e, p, n);
            if (p < T) gs[e] = red_sum<T>(b.red, KS, L.DP, p, n);
        }
        __syncthreads();
    }
}

// eps in b.zs -> x in b.zs, forward-pass log q in lq_out[p].
template <int T>
__device__ void flow_sample(const TileLayout& L, const TileBufs& b, const fab_flow_desc& f,
                            const float* __restrict__ blob, float* lq_out) {
    constexpr int TP = TileDims<T>::TP;
    {   // base: z = loc + exp(log_scale)*eps ; log p0 = -d/2 log 2pi - sum(log_scale + eps^2/2)
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int p = warp; p < T; p += FAB_NWARPS) {
            float s = 0.f;
            for (int j = lane; j < L.d; j += 32) {
                const float ls = b.lsc[j];
                const int e = kidx<T>(p, j);
                const float ev = b.zs[e];
                s += ls + 0.5f * ev * ev;
                b.zs[e] = b.loc[j] + expf(ls) * ev;
            }
            s = warp_sum(s);
            if (lane == 0) b.ld[p] = -0.5f * (float)L.d * 1.8378770664093453f - s;
        }
    }
    set_one_rows<T>(b.h1, L.WP);
    __syncthreads();
    for (int k = 0; k < L.K; ++k) {
        const float* lay = blob + f.off_layers + (size_t)k * f.layer_stride;
        for (int e = threadIdx.x; e < TP * L.d1; e += FAB_NT) b.z1b[e] = b.zs[e];
        __syncthreads();
        int KS = tile_gemm<T>(b.z1b, L.D1P / 4 + 1, reinterpret_cast<const float4*>(lay + f.o_w1),
                              L.WP, b.red, L.red_floats);
        tile_gemm_prefetch<T>(L.WP / 4 + 1, reinterpret_cast<const float4*>(lay + f.o_w2), L.WP,
                              L.red_floats);
        __syncthreads();
        hidden_epilogue<T, true, false>(L, b.red, KS, L.WP, 0, b.h1, nullptr);
        __syncthreads();
        KS = mlp_tail<T, false>(L, b, lay, f, k, lay + f.o_mix_inv, L.DP / 4, L.DP);
        for (int i = threadIdx.x; i < T * L.d2; i += FAB_NT) {
            const int j = i / T, p = i - j * T;
            const float shift = red_sum<T>(b.red, KS, L.P2, p, j);
            const float scale = red_sum<T>(b.red, KS, L.P2, p, L.d2 + j);
            const int e = kidx<T>(p, L.d1 + j);
            b.zs[e] = b.zs[e] * expf(scale) + shift;
            b.scl[p * L.d2 + j] = scale;
        }
        __syncthreads();
        // log q -= sum(scale);  log q -= (-sum log_S)
        logdet_accumulate<T>(L, b, b.logs[k]);
        // u' = [v1,y2] @ Wmix^-1
        KS = tile_gemm<T>(b.zs, L.DP / 4, reinterpret_cast<const float4*>(lay + f.o_mix_inv), L.DP,
                          b.red, L.red_floats);
        if (k + 1 < L.K)
            tile_gemm_prefetch<T>(L.D1P / 4 + 1,
                                  reinterpret_cast<const float4*>(lay + f.layer_stride + f.o_w1),
                                  L.WP, L.red_floats);
        __syncthreads();
        for (int e = threadIdx.x; e < TP * L.d; e += FAB_NT) {
            int p, n;
            kdecode<T>(e, p, n);
            if (p < T) b.zs[e] = red_sum<T>(b.red, KS, L.DP, p, n);
        }
        __syncthreads();
    }
    for (int p = threadIdx.x; p < T; p += FAB_NT) lq_out[p] = b.ld[p];
    __syncthreads();
}
