// Shared device helpers for the tile kernels (sm_100a).
//
// Execution model: one thread block ("tile CTA", FAB_NT threads) carries T particles through a
// whole flow evaluation / HMC outer step with every intermediate in shared memory; only the
// Point rows and noise touch HBM.  Weights stream from L2 straight into registers (each packed
// weight word is consumed by exactly one thread of the CTA), see tile_gemm.cuh.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include "fab_b200.h"

#ifndef FAB_NT
#define FAB_NT 256              // threads per tile CTA (8 warps = 2 per scheduler; best of the sweep in profiles/)
#endif
#ifndef FAB_MIN_CTAS
#define FAB_MIN_CTAS 1          // co-resident tile CTAs per SM the kernels are compiled for
#endif
#define FAB_NWARPS (FAB_NT / 32)
#ifndef FAB_TN
#define FAB_TN 4                // output columns per GEMM unit (tile_gemm.cuh)
#endif
#define FAB_FULL 0xffffffffu

__host__ __device__ __forceinline__ int fab_round4(int v) { return (v + 3) & ~3; }

// Shared-memory carve-up of one tile CTA; all offsets in floats.  Filled on the host, passed by
// value to the kernels (so host and device agree by construction).  zs, vs, z1b, par, h1, h2 are
// GEMM operands in the k-major / particle-fastest layout of tile_gemm.cuh (act[k][TP], TP = T
// rounded up to 4); zs, z1b, h1, h2 carry 4 extra rows at the end (the constant-one bias row; for
// h1 also the [gv] rows of the merged backward GEMM).  Everything else is row-major per particle.
struct TileLayout {
    int T, TP;          // particles per CTA, rounded up to 4
    int d, DP;          // dim and round_up(dim,4)
    int d1, d2, D1P, P2;// conditioner width, transformed width, pads (P2 = round_up(2*d2,4))
    int WP;             // padded hidden width
    int MW;             // ReLU-mask words per layer and MLP stage
    int K;              // coupling layers
    int red_floats;     // capacity of the split-K reduction buffer
    // offsets
    int o_zs, o_vs, o_z1b, o_par, o_h1, o_h2, o_red;
    int o_sy2, o_ses, o_m1, o_m2;       // saved-for-backward, [K][...]
    int o_ld;                           // [T] running log-det
    int o_scl;                          // [T][d2] coupling scales of the current layer
    int o_const;                        // loc[DP], log_scale[DP], 1/scale[DP], logs[K]
    int o_state;                        // kernel-specific state area
    int total_floats;
};

// state_floats: extra per-CTA floats the calling kernel wants after the evaluation buffers.
__host__ inline TileLayout make_tile_layout(const fab_flow_desc& f, int T, bool with_grad,
                                            int state_floats) {
    TileLayout L{};
    L.T = T; L.TP = fab_round4(T); L.d = f.dim; L.DP = fab_round4(f.dim);
    const int TP = L.TP;
    L.d1 = f.d1; L.d2 = f.d2; L.D1P = fab_round4(f.d1 > 0 ? f.d1 : 1);
    L.P2 = fab_round4(2 * f.d2 > 0 ? 2 * f.d2 : 1);
    L.WP = f.width_pad > 0 ? f.width_pad : 4;
    L.MW = 4 * ((TP * L.WP / 4 + 31) / 32);
    L.K = f.n_layers;
    // reduction buffer: large enough for the k-splits that let each wide GEMM (N >= WP) occupy all
    // FAB_NT threads (same arithmetic as gemm_plan in tile_gemm.cuh); rows are padded by 4 floats.
    auto need = [&](int NP, int K4) {
        int ks = FAB_NT / (NP / FAB_TN > 0 ? NP / FAB_TN : 1);
        if (ks < 1) ks = 1;
        if (ks > K4) ks = K4;
        if (ks > 8) ks = 8;
        return ks * T * (NP + 4);
    };
    L.red_floats = need(L.DP + L.WP, L.DP / 4 + 1);
    if (need(L.WP, L.WP / 4 + 1) > L.red_floats) L.red_floats = need(L.WP, L.WP / 4 + 1);
    if (need(L.WP, L.P2 / 4) > L.red_floats) L.red_floats = need(L.WP, L.P2 / 4);
    if (need(L.DP, (L.WP + L.DP) / 4) > L.red_floats) L.red_floats = need(L.DP, (L.WP + L.DP) / 4);
    int o = 0;
    auto take = [&](int n) { int r = o; o += fab_round4(n); return r; };
    L.o_zs = take(TP * (L.DP + 4));
    L.o_vs = take(TP * L.DP);
    L.o_z1b = take(TP * (L.D1P + 4));
    L.o_par = take(TP * L.P2);
    L.o_h1 = take(TP * (L.WP + (L.DP > 4 ? L.DP : 4)));
    L.o_h2 = take(TP * (L.WP + 4));
    L.o_red = take(L.red_floats);
    int KS = with_grad ? L.K : 0;
    L.o_sy2 = take(KS * T * L.d2);
    L.o_ses = take(KS * T * L.d2);
    L.o_m1 = take(KS * L.MW);
    L.o_m2 = take(KS * L.MW);
    L.o_ld = take(T);
    L.o_scl = take(T * L.d2);
    L.o_const = take(3 * L.DP + L.K);
    L.o_state = take(state_floats);
    L.total_floats = o;
    return L;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FAB_FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FAB_FULL, v, o));
    return v;
}

__device__ __forceinline__ bool fab_isfinite(float v) { return fabsf(v) <= 3.402823466e38f; }

// gamma(x) = cq*log_q + cp*log_p with every operation rounded separately (no FMA contraction),
// i.e. the op-by-op fp32 arithmetic torch performs for base.py:94/97.
__device__ __forceinline__ float gamma_of(const fab_gamma& g, float lq, float lp) {
    return __fadd_rn(__fmul_rn(g.cq, lq), __fmul_rn(g.cp, lp));
}
// grad U = nan_to_num(clamp(-(gq*grad_q + gp*grad_p), +-max_grad), nan=0)   (hmc.py:194-199)
__device__ __forceinline__ float grad_u_of(const fab_gamma& g, float dq, float dp, float max_grad) {
    float v = -__fadd_rn(__fmul_rn(g.gq, dq), __fmul_rn(g.gp, dp));
    if (v != v) return 0.0f;                       // NaN survives clamp, then -> 0
    return fminf(fmaxf(v, -max_grad), max_grad);   // +-inf clamp to +-max_grad first
}
