// Shared device helpers for the tile kernels (sm_100a).
//
// Execution model: one thread block ("tile CTA", FAB_NT threads) carries T particles through a
// whole flow evaluation / HMC outer step with every intermediate in shared memory; only the
// Point rows and noise touch HBM.  Weights stream from L2 straight into registers in MMA
// fragment order (each weight word is consumed by exactly one warp of the CTA), see mma_gemm.cuh.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include "fab_b200.h"

#define FAB_NT 256              // threads per tile CTA: 8 warps, fixed by the warp -> n-tile maps of
                                // mma_gemm.cuh (16 warps were measured slower: the operand splits of
                                // the shared A fragments are repeated per warp)
#ifndef FAB_MIN_CTAS
#define FAB_MIN_CTAS 1          // co-resident tile CTAs per SM the kernels are compiled for
#endif
#define FAB_NWARPS (FAB_NT / 32)
#define FAB_FULL 0xffffffffu

__host__ __device__ __forceinline__ int fab_round4(int v) { return (v + 3) & ~3; }

// Optional per-phase cycle profile of CTA 0 (-DFAB_PROF; experiment builds only, see
// profiles/phase_profile.py).  prof_mark(id) charges the cycles since the previous mark to `id`.
#ifdef FAB_PROF
__device__ unsigned long long g_fab_prof[32];
__device__ unsigned long long g_fab_cta_cycles[1024];      // per-CTA cycles of the last k_hmc_step
__device__ unsigned int g_fab_cta_smid[1024];
__device__ __forceinline__ void prof_mark(int id) {
    __shared__ long long s_prof_last;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const long long t = clock64();
        if (id >= 0) atomicAdd(&g_fab_prof[id], (unsigned long long)(t - s_prof_last));
        s_prof_last = t;
    }
}
#else
__device__ __forceinline__ void prof_mark(int) {}
#endif

// Shared-memory carve-up of one tile CTA; all offsets in floats.  Filled on the host, passed by
// value to the kernels (so host and device agree by construction).
//
// A CTA carries T <= TP particles in TP "slots" (TP = 16, or 8 for small batches / very wide
// flows); slot = row of the warp-level MMA tile (mma_gemm.cuh).  zs0/zs1, gs, par, h1, h2 are MMA
// operands in k-major layout act[k][S] (S = slot stride, 24 floats for TP=16 so that the
// fragment loads are bank-conflict free, 8 for TP=8); rows are padded to a multiple of 16 (one
// k-tile pair) and pad rows / pad slots always hold finite values (0 unless an epilogue wrote
// bias-only garbage into a pad slot).  h1 carries D16 extra rows behind the hidden rows: the
// [gv] part of the merged backward GEMM.  Everything else is row-major per particle.
struct TileLayout {
    int T, TP, S;       // particles per CTA, slots (8|16), slot stride of the operand buffers
    int d, DP;          // dim and round_up(dim,4) (row-major state tiles)
    int d1, d2;         // conditioner width, transformed width
    int D8, D16, W8, W16, P8, P16, D1K;   // n-pads (8) / k-pads (16) of d, W, 2*d2; D1K = k-pad of d1
    int NTH;            // hidden n-tiles = W8/8
    int MW;             // ReLU-mask words per layer and MLP stage
    int K;              // coupling layers
    int RS;             // slot stride of the split-K partial buffer
    int red_floats;     // capacity of the split-K partial buffer
    // offsets
    int o_zs0, o_zs1, o_gs, o_par, o_h1, o_h2, o_red;
    int o_sy2, o_ses, o_m1, o_m2;       // saved-for-backward, [K][...]
    int o_ld;                           // [TP] running log-det
    int o_scl;                          // [DP + d2][TP] per-(row, slot) terms of the per-particle sums
    int o_const;                        // loc[DP], log_scale[DP], 1/scale[DP], logs[K], sum(logs)
    int o_state;                        // kernel-specific state area
    int total_floats;
};

__host__ __device__ __forceinline__ int fab_round8(int v) { return (v + 7) & ~7; }
__host__ __device__ __forceinline__ int fab_round16(int v) { return (v + 15) & ~15; }

// k-split plan of the narrow GEMMs (mma_gemm.cuh: mma_gemm_ksplit): NT n-tiles are spread over
// NGR warp groups (<= 4 tiles per warp and pass), the remaining factor of the warps splits K.
__host__ __device__ inline void fab_ksplit_plan(int NT, int& NGR, int& KS) {
    const int need = (NT + 3) / 4;
    NGR = 1;
    while (NGR < need && NGR < FAB_NWARPS) NGR <<= 1;
    KS = FAB_NWARPS / NGR;
}

// state_floats: extra per-CTA floats the calling kernel wants after the evaluation buffers.
__host__ inline TileLayout make_tile_layout(const fab_flow_desc& f, int T, int TP, bool with_grad,
                                            int state_floats) {
    TileLayout L{};
    L.T = T; L.TP = TP; L.S = TP == 16 ? 24 : 8; L.RS = TP == 16 ? 20 : 12;
    L.d = f.dim; L.DP = fab_round4(f.dim);
    L.d1 = f.d1; L.d2 = f.d2;
    L.D8 = fab_round8(f.dim); L.D16 = fab_round16(f.dim);
    L.W8 = f.width_pad; L.W16 = f.width_kpad;
    L.P8 = fab_round8(2 * f.d2); L.P16 = fab_round16(2 * f.d2);
    L.D1K = fab_round16(f.d1);
    L.NTH = L.W8 / 8;
    L.MW = L.NTH * (TP == 16 ? 4 : 2);
    L.K = f.n_layers;
    auto need = [&](int N8) {
        int NGR, KS; fab_ksplit_plan(N8 / 8, NGR, KS);
        return KS * N8 * L.RS;
    };
    L.red_floats = need(L.P8) > need(L.D8) ? need(L.P8) : need(L.D8);
    const int S = L.S;
    int o = 0;
    auto take = [&](int n) { int r = o; o += fab_round4(n); return r; };
    L.o_zs0 = take(L.D16 * S);
    L.o_zs1 = take(L.D16 * S);
    L.o_gs = take(L.D16 * S);
    L.o_par = take(L.P16 * S);
    L.o_h1 = take((L.W16 + L.D16) * S);
    L.o_h2 = take((L.W16 > 0 ? L.W16 : 16) * S);
    L.o_red = take(L.red_floats);
    const int KSV = with_grad ? L.K : 0;
    L.o_sy2 = take(KSV * L.d2 * TP);
    L.o_ses = take(KSV * L.d2 * TP);
    L.o_m1 = take(KSV * L.MW);
    L.o_m2 = take(KSV * L.MW);
    L.o_ld = take(TP);
    L.o_scl = take((L.DP + L.d2) * TP);
    L.o_const = take(3 * L.DP + L.K + 1);
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
