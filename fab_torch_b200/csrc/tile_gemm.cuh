// Tile GEMM: out[p][n] = sum_k act[p][k] * M[k][n] for the T particles of one CTA.
//
//   act : shared memory, [T][lda] fp32, zero-padded to a multiple of 4 columns
//   Wp  : global (L2-resident) packed operand, float4 [K4][NP] (see include/fab_b200.h)
//
// Work decomposition.  A *unit* is (k-split ks, column group ng); it owns the two output columns
// n0 = ng and n1 = ng + NP/2 for ALL T particles over the k-range of its split, i.e. 2*T fp32
// accumulators in registers.  Each packed weight word is therefore loaded from L2 exactly once
// per CTA, by exactly one thread, directly into registers (coalesced: consecutive ng read
// consecutive float4), and activations are read with warp-broadcast LDS.128.  Per 4-wide k step
// a unit issues 2 LDG.128 + T LDS.128 + 8*T FFMA.  Weight loads run PF steps ahead in a register
// ring so L2 latency is covered by the FFMA stream.  Partial sums of the k-splits go to a
// shared reduction buffer red[ks][p][n]; the caller sums them in its fused epilogue.
//
// Roofline: FP32 FFMA pipe (tensor cores cannot hold the 1e-5-relative fp32 parity bar at one
// pass, and at <= 2048 particles per GPU a 128-row UMMA tile would leave 132 of 148 SMs idle;
// see DESIGN.md §4).
#pragma once
#include "common.cuh"

#define FAB_PF 4   // weight prefetch distance in k4 steps

struct GemmSplit {
    int KS;       // number of k-splits actually used
};

template <int T>
__device__ __forceinline__ void tile_gemm_accumulate(float (&acc)[T][2], const float* act, int lda,
                                                     const float4* __restrict__ w0p, int NG, int NP,
                                                     int k4b, int nsteps) {
    // w0p points at Wp[k4b][ng]; the second column lives NG float4 further.
    float4 ring[FAB_PF][2];
#pragma unroll
    for (int i = 0; i < FAB_PF; ++i) {
        if (i < nsteps) {
            ring[i][0] = __ldg(w0p + (size_t)i * NP);
            ring[i][1] = __ldg(w0p + (size_t)i * NP + NG);
        }
    }
    const float* a_base = act + k4b * 4;
    for (int s = 0; s < nsteps; s += FAB_PF) {
#pragma unroll
        for (int i = 0; i < FAB_PF; ++i) {
            if (s + i < nsteps) {
                const float4 w0 = ring[i][0];
                const float4 w1 = ring[i][1];
                if (s + i + FAB_PF < nsteps) {
                    ring[i][0] = __ldg(w0p + (size_t)(s + i + FAB_PF) * NP);
                    ring[i][1] = __ldg(w0p + (size_t)(s + i + FAB_PF) * NP + NG);
                }
                const float* a = a_base + (s + i) * 4;
#pragma unroll
                for (int p = 0; p < T; ++p) {
                    const float4 av = *reinterpret_cast<const float4*>(a + p * lda);
                    acc[p][0] = fmaf(av.x, w0.x, acc[p][0]);
                    acc[p][1] = fmaf(av.x, w1.x, acc[p][1]);
                    acc[p][0] = fmaf(av.y, w0.y, acc[p][0]);
                    acc[p][1] = fmaf(av.y, w1.y, acc[p][1]);
                    acc[p][0] = fmaf(av.z, w0.z, acc[p][0]);
                    acc[p][1] = fmaf(av.z, w1.z, acc[p][1]);
                    acc[p][0] = fmaf(av.w, w0.w, acc[p][0]);
                    acc[p][1] = fmaf(av.w, w1.w, acc[p][1]);
                }
            }
        }
    }
}

// Computes all partial sums into red[ks][p][n] (n < NP).  Caller must __syncthreads() before
// reading `red` and again before the next tile_gemm overwrites it.
template <int T>
__device__ __forceinline__ GemmSplit tile_gemm(const float* act, int lda, int K4,
                                               const float4* __restrict__ Wp, int NP, float* red,
                                               int red_floats) {
    const int NG = NP >> 1;
    int KS = FAB_NT / NG;
    if (KS < 1) KS = 1;
    if (KS > K4) KS = K4;
    const int cap = red_floats / (T * NP);
    if (KS > cap) KS = cap;
    const int ksteps = (K4 + KS - 1) / KS;
    const int units = NG * KS;
    for (int u = threadIdx.x; u < units; u += FAB_NT) {
        const int ks = u / NG;
        const int ng = u - ks * NG;
        const int k4b = ks * ksteps;
        int nsteps = K4 - k4b;
        if (nsteps > ksteps) nsteps = ksteps;
        float acc[T][2];
#pragma unroll
        for (int p = 0; p < T; ++p) { acc[p][0] = 0.f; acc[p][1] = 0.f; }
        if (nsteps > 0)
            tile_gemm_accumulate<T>(acc, act, lda, Wp + (size_t)k4b * NP + ng, NG, NP, k4b, nsteps);
        float* r = red + (size_t)ks * T * NP + ng;
#pragma unroll
        for (int p = 0; p < T; ++p) {
            r[p * NP] = acc[p][0];
            r[p * NP + NG] = acc[p][1];
        }
    }
    GemmSplit g; g.KS = KS;
    return g;
}

template <int T>
__device__ __forceinline__ float red_sum(const float* red, int KS, int NP, int p, int n) {
    float s = red[p * NP + n];
    for (int ks = 1; ks < KS; ++ks) s += red[(size_t)(ks * T + p) * NP + n];
    return s;
}
