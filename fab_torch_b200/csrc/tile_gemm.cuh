// Tile GEMM: out[p][n] = sum_k act[p][k] * M[k][n] for the T particles of one CTA.
//
//   act : shared memory, "k4-major" operand layout  float4 act4[K4][T]  (act4[k4][p] holds
//         act[p][4*k4 .. 4*k4+3]); element (p,k) lives at float index kidx<T>(p,k).  With T a
//         compile-time constant every LDS.128 of the inner loop has an immediate offset.
//         Biases are part of the operand: the last k4 block of `act` is the constant (1,0,0,0)
//         and the matching weight rows hold (bias,0,0,0) -- see include/fab_b200.h.
//   Wp  : global (L2-resident) packed operand, float4 [K4][NP]
//   red : shared reduction buffer [KS][T][NP+4] (row padded by 4 floats so that the fused
//         epilogues, which walk particles fastest, read it without bank conflicts)
//
// Work decomposition.  A *unit* is (k-split ks, column group ng); it owns the FAB_TN output
// columns n_t = ng + t*NP/FAB_TN for ALL T particles over the k-range of its split, i.e.
// FAB_TN*T fp32 accumulators in registers.  Each packed weight word is therefore loaded from L2
// exactly once per CTA, by exactly one thread, straight into registers (coalesced: consecutive ng
// read consecutive float4; non-allocating so the stream does not evict the small L1-resident
// tables), and activations are read with warp-broadcast LDS.128.  Per 4-wide k step a unit issues
// FAB_TN LDG.128 + T LDS.128 + 4*FAB_TN*T FFMA.  Weight loads run FAB_PF steps ahead in a register
// ring; the steady-state loop is branch-free.
// tile_gemm_prefetch() pulls the first weight tiles of the NEXT GEMM into L1 while the current
// epilogue runs, hiding the L2 latency that would otherwise be exposed after every barrier.
//
// Roofline: FP32 FFMA pipe (tensor cores cannot hold the 1e-5-relative fp32 parity bar in one
// pass, and at <= 2048 particles per GPU a 128-row UMMA tile would leave 132 of 148 SMs idle;
// see DESIGN.md §4).
#pragma once
#include "common.cuh"

#ifndef FAB_TN
#define FAB_TN 4   // output columns per unit
#endif
#ifndef FAB_PF
#define FAB_PF 2   // weight prefetch distance in k4 steps
#endif

template <int T>
__device__ __forceinline__ int kidx(int p, int n) { return (((n >> 2) * T + p) << 2) | (n & 3); }
// inverse of kidx for a linear element index e of a k4-major buffer
template <int T>
__device__ __forceinline__ void kdecode(int e, int& p, int& n) {
    const int q = e >> 2;              // float4 index = k4*T + p
    const int k4 = q / T;
    p = q - k4 * T;
    n = (k4 << 2) | (e & 3);
}

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
        : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

struct GemmPlan { int NG, KS, ksteps; };

template <int T>
__device__ __forceinline__ GemmPlan gemm_plan(int K4, int NP, int red_floats) {
    GemmPlan g;
    g.NG = NP / FAB_TN;
    int KS = FAB_NT / g.NG;
    if (KS < 1) KS = 1;
    if (KS > K4) KS = K4;
    const int cap = red_floats / (T * (NP + 4));
    if (KS > cap) KS = cap;
    g.KS = KS;
    g.ksteps = (K4 + KS - 1) / KS;
    return g;
}

template <int T>
__device__ __forceinline__ void gemm_step(float (&acc)[T][FAB_TN], const float4* a4,
                                          const float4 (&w)[FAB_TN]) {
#pragma unroll
    for (int p = 0; p + 1 < T; p += 2) {
        const float4 a0 = a4[p];
        const float4 a1 = a4[p + 1];
        // component-major over the pair: 2*FAB_TN independent FFMAs between dependent ones
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float x0 = c == 0 ? a0.x : c == 1 ? a0.y : c == 2 ? a0.z : a0.w;
            const float x1 = c == 0 ? a1.x : c == 1 ? a1.y : c == 2 ? a1.z : a1.w;
#pragma unroll
            for (int t = 0; t < FAB_TN; ++t) {
                const float wv = c == 0 ? w[t].x : c == 1 ? w[t].y : c == 2 ? w[t].z : w[t].w;
                acc[p][t] = fmaf(x0, wv, acc[p][t]);
                acc[p + 1][t] = fmaf(x1, wv, acc[p + 1][t]);
            }
        }
    }
    if (T & 1) {
        const float4 a0 = a4[T - 1];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float x0 = c == 0 ? a0.x : c == 1 ? a0.y : c == 2 ? a0.z : a0.w;
#pragma unroll
            for (int t = 0; t < FAB_TN; ++t) {
                const float wv = c == 0 ? w[t].x : c == 1 ? w[t].y : c == 2 ? w[t].z : w[t].w;
                acc[T - 1][t] = fmaf(x0, wv, acc[T - 1][t]);
            }
        }
    }
}

// Pull the weight words the first FAB_PF steps of tile_gemm(K4, Wp, NP) will read into L1.
template <int T>
__device__ __forceinline__ void tile_gemm_prefetch(int K4, const float4* __restrict__ Wp, int NP,
                                                   int red_floats) {
    const GemmPlan g = gemm_plan<T>(K4, NP, red_floats);
    const int u = threadIdx.x;
    if (u >= g.NG * g.KS) return;
    const int ks = u / g.NG;
    const int ng = u - ks * g.NG;
    const int k4b = ks * g.ksteps;
#pragma unroll
    for (int i = 0; i < FAB_PF; ++i) {
        if (k4b + i < K4) {
            const float4* wp = Wp + (size_t)(k4b + i) * NP + ng;
#pragma unroll
            for (int t = 0; t < FAB_TN; ++t)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(wp + (size_t)t * g.NG));
        }
    }
}

// Computes all partial sums into red[ks][p][n] (row stride NP+4).  Returns the number of k-splits
// KS.  Caller must __syncthreads() before reading `red` and again before the next tile_gemm.
template <int T>
__device__ __noinline__ int tile_gemm(const float* act, int K4, const float4* __restrict__ Wp,
                                      int NP, float* red, int red_floats) {
    const GemmPlan g = gemm_plan<T>(K4, NP, red_floats);
    const int NG = g.NG;
    const int NPs = NP + 4;
    const int units = NG * g.KS;
    for (int u = threadIdx.x; u < units; u += FAB_NT) {
        const int ks = u / NG;
        const int ng = u - ks * NG;
        const int k4b = ks * g.ksteps;
        int nsteps = K4 - k4b;
        if (nsteps > g.ksteps) nsteps = g.ksteps;
        float acc[T][FAB_TN];
#pragma unroll
        for (int p = 0; p < T; ++p)
#pragma unroll
            for (int t = 0; t < FAB_TN; ++t) acc[p][t] = 0.f;
        if (nsteps > 0) {
            const float4* wp = Wp + (size_t)k4b * NP + ng;          // column t: + t*NG
            const float4* a4 = reinterpret_cast<const float4*>(act) + (size_t)k4b * T;
            float4 ring[FAB_PF][FAB_TN];
#pragma unroll
            for (int i = 0; i < FAB_PF; ++i) {
                if (i < nsteps) {
#pragma unroll
                    for (int t = 0; t < FAB_TN; ++t)
                        ring[i][t] = ldg_stream(wp + (size_t)i * NP + t * NG);
                }
            }
            int s = 0;
            // steady state: every step of the group and its prefetch target are in range
            for (; s + 2 * FAB_PF <= nsteps; s += FAB_PF) {
#pragma unroll
                for (int i = 0; i < FAB_PF; ++i) {
                    float4 w[FAB_TN];
#pragma unroll
                    for (int t = 0; t < FAB_TN; ++t) {
                        w[t] = ring[i][t];
                        ring[i][t] = ldg_stream(wp + (size_t)(s + i + FAB_PF) * NP + t * NG);
                    }
                    gemm_step<T>(acc, a4 + (size_t)(s + i) * T, w);
                }
            }
            // tail: fewer than 2*FAB_PF steps left
            for (; s < nsteps; s += FAB_PF) {
#pragma unroll
                for (int i = 0; i < FAB_PF; ++i) {
                    if (s + i < nsteps) {
                        float4 w[FAB_TN];
#pragma unroll
                        for (int t = 0; t < FAB_TN; ++t) w[t] = ring[i][t];
                        if (s + i + FAB_PF < nsteps) {
#pragma unroll
                            for (int t = 0; t < FAB_TN; ++t)
                                ring[i][t] = ldg_stream(wp + (size_t)(s + i + FAB_PF) * NP + t * NG);
                        }
                        gemm_step<T>(acc, a4 + (size_t)(s + i) * T, w);
                    }
                }
            }
        }
        float* r = red + (size_t)ks * T * NPs + ng;
#pragma unroll
        for (int p = 0; p < T; ++p)
#pragma unroll
            for (int t = 0; t < FAB_TN; ++t) r[p * NPs + t * NG] = acc[p][t];
    }
    return g.KS;
}

// sum of the k-split partials of one output element / of four consecutive columns
template <int T>
__device__ __forceinline__ float red_sum(const float* red, int KS, int NP, int p, int n) {
    const int NPs = NP + 4;
    const float* r = red + p * NPs + n;
    float s = r[0];
    const int stride = T * NPs;
#pragma unroll 4
    for (int ks = 1; ks < KS; ++ks) { r += stride; s += r[0]; }
    return s;
}
template <int T>
__device__ __forceinline__ float4 red_sum4(const float* red, int KS, int NP, int p, int n0) {
    const int NPs = NP + 4;
    const float* r = red + p * NPs + n0;
    float4 s = *reinterpret_cast<const float4*>(r);
    const int stride = T * NPs;
#pragma unroll 4
    for (int ks = 1; ks < KS; ++ks) {
        r += stride;
        const float4 v = *reinterpret_cast<const float4*>(r);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    return s;
}
