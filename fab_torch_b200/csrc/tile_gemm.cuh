// Tile GEMM: out[p][n] = sum_k act[p][k] * M[k][n] for the T particles of one CTA.
//
//   act : shared memory operand, k-major with particles fastest:  act[k][TP]  (TP = T rounded up
//         to 4), element (p,k) at float index kidx<T>(p,k) = k*TP + p.  One LDS.128 therefore
//         yields two *particle pairs* of one k, ready as the packed operands of fma.rn.f32x2.
//         Biases are part of the operand: after the K real rows comes a row of ones (then three
//         zero rows) and the matching weight rows hold (bias,0,0,0) -- see include/fab_b200.h.
//   Wp  : global (L2-resident) packed operand, float4 [K4][NP]
//   red : shared reduction buffer [KS][T][NP+4] (row padded by 4 floats: conflict-free for the
//         fused epilogues)
//
// Work decomposition.  A *unit* is (k-split ks, column group ng); it owns the FAB_TN output
// columns n_t = ng + t*NP/FAB_TN for ALL T particles over the k-range of its split, i.e.
// FAB_TN*T fp32 accumulators held as FAB_TN*T/2 packed pairs.  Each packed weight word is loaded
// from L2 exactly once per CTA, by exactly one thread, straight into registers (coalesced,
// non-allocating), and activations are read with warp-broadcast LDS.128.  The inner product runs
// on Blackwell's packed FP32 pipe: per 4-wide k step a unit issues FAB_TN LDG.128 + ~TP LDS.128 +
// 2*FAB_TN*T FFMA2 (fma.rn.f32x2, two particles per instruction), which halves the issue slots
// per FLOP -- measured on B200 (profiles/microbench_ffma.cu) this loop sustains 80-84 % of the
// 74.4 TFLOP/s FP32 peak versus 60 % for scalar FFMA.  Weight loads run FAB_PF steps ahead in a
// register ring; the steady-state loop is branch-free.
// tile_gemm_prefetch() pulls the first weight tiles of the NEXT GEMM into L1 while the current
// epilogue runs, hiding the L2 latency that would otherwise be exposed after every barrier.
//
// Roofline: FP32 FMA pipe (tensor cores cannot hold the 1e-5-relative fp32 parity bar in one
// pass, and at <= 2048 particles per GPU a 128-row UMMA tile would leave 132 of 148 SMs idle;
// see DESIGN.md §4).
#pragma once
#include "common.cuh"

#ifndef FAB_PF
#define FAB_PF 2   // weight prefetch distance in k4 steps
#endif

typedef unsigned long long fab_u64;
__device__ __forceinline__ fab_u64 pack2(float a, float b) {
    fab_u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(fab_u64 v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ fab_u64 ffma2(fab_u64 a, fab_u64 b, fab_u64 c) {
    fab_u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

template <int T> struct TileDims { static constexpr int TP = (T + 3) & ~3; };

template <int T>
__device__ __forceinline__ int kidx(int p, int n) { return n * TileDims<T>::TP + p; }
// inverse of kidx for a linear element index e of an operand buffer (p may be a pad slot >= T)
template <int T>
__device__ __forceinline__ void kdecode(int e, int& p, int& n) {
    n = e / TileDims<T>::TP;
    p = e - n * TileDims<T>::TP;
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

// one 4-wide k step: acc[q][t] (+)= (act[2q][k], act[2q+1][k]) * (w_t[k], w_t[k]) for 4 k's
template <int T>
__device__ __forceinline__ void gemm_step(fab_u64 (&acc)[T / 2][FAB_TN], const float* a,
                                          const float4 (&w)[FAB_TN]) {
    constexpr int TP = TileDims<T>::TP;
    constexpr int NQ = T / 2;                       // particle pairs
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        fab_u64 w2[FAB_TN];
#pragma unroll
        for (int t = 0; t < FAB_TN; ++t) {
            const float wv = c == 0 ? w[t].x : c == 1 ? w[t].y : c == 2 ? w[t].z : w[t].w;
            w2[t] = pack2(wv, wv);
        }
        const fab_u64* arow = reinterpret_cast<const fab_u64*>(a + c * TP);
#pragma unroll
        for (int q = 0; q < NQ; q += 2) {
            fab_u64 p0, p1 = 0;
            if (q + 1 < NQ) {                       // LDS.128: two pairs
                const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(arow + q);
                p0 = v.x; p1 = v.y;
            } else {                                // trailing pair: LDS.64
                p0 = arow[q];
            }
#pragma unroll
            for (int t = 0; t < FAB_TN; ++t) {
                acc[q][t] = ffma2(p0, w2[t], acc[q][t]);
                if (q + 1 < NQ) acc[q + 1][t] = ffma2(p1, w2[t], acc[q + 1][t]);
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
    static_assert(T % 2 == 0, "particles are processed in pairs");
    constexpr int TP = TileDims<T>::TP;
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
        fab_u64 acc[T / 2][FAB_TN];
#pragma unroll
        for (int q = 0; q < T / 2; ++q)
#pragma unroll
            for (int t = 0; t < FAB_TN; ++t) acc[q][t] = 0ull;
        if (nsteps > 0) {
            const float4* wp = Wp + (size_t)k4b * NP + ng;          // column t: + t*NG
            const float* a = act + (size_t)k4b * 4 * TP;
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
                    gemm_step<T>(acc, a + (size_t)(s + i) * 4 * TP, w);
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
                        gemm_step<T>(acc, a + (size_t)(s + i) * 4 * TP, w);
                    }
                }
            }
        }
        float* r = red + (size_t)ks * T * NPs + ng;
#pragma unroll
        for (int q = 0; q < T / 2; ++q)
#pragma unroll
            for (int t = 0; t < FAB_TN; ++t) {
                float lo, hi;
                unpack2(acc[q][t], lo, hi);
                r[(2 * q) * NPs + t * NG] = lo;
                r[(2 * q + 1) * NPs + t * NG] = hi;
            }
    }
    return g.KS;
}

// sum of the k-split partials of one output element
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
