// Microbenchmark (B200): which structural choice of the 3xTF32 tile-GEMM loop costs what.
// Same job as microbench_mma_gemm.cu ([16x320]x[320x320] + ReLU/mask epilogue + barrier, weights
// from L2 in fragment order), but with a deliberately simple loop whose features are switched by
// macros:  -DV_FRESH (fresh accumulator per k-tile + FADD), -DV_RN (+0x1000 on the A lo part),
// -DV_NOVOL (non-volatile asm), -DV_INTERLEAVE (tile-interleaved MMA order), -DV_RING3.
#include <cstdio>
#include <vector>
#include "flow_tile.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

#ifdef V_NOVOL
#define ASMV asm
#else
#define ASMV asm volatile
#endif
__device__ __forceinline__ void mma_v(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    ASMV("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
         : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_a(float v, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(v) & 0xffffe000u;
#ifdef V_RN
    lo = __float_as_uint(v - __uint_as_float(hi)) + 0x1000u;
#else
    lo = __float_as_uint(v - __uint_as_float(hi));
#endif
}

__device__ __forceinline__ void add2(float& a0, float& a1, float b0, float b1) {
#ifdef V_FADD2
    unsigned long long a, b;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(a));
#else
    a0 += b0; a1 += b1;
#endif
}
// lo parts of two weights at once
__device__ __forceinline__ void split_w2(float v0, float v1, uint32_t& h0, uint32_t& h1, uint32_t& l0, uint32_t& l1) {
#ifdef V_SPLIT2
    h0 = __float_as_uint(v0) & 0xffffe000u; h1 = __float_as_uint(v1) & 0xffffe000u;
    unsigned long long v, h, l;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(v0), "f"(v1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(h) : "r"(h0 ^ 0x80000000u), "r"(h1 ^ 0x80000000u));   // -hi
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(l) : "l"(v), "l"(h));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(l0), "=r"(l1) : "l"(l));
#else
    split_w(v0, h0, l0); split_w(v1, h1, l1);
#endif
}

template <class Epi>
__device__ __forceinline__ void gemm_simple(const float* act, int KT2, const float4* __restrict__ Wf, int NT,
                                            const float* __restrict__ bias, Epi epi) {
    constexpr int S = 24, NTW = 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    float c[NTW][4];
#pragma unroll
    for (int i = 0; i < NTW; ++i) {
        const float2 bv = *reinterpret_cast<const float2*>(bias + (warp + 8 * i) * 8 + 2 * t);
        c[i][0] = bv.x; c[i][1] = bv.y; c[i][2] = bv.x; c[i][3] = bv.y;
    }
    const float4* W = Wf + (size_t)warp * 32 + lane;
    auto do_pair = [&](const float4 (&w)[NTW], int kp) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int kb = kp * 16 + h * 8;
            uint32_t ah[4], al[4];
            const float* ap = act + (size_t)(kb + t) * S + g;
            split_a(ap[0], ah[0], al[0]); split_a(ap[8], ah[1], al[1]);
            split_a(ap[4 * S], ah[2], al[2]); split_a(ap[4 * S + 8], ah[3], al[3]);
            uint32_t bh[NTW][2], bl[NTW][2];
#pragma unroll
            for (int i = 0; i < NTW; ++i)
                split_w2(h == 0 ? w[i].x : w[i].z, h == 0 ? w[i].y : w[i].w, bh[i][0], bh[i][1], bl[i][0], bl[i][1]);
#ifdef V_FRESH
            float cp[NTW][4];
#pragma unroll
            for (int i = 0; i < NTW; ++i) cp[i][0] = cp[i][1] = cp[i][2] = cp[i][3] = 0.f;
#define ACC cp
#else
#define ACC c
#endif
#ifdef V_INTERLEAVE
#pragma unroll
            for (int i = 0; i < NTW; ++i) mma_v(ACC[i], al, bh[i][0], bh[i][1]);
#pragma unroll
            for (int i = 0; i < NTW; ++i) mma_v(ACC[i], ah, bl[i][0], bl[i][1]);
#pragma unroll
            for (int i = 0; i < NTW; ++i) mma_v(ACC[i], ah, bh[i][0], bh[i][1]);
#else
#pragma unroll
            for (int i = 0; i < NTW; ++i) {
                mma_v(ACC[i], al, bh[i][0], bh[i][1]);
                mma_v(ACC[i], ah, bl[i][0], bl[i][1]);
                mma_v(ACC[i], ah, bh[i][0], bh[i][1]);
            }
#endif
#ifdef V_FRESH
#pragma unroll
            for (int i = 0; i < NTW; ++i) { add2(c[i][0], c[i][1], cp[i][0], cp[i][1]); add2(c[i][2], c[i][3], cp[i][2], cp[i][3]); }
#endif
        }
    };
    auto load = [&](float4 (&dst)[NTW], int kp) {
#pragma unroll
        for (int i = 0; i < NTW; ++i) dst[i] = __ldg(W + ((size_t)kp * NT + i * 8) * 32);
    };
    float4 wn[NTW];
    load(wn, 0);
#ifdef V_UNROLL2
    float4 wm[NTW];
    if (KT2 > 1) load(wm, 1);
    int kp = 0;
    for (; kp + 2 <= KT2; kp += 2) {
        do_pair(wn, kp);
        if (kp + 2 < KT2) load(wn, kp + 2);
        do_pair(wm, kp + 1);
        if (kp + 3 < KT2) load(wm, kp + 3);
    }
    if (kp < KT2) do_pair(wn, kp);
#elif defined(V_RING3)
    float4 wn2[NTW];
    load(wn2, 1);
    for (int kp = 0; kp < KT2; ++kp) {
        float4 w[NTW];
#pragma unroll
        for (int i = 0; i < NTW; ++i) { w[i] = wn[i]; wn[i] = wn2[i]; }
        if (kp + 2 < KT2) load(wn2, kp + 2);
        do_pair(w, kp);
    }
#else
    for (int kp = 0; kp < KT2; ++kp) {
        float4 w[NTW];
#pragma unroll
        for (int i = 0; i < NTW; ++i) w[i] = wn[i];
        if (kp + 1 < KT2) load(wn, kp + 1);
        do_pair(w, kp);
    }
#endif
#pragma unroll
    for (int i = 0; i < NTW; ++i) epi(warp + 8 * i, c[i]);
}

__global__ void __launch_bounds__(256, 1)
k_gemm_only(const float4* __restrict__ Wf, const float* __restrict__ bias, int W, int layers, int reps,
            float* out, long long* cycles) {
    constexpr int TP = 16, S = 24;
    float* h1 = fab_smem;
    float* h2 = fab_smem + (size_t)W * S;
    uint32_t* mask = reinterpret_cast<uint32_t*>(fab_smem + (size_t)2 * W * S);
    for (int i = threadIdx.x; i < 2 * W * S; i += 256) fab_smem[i] = 0.001f * (i % 97) - 0.04f;
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int KT2 = W / 16, NT = W / 8;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        for (int l = 0; l < layers; ++l) {
            float* src = (l & 1) ? h2 : h1;
            float* dst = (l & 1) ? h1 : h2;
            gemm_simple(src, KT2, Wf + (size_t)l * KT2 * NT * 32, NT, bias,
                        [&](int nt, const float (&c)[4]) { hidden_fwd<TP, true>(dst, mask, nt, g, t, c); });
            __syncthreads();
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * 256 + threadIdx.x] = h1[threadIdx.x];
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

int main() {
    const int W = 320, layers = 10, reps = 20, KT2 = W / 16, NT = W / 8;
    float4* Wf; CK(cudaMalloc(&Wf, (size_t)layers * KT2 * NT * 32 * sizeof(float4)));
    {
        std::vector<float> hw((size_t)layers * KT2 * NT * 128);
        for (size_t i = 0; i < hw.size(); ++i) hw[i] = 0.002f * ((int)(i * 2654435761u % 61) - 30);
        CK(cudaMemcpy(Wf, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice));
    }
    float* bias; CK(cudaMalloc(&bias, W * 4)); CK(cudaMemset(bias, 0, W * 4));
    float* out; CK(cudaMalloc(&out, 148 * 1024 * sizeof(float)));
    long long* cyc; CK(cudaMalloc(&cyc, 8));
    const size_t sm = ((size_t)2 * W * 24 + NT * 4) * sizeof(float);
    CK(cudaFuncSetAttribute(k_gemm_only, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    for (int rep = 0; rep < 2; ++rep) { k_gemm_only<<<147, 256, sm>>>(Wf, bias, W, layers, reps, out, cyc); CK(cudaDeviceSynchronize()); }
    long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("%.0f cycles per GEMM\n", (double)h / (layers * reps));
    return 0;
}
