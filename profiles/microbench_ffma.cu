// Microbenchmark (B200): FP32 FMA issue rates that bound the tile GEMM.
//   1. FFMA   : independent chains, registers only
//   2. FFMA2  : fma.rn.f32x2 (packed pair), registers only
//   3. step   : the tile-GEMM inner step fed from shared memory (LDS.128 broadcast + FFMA),
//               scalar and packed variants, T particles x TN columns per thread
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb profiles/microbench_ffma.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int NCH>
__global__ void k_ffma(float* out, int iters, float a, float b) {
    float acc[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) acc[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NCH>
__global__ void k_ffma2(float* out, int iters, float a, float b) {
    u64 acc[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) acc[i] = pack2(threadIdx.x * 0.001f + i, 1.f + i);
    const u64 a2 = pack2(a, a * 1.0001f), b2 = pack2(b, b * 0.999f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) acc[i] = ffma2(acc[i], a2, b2);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) { float x, y; unpack2(acc[i], x, y); s += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// scalar step: act layout [k4][T] float4 (k4-major), acc[T][TN]
template <int T, int TN>
__global__ void k_step_scalar(float* out, const float4* __restrict__ wglob, int steps, int reps) {
    extern __shared__ float4 act4[];       // [steps][T]
    for (int i = threadIdx.x; i < steps * T; i += blockDim.x)
        act4[i] = make_float4(0.001f * i, 0.002f, -0.001f, 0.0005f * (i & 7));
    __syncthreads();
    float acc[T][TN];
#pragma unroll
    for (int p = 0; p < T; ++p)
#pragma unroll
        for (int t = 0; t < TN; ++t) acc[p][t] = 0.f;
    float4 w[TN];
#pragma unroll
    for (int t = 0; t < TN; ++t) w[t] = wglob[threadIdx.x * TN + t];
    for (int r = 0; r < reps; ++r) {
        for (int s = 0; s < steps; ++s) {
            const float4* a4 = act4 + s * T;
#pragma unroll
            for (int p = 0; p + 1 < T; p += 2) {
                const float4 a0 = a4[p], a1 = a4[p + 1];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float x0 = c == 0 ? a0.x : c == 1 ? a0.y : c == 2 ? a0.z : a0.w;
                    const float x1 = c == 0 ? a1.x : c == 1 ? a1.y : c == 2 ? a1.z : a1.w;
#pragma unroll
                    for (int t = 0; t < TN; ++t) {
                        const float wv = c == 0 ? w[t].x : c == 1 ? w[t].y : c == 2 ? w[t].z : w[t].w;
                        acc[p][t] = fmaf(x0, wv, acc[p][t]);
                        acc[p + 1][t] = fmaf(x1, wv, acc[p + 1][t]);
                    }
                }
            }
        }
    }
    float sum = 0.f;
#pragma unroll
    for (int p = 0; p < T; ++p)
#pragma unroll
        for (int t = 0; t < TN; ++t) sum += acc[p][t];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
}

// packed step: act layout [k][TP] floats with particles fastest (TP = T rounded up to 4), so one
// LDS.128 yields two particle pairs of one k; acc2[T/2][TN] pairs over particles, weights
// duplicated into (w,w) pairs once per step.
template <int T, int TN>
__global__ void k_step_packed(float* out, const float4* __restrict__ wglob, int steps, int reps) {
    constexpr int TP = (T + 3) & ~3;
    extern __shared__ float4 act4[];       // [steps*4][TP/4] float4
    for (int i = threadIdx.x; i < steps * 4 * (TP / 4); i += blockDim.x)
        act4[i] = make_float4(0.001f * i, 0.002f, -0.001f, 0.0005f * (i & 7));
    __syncthreads();
    u64 acc[T / 2][TN];
#pragma unroll
    for (int p = 0; p < T / 2; ++p)
#pragma unroll
        for (int t = 0; t < TN; ++t) acc[p][t] = pack2(0.f, 0.f);
    float4 w[TN];
#pragma unroll
    for (int t = 0; t < TN; ++t) w[t] = wglob[threadIdx.x * TN + t];
    for (int r = 0; r < reps; ++r) {
        for (int s = 0; s < steps; ++s) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                u64 w2[TN];
#pragma unroll
                for (int t = 0; t < TN; ++t) {
                    const float wv = c == 0 ? w[t].x : c == 1 ? w[t].y : c == 2 ? w[t].z : w[t].w;
                    w2[t] = pack2(wv, wv);
                }
                const u64* arow = reinterpret_cast<const u64*>(act4 + (s * 4 + c) * (TP / 4));
#pragma unroll
                for (int q = 0; q < T / 2; q += 2) {
                    // one LDS.128 = two particle pairs (or one LDS.64 for a trailing pair)
                    u64 p0, p1 = 0;
                    if (q + 1 < T / 2) {
                        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(arow + q);
                        p0 = v.x; p1 = v.y;
                    } else {
                        p0 = arow[q];
                    }
#pragma unroll
                    for (int t = 0; t < TN; ++t) {
                        acc[q][t] = ffma2(p0, w2[t], acc[q][t]);
                        if (q + 1 < T / 2) acc[q + 1][t] = ffma2(p1, w2[t], acc[q + 1][t]);
                    }
                }
            }
        }
    }
    float sum = 0.f;
#pragma unroll
    for (int p = 0; p < T / 2; ++p)
#pragma unroll
        for (int t = 0; t < TN; ++t) { float x, y; unpack2(acc[p][t], x, y); sum += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
}

template <typename F>
float time_ms(F launch) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(a);
    launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
    int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    const double peak = nsm * 128.0 * 2 * clk * 1e3 / 1e12;
    printf("SMs %d, clock %d kHz, FP32 peak (128 lanes) %.1f TFLOP/s\n", nsm, clk, peak);
    float* out; CK(cudaMalloc(&out, 148 * 1024 * 8 * sizeof(float)));
    float4* w; CK(cudaMalloc(&w, 1024 * 8 * sizeof(float4))); CK(cudaMemset(w, 0, 1024 * 8 * sizeof(float4)));
    const int iters = 4096;
    for (int nt : {128, 256, 320, 512, 640, 1024}) {
        float ms = time_ms([&] { k_ffma<16><<<nsm, nt>>>(out, iters, 1.0001f, 0.5f); });
        double tf = 2.0 * 16 * iters * nt * nsm / (ms * 1e-3) / 1e12;
        float ms2 = time_ms([&] { k_ffma2<16><<<nsm, nt>>>(out, iters, 1.0001f, 0.5f); });
        double tf2 = 2.0 * 2 * 16 * iters * nt * nsm / (ms2 * 1e-3) / 1e12;
        printf("threads/SM %4d : FFMA %.1f TF (%.0f%%)   FFMA2 %.1f TF (%.0f%%)\n", nt, tf, 100 * tf / peak, tf2, 100 * tf2 / peak);
    }
    const int steps = 80, reps = 40;
#define RUN_STEP(T, TN, NT)                                                                              \
    {                                                                                                    \
        size_t sm = (size_t)steps * 4 * (((T) + 3) & ~3) * sizeof(float);                                \
        cudaFuncSetAttribute(k_step_scalar<T, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
        cudaFuncSetAttribute(k_step_packed<T, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
        float ms = time_ms([&] { k_step_scalar<T, TN><<<nsm, NT, sm>>>(out, w, steps, reps); });          \
        double fl = 2.0 * 4 * T * TN * steps * reps * (double)NT * nsm;                                   \
        float ms2 = time_ms([&] { k_step_packed<T, TN><<<nsm, NT, sm>>>(out, w, steps, reps); });         \
        printf("step T=%2d TN=%d NT=%4d : scalar %.1f TF (%.0f%%)   packed %.1f TF (%.0f%%)\n", T, TN, NT, \
               fl / (ms * 1e-3) / 1e12, 100 * fl / (ms * 1e-3) / 1e12 / peak, fl / (ms2 * 1e-3) / 1e12,   \
               100 * fl / (ms2 * 1e-3) / 1e12 / peak);                                                   \
    }
    RUN_STEP(14, 4, 320)
    RUN_STEP(14, 2, 640)
    RUN_STEP(16, 4, 320)
    RUN_STEP(16, 2, 640)
    RUN_STEP(16, 4, 512)
    RUN_STEP(8, 4, 640)
    RUN_STEP(8, 8, 320)
    RUN_STEP(16, 2, 320)
    cudaError_t e = cudaDeviceSynchronize();
    printf("done: %s\n", cudaGetErrorString(e));
    return 0;
}
