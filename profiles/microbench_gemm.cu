// Microbenchmark (B200): the product's tile_gemm<T> in isolation -- 147 CTAs stream the hidden
// layer weights (WP x WP, K layers) from L2 exactly as flow_inverse does, no epilogues.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I include -I fab_torch_b200/csrc \
//        [-DFAB_NT=.. -DFAB_TN=.. -DFAB_PF=..] -o profiles/mb_gemm profiles/microbench_gemm.cu
#include <cstdio>
#include <vector>
#include "tile_gemm.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int T>
__global__ void __launch_bounds__(FAB_NT, 1)
k_gemm_only(const float4* __restrict__ W, int WP, int K4, int layers, int reps, int red_floats,
            float* out, int sync_mode) {
    extern __shared__ __align__(16) float smem[];
    constexpr int TP = TileDims<T>::TP;
    float* act = smem;                             // [K4*4][TP]
    float* red = smem + (size_t)K4 * 4 * TP;
    for (int i = threadIdx.x; i < K4 * 4 * TP; i += FAB_NT) act[i] = 0.001f * (i % 97) - 0.04f;
    __syncthreads();
    float s = 0.f;
    for (int r = 0; r < reps; ++r) {
        for (int l = 0; l < layers; ++l) {
            const int KS = tile_gemm<T>(act, K4, W + (size_t)l * K4 * WP, WP, red, red_floats);
            if (sync_mode) __syncthreads();
            s += red[threadIdx.x] * (float)KS;
            if (sync_mode) __syncthreads();
        }
    }
    out[blockIdx.x * FAB_NT + threadIdx.x] = s;
}

int main() {
    int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
    int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    const double peak = nsm * 128.0 * 2 * clk * 1e3 / 1e12;
    const int WP = 320, K4 = WP / 4 + 1, layers = 10, reps = 20;
    float4* W; CK(cudaMalloc(&W, (size_t)layers * K4 * WP * sizeof(float4)));
    CK(cudaMemset(W, 0, (size_t)layers * K4 * WP * sizeof(float4)));
    float* out; CK(cudaMalloc(&out, 148 * 1024 * sizeof(float)));
    printf("FAB_NT=%d FAB_TN=%d FAB_PF=%d  peak %.1f TF\n", FAB_NT, FAB_TN, FAB_PF, peak);
#define RUN(T, GRID)                                                                                    \
    {                                                                                                    \
        constexpr int TP = TileDims<T>::TP;                                                              \
        int ks = FAB_NT / (WP / FAB_TN); if (ks < 1) ks = 1; if (ks > 8) ks = 8;                          \
        int red_floats = ks * T * (WP + 4);                                                              \
        size_t sm = ((size_t)K4 * 4 * TP + red_floats) * sizeof(float);                                  \
        if (sm <= 227 * 1024) {                                                                          \
        CK(cudaFuncSetAttribute(k_gemm_only<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));   \
        for (int mode = 0; mode < 2; ++mode) {                                                           \
            k_gemm_only<T><<<GRID, FAB_NT, sm>>>(W, WP, K4, layers, 2, red_floats, out, mode);            \
            CK(cudaDeviceSynchronize());                                                                 \
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);                                   \
            cudaEventRecord(a);                                                                          \
            k_gemm_only<T><<<GRID, FAB_NT, sm>>>(W, WP, K4, layers, reps, red_floats, out, mode);         \
            cudaEventRecord(b); CK(cudaEventSynchronize(b));                                              \
            float ms; cudaEventElapsedTime(&ms, a, b);                                                   \
            double fl = 2.0 * T * (double)WP * (K4 * 4) * layers * reps * GRID;                           \
            printf("T=%2d grid=%3d KS=%d sync=%d: %.3f ms  %.1f TF (%.0f%% of peak, %.0f%% of the SMs' share)\n", T, GRID, \
                   ks, mode, ms, fl / (ms * 1e-3) / 1e12, 100 * fl / (ms * 1e-3) / 1e12 / peak,            \
                   100 * fl / (ms * 1e-3) / 1e12 / (peak * GRID / nsm));                                  \
        } } else printf("T=%d: smem %zu too large\n", T, sm);                                             \
    }
    RUN(14, 147)
    RUN(16, 128)
    RUN(8, 148)
    RUN(14, 1)
    printf("done\n");
    return 0;
}
