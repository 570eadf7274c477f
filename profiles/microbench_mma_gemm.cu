// Microbenchmark (B200): the product's mma_gemm_wide<16> in isolation -- `grid` CTAs stream the
// hidden-layer weights (W x W, `layers` layers) from L2 exactly as flow_inverse's second GEMM does,
// with the ReLU + mask-ballot epilogue, one barrier per GEMM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I include -I fab_torch_b200/csrc \
//        [-DFAB_NT=256|512 ...] -o profiles/mb_mma_gemm profiles/microbench_mma_gemm.cu
#include <cstdio>
#include <vector>
#include "flow_tile.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void __launch_bounds__(FAB_NT, 1)
k_gemm_only(const float4* __restrict__ Wf, const float* __restrict__ bias, int W, int layers, int reps,
            float* out, long long* cycles) {
    constexpr int TP = 16, S = ActL<TP>::S;
    float* h1 = fab_smem;                          // [W][S]
    float* h2 = fab_smem + (size_t)W * S;
    uint32_t* mask = reinterpret_cast<uint32_t*>(fab_smem + (size_t)2 * W * S);
    for (int i = threadIdx.x; i < 2 * W * S; i += FAB_NT) fab_smem[i] = 0.001f * (i % 97) - 0.04f;
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int KT2 = W / 16, NT = W / 8;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        for (int l = 0; l < layers; ++l) {
            float* src = (l & 1) ? h2 : h1;
            float* dst = (l & 1) ? h1 : h2;
            mma_gemm_wide<TP>(src, KT2, Wf + (size_t)l * KT2 * NT * 32, NT, bias,
                              [&](int nt, const float (&c)[4]) { hidden_fwd<TP, true>(dst, mask, nt, g, t, c); });
            __syncthreads();
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * FAB_NT + threadIdx.x] = h1[threadIdx.x];
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
    for (int grid : {1, 147}) {
        for (int rep = 0; rep < 2; ++rep) {
            k_gemm_only<<<grid, FAB_NT, sm>>>(Wf, bias, W, layers, reps, out, cyc);
            CK(cudaDeviceSynchronize());
        }
        long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("FAB_NT=%d grid=%3d: %.0f cycles per [16x%d]x[%dx%d] GEMM + epilogue + barrier "
               "(tensor-pipe floor 9600)\n", FAB_NT, grid, (double)h / (layers * reps), W, W, W);
    }
    return 0;
}
