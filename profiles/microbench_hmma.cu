// Microbenchmark (B200): is the legacy warp-level tensor path (mma.sync m16n8k8 TF32, SASS HMMA)
// a faster engine than the packed-FP32 tile GEMM for ONE CTA's 14..16 particles?
//
//  (1) raw issue rate of mma.sync.m16n8k8.tf32 per SM sub-partition (independent accumulators)
//  (2) a 3xTF32 (error-compensated, ~fp32 accurate) [16 x 320] x [320 x 320] tile GEMM: weights
//      stream from L2 in fragment order, activations come from shared memory, hi/lo split on the
//      fly -- the job the first (FP32 FFMA2) version of the tile GEMM did in ~21.6k cycles.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/mb_hmma profiles/microbench_hmma.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int ILP>
__global__ void k_raw(int iters, float* out, long long* cycles) {
    float c[ILP][4];
    uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f810000u, 0x3f820000u, 0x3f830000u};
    uint32_t b[2] = {0x3f000000u, 0x3f010000u + threadIdx.x};
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) mma_tf32(c[i], a, b);
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// ---- 3xTF32 tile GEMM -------------------------------------------------------------------------
constexpr int S = 24;                 // activation row stride (floats): conflict-free fragment loads
constexpr int NTW = 5;                // n-tiles (8 columns each) per warp: 8 warps x 5 x 8 = 320

__device__ __forceinline__ void split(float v, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(v) & 0xffffe000u;                  // top 10 mantissa bits (tf32 grid)
    lo = __float_as_uint(v - __uint_as_float(hi));          // exact remainder; HW drops its low bits
}

// Wf: fragment-ordered weights [KT/2][NT][32 lanes][4] = (b0,b1 of even k-tile, b0,b1 of odd k-tile)
template <bool PRESPLIT_A>
__global__ void __launch_bounds__(256, 1)
k_gemm3(const float4* __restrict__ Wf, int KT, int NT, int layers, int reps, float* out,
        long long* cycles) {
    extern __shared__ __align__(16) float smem[];
    float* act = smem;                                // [KT*8][S]   (hi plane when PRESPLIT_A)
    float* act_lo = smem + (size_t)KT * 8 * S;        // lo plane
    float* hout = act_lo + (size_t)KT * 8 * S;        // [NT*8][S] epilogue target
    for (int i = threadIdx.x; i < KT * 8 * S; i += 256) {
        const float v = 0.001f * (i % 97) - 0.04f;
        uint32_t h, l; split(v, h, l);
        act[i] = PRESPLIT_A ? __uint_as_float(h) : v;
        act_lo[i] = __uint_as_float(l);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        for (int l = 0; l < layers; ++l) {
            const float4* W = Wf + (size_t)l * (KT / 2) * NT * 32;
            float c[NTW][4];
#pragma unroll
            for (int i = 0; i < NTW; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
            float4 wn[NTW];
#pragma unroll
            for (int i = 0; i < NTW; ++i) wn[i] = __ldg(W + ((size_t)0 * NT + warp * NTW + i) * 32 + lane);
            for (int k2 = 0; k2 < KT / 2; ++k2) {
                float4 w[NTW];
#pragma unroll
                for (int i = 0; i < NTW; ++i) w[i] = wn[i];
                if (k2 + 1 < KT / 2) {
#pragma unroll
                    for (int i = 0; i < NTW; ++i)
                        wn[i] = __ldg(W + ((size_t)(k2 + 1) * NT + warp * NTW + i) * 32 + lane);
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int kb = (k2 * 2 + h) * 8;
                    uint32_t ah[4], al[4];
                    const float* ap = act + (size_t)(kb + t) * S + g;
                    if (PRESPLIT_A) {
                        const float* lp = act_lo + (size_t)(kb + t) * S + g;
                        ah[0] = __float_as_uint(ap[0]); ah[1] = __float_as_uint(ap[8]);
                        ah[2] = __float_as_uint(ap[4 * S]); ah[3] = __float_as_uint(ap[4 * S + 8]);
                        al[0] = __float_as_uint(lp[0]); al[1] = __float_as_uint(lp[8]);
                        al[2] = __float_as_uint(lp[4 * S]); al[3] = __float_as_uint(lp[4 * S + 8]);
                    } else {
                        split(ap[0], ah[0], al[0]); split(ap[8], ah[1], al[1]);
                        split(ap[4 * S], ah[2], al[2]); split(ap[4 * S + 8], ah[3], al[3]);
                    }
#pragma unroll
                    for (int i = 0; i < NTW; ++i) {
                        uint32_t bh[2], bl[2];
                        split(h == 0 ? w[i].x : w[i].z, bh[0], bl[0]);
                        split(h == 0 ? w[i].y : w[i].w, bh[1], bl[1]);
                        mma_tf32(c[i], al, bh);       // small terms first
                        mma_tf32(c[i], ah, bl);
                        mma_tf32(c[i], ah, bh);
                    }
                }
            }
            // epilogue straight from the accumulator fragments: relu -> next operand (k-major)
#pragma unroll
            for (int i = 0; i < NTW; ++i) {
                const int n = (warp * NTW + i) * 8 + 2 * t;
                hout[(size_t)n * S + g] = fmaxf(c[i][0], 0.f);
                hout[(size_t)(n + 1) * S + g] = fmaxf(c[i][1], 0.f);
                hout[(size_t)n * S + g + 8] = fmaxf(c[i][2], 0.f);
                hout[(size_t)(n + 1) * S + g + 8] = fmaxf(c[i][3], 0.f);
            }
            __syncthreads();
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * 256 + threadIdx.x] = hout[threadIdx.x];
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

int main() {
    float* out; CK(cudaMalloc(&out, 148 * 1024 * sizeof(float)));
    long long* cyc; CK(cudaMalloc(&cyc, 8));
    long long h;
    const int iters = 4096;
#define RAW(ILP, NTHR)                                                                            \
    {                                                                                             \
        k_raw<ILP><<<148, NTHR>>>(iters, out, cyc);                                               \
        CK(cudaDeviceSynchronize());                                                              \
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));                                       \
        const double per_smsp = (double)h / ((double)iters * ILP * (NTHR / 32) / 4.0);            \
        printf("raw HMMA.1688.TF32  warps/SM=%2d ILP=%d : %.2f cycles per HMMA per sub-partition " \
               "(%.0f dense TF32 TFLOP/s chip-wide at 1.965 GHz)\n", NTHR / 32, ILP, per_smsp,     \
               148 * 4 * 2048.0 / per_smsp * 1.965e9 / 1e12);                                     \
    }
    RAW(1, 128) RAW(2, 128) RAW(4, 128) RAW(8, 128)
    RAW(1, 256) RAW(2, 256) RAW(4, 256) RAW(8, 256)
    RAW(4, 512) RAW(8, 512)

    const int KT = 40, NT = 40, layers = 10, reps = 20;
    float4* W; CK(cudaMalloc(&W, (size_t)layers * (KT / 2) * NT * 32 * sizeof(float4)));
    {
        std::vector<float> hw((size_t)layers * (KT / 2) * NT * 32 * 4);
        for (size_t i = 0; i < hw.size(); ++i) hw[i] = 0.01f * ((int)(i % 61) - 30);
        CK(cudaMemcpy(W, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice));
    }
    const size_t sm = ((size_t)KT * 8 * S * 2 + (size_t)NT * 8 * S) * sizeof(float);
    CK(cudaFuncSetAttribute(k_gemm3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    CK(cudaFuncSetAttribute(k_gemm3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    for (int pres = 0; pres < 2; ++pres)
        for (int grid : {1, 147}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (pres) k_gemm3<true><<<grid, 256, sm>>>(W, KT, NT, layers, reps, out, cyc);
                else k_gemm3<false><<<grid, 256, sm>>>(W, KT, NT, layers, reps, out, cyc);
                CK(cudaDeviceSynchronize());
            }
            CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
            printf("3xTF32 tile GEMM [16x320]x[320x320], presplit_A=%d grid=%3d: %.0f cycles per GEMM "
                   "(first-version FFMA2 tile GEMM: ~21600)\n", pres, grid, (double)h / (layers * reps));
        }
    printf("done\n");
    return 0;
}
