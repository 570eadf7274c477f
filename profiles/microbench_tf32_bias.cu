// Calibration: mean relative shrink of ONE 3xTF32 k-tile partial sum (8 terms, fresh accumulator,
// small terms first, lo operands rounded to the tf32 grid) on the warp-level tensor path.
// The tensor core truncates toward zero, so E[(got - exact)/exact] < 0; the product kernels
// undo the mean with c = fma(cp, 1 + kappa*2^-24, c) (mma_gemm.cuh).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/mb_tf32bias profiles/microbench_tf32_bias.cu
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <vector>
#include <random>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void split(float v, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(v) & 0xffffe000u;
    lo = __float_as_uint(v - __uint_as_float(hi)) + 0x1000u;
}
// A [16][8], B [8][N] -> C [16][N] single k-tile
__global__ void k_tile(const float* A, const float* B, float* C, int N) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int nt = blockIdx.x * (blockDim.x / 32) + warp; nt < N / 8; nt += gridDim.x * (blockDim.x / 32)) {
        float av[4] = {A[g * 8 + t], A[(g + 8) * 8 + t], A[g * 8 + t + 4], A[(g + 8) * 8 + t + 4]};
        float bv[2] = {B[t * N + nt * 8 + g], B[(t + 4) * N + nt * 8 + g]};
        uint32_t ah[4], al[4], bh[2], bl[2];
        for (int i = 0; i < 4; ++i) split(av[i], ah[i], al[i]);
        for (int i = 0; i < 2; ++i) split(bv[i], bh[i], bl[i]);
        float c[4] = {0, 0, 0, 0};
        mma_tf32(c, al, bh); mma_tf32(c, ah, bl); mma_tf32(c, ah, bh);
        C[g * N + nt * 8 + 2 * t] = c[0]; C[g * N + nt * 8 + 2 * t + 1] = c[1];
        C[(g + 8) * N + nt * 8 + 2 * t] = c[2]; C[(g + 8) * N + nt * 8 + 2 * t + 1] = c[3];
    }
}
int main() {
    const int N = 1 << 16;
    std::mt19937 rng(7);
    std::normal_distribution<float> nd(0.f, 1.f);
    for (int mode = 0; mode < 3; ++mode) {
        std::vector<float> hA(16 * 8), hB(8 * N), hC(16 * N);
        for (auto& v : hA) { v = nd(rng); if (mode >= 1) v = fabsf(v); }      // 1: relu-like A, signed B; 2: all positive
        for (auto& v : hB) { v = nd(rng) * 0.05f; if (mode == 2) v = fabsf(v); }
        float *A, *B, *C; cudaMalloc(&A, hA.size() * 4); cudaMalloc(&B, hB.size() * 4); cudaMalloc(&C, hC.size() * 4);
        cudaMemcpy(A, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(B, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
        k_tile<<<148, 256>>>(A, B, C, N);
        cudaMemcpy(hC.data(), C, hC.size() * 4, cudaMemcpyDeviceToHost);
        double sum_rel = 0, sum_w = 0, sq = 0, num = 0, den = 0; long cnt = 0;
        for (int p = 0; p < 16; ++p) for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < 8; ++k) s += (double)hA[p * 8 + k] * hB[k * N + n];
            const double e = hC[p * N + n] - s;
            if (fabs(s) > 1e-3) { sum_rel += e / s; sq += (e / s) * (e / s); ++cnt; }
            num += e * s; den += s * s;                       // least-squares shrink factor
        }
        printf("%s: mean rel err %+.3e (= %+.3f * 2^-24), rms %.3e, least-squares shrink %+.3e (= %+.3f * 2^-24)\n",
               mode == 0 ? "signed x signed " : mode == 1 ? "relu(A) x signed" : "positive        ",
               sum_rel / cnt, sum_rel / cnt * 16777216.0, sqrt(sq / cnt), num / den, num / den * 16777216.0);
        cudaFree(A); cudaFree(B); cudaFree(C);
    }
    return 0;
}
