// Accuracy of error-compensated TF32 tile GEMMs on the warp-level tensor path (B200), versus an
// fp64 host reference and a plain fp32 FMA loop: C[16x320] = A[16x320] * B[320x320].
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/mb_tf32acc profiles/microbench_tf32_accuracy.cu
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <vector>
#include <random>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void split_trunc(float v, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(v) & 0xffffe000u;
    lo = __float_as_uint(v - __uint_as_float(hi));
}
__device__ __forceinline__ void split_rn(float v, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
    const float r = v - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
// variant: 0 = 1xTF32 (trunc), 1 = 3x trunc one acc (small first), 2 = 3x RN one acc (small first),
// 3 = 3x RN, small terms in their own accumulator, 4 = 4x RN (adds lo*lo) separate small acc,
// 5 = 3x RN one acc (big first), 6 = 3x trunc separate small acc,
// 7 / 8 = 3x trunc into a FRESH accumulator per 2 / 4 k-tiles, flushed into the total with a
// round-to-nearest FADD (the tensor core truncates every accumulation: short chains keep the bias small)
template <int V>
__global__ void k_acc(const float* A, const float* B, float* C, int K, int N) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int nt = warp; nt < N / 8; nt += blockDim.x / 32) {
        float c[4] = {0, 0, 0, 0}, s[4] = {0, 0, 0, 0};
        for (int kb = 0; kb < K; kb += 8) {
            float av[4] = {A[g * K + kb + t], A[(g + 8) * K + kb + t], A[g * K + kb + t + 4], A[(g + 8) * K + kb + t + 4]};
            float bv[2] = {B[(kb + t) * N + nt * 8 + g], B[(kb + t + 4) * N + nt * 8 + g]};
            uint32_t ah[4], al[4], bh[2], bl[2];
            for (int i = 0; i < 4; ++i) { if (V == 0 || V == 1 || V >= 6) split_trunc(av[i], ah[i], al[i]); else split_rn(av[i], ah[i], al[i]); }
            for (int i = 0; i < 2; ++i) { if (V == 0 || V == 1 || V >= 6) split_trunc(bv[i], bh[i], bl[i]); else split_rn(bv[i], bh[i], bl[i]); }
            if (V == 0) { mma_tf32(c, ah, bh); }
            else if (V == 1 || V == 2) { mma_tf32(c, al, bh); mma_tf32(c, ah, bl); mma_tf32(c, ah, bh); }
            else if (V == 3 || V == 6) { mma_tf32(s, al, bh); mma_tf32(s, ah, bl); mma_tf32(c, ah, bh); }
            else if (V == 4) { mma_tf32(s, al, bl); mma_tf32(s, al, bh); mma_tf32(s, ah, bl); mma_tf32(c, ah, bh); }
            else if (V == 5) { mma_tf32(c, ah, bh); mma_tf32(c, al, bh); mma_tf32(c, ah, bl); }
            else if (V == 7 || V == 8) {
                mma_tf32(s, al, bh); mma_tf32(s, ah, bl); mma_tf32(s, ah, bh);
                const int J = V == 7 ? 2 : 4;
                if (((kb / 8) + 1) % J == 0) { for (int i = 0; i < 4; ++i) { c[i] += s[i]; s[i] = 0.f; } }
            }
        }
        for (int i = 0; i < 4; ++i) c[i] += s[i];
        C[g * N + nt * 8 + 2 * t] = c[0]; C[g * N + nt * 8 + 2 * t + 1] = c[1];
        C[(g + 8) * N + nt * 8 + 2 * t] = c[2]; C[(g + 8) * N + nt * 8 + 2 * t + 1] = c[3];
    }
}
__global__ void k_fma(const float* A, const float* B, float* C, int K, int N) {
    for (int e = threadIdx.x; e < 16 * N; e += blockDim.x) {
        const int p = e / N, n = e % N;
        float s = 0.f;
        for (int k = 0; k < K; ++k) s = fmaf(A[p * K + k], B[k * N + n], s);
        C[e] = s;
    }
}
int main() {
    const int K = 320, N = 320;
    std::mt19937 rng(1);
    std::normal_distribution<float> nd(0.f, 1.f);
    for (int mode = 0; mode < 2; ++mode) {
        std::vector<float> hA(16 * K), hB(K * N), hC(16 * N);
        for (auto& v : hA) v = mode == 0 ? nd(rng) : fabsf(nd(rng));          // mode 1: all positive (no cancellation, relu-like)
        for (auto& v : hB) v = (mode == 0 ? nd(rng) : fabsf(nd(rng))) * 0.05f;
        std::vector<double> ref(16 * N), mag(16 * N);
        for (int p = 0; p < 16; ++p) for (int n = 0; n < N; ++n) {
            double s = 0, m = 0;
            for (int k = 0; k < K; ++k) { s += (double)hA[p * K + k] * hB[k * N + n]; m += fabs((double)hA[p * K + k] * hB[k * N + n]); }
            ref[p * N + n] = s; mag[p * N + n] = m;
        }
        float *A, *B, *C; CK(cudaMalloc(&A, hA.size() * 4)); CK(cudaMalloc(&B, hB.size() * 4)); CK(cudaMalloc(&C, hC.size() * 4));
        CK(cudaMemcpy(A, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(B, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
        const char* names[10] = {"1xTF32 trunc", "3xTF32 trunc, one acc, small first", "3xTF32 RN, one acc, small first",
                                "3xTF32 RN, small terms own acc", "4xTF32 RN, small terms own acc", "3xTF32 RN, one acc, big first",
                                "3xTF32 trunc, small terms own acc", "3xTF32 trunc, fresh acc per 2 k-tiles",
                                "3xTF32 trunc, fresh acc per 4 k-tiles", "fp32 FMA chain"};
        printf("%s operands: error relative to sum|a||b| (max / rms), signed mean\n", mode == 0 ? "signed" : "positive");
        for (int v = 0; v < 10; ++v) {
            switch (v) {
                case 0: k_acc<0><<<1, 256>>>(A, B, C, K, N); break; case 1: k_acc<1><<<1, 256>>>(A, B, C, K, N); break;
                case 2: k_acc<2><<<1, 256>>>(A, B, C, K, N); break; case 3: k_acc<3><<<1, 256>>>(A, B, C, K, N); break;
                case 4: k_acc<4><<<1, 256>>>(A, B, C, K, N); break; case 5: k_acc<5><<<1, 256>>>(A, B, C, K, N); break;
                case 6: k_acc<6><<<1, 256>>>(A, B, C, K, N); break; case 7: k_acc<7><<<1, 256>>>(A, B, C, K, N); break;
                case 8: k_acc<8><<<1, 256>>>(A, B, C, K, N); break; default: k_fma<<<1, 256>>>(A, B, C, K, N); break;
            }
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(hC.data(), C, hC.size() * 4, cudaMemcpyDeviceToHost));
            double mx = 0, sq = 0, mean = 0;
            for (int i = 0; i < 16 * N; ++i) { const double e = (hC[i] - ref[i]) / mag[i]; mx = fmax(mx, fabs(e)); sq += e * e; mean += e; }
            printf("  %-36s max %.3e  rms %.3e  mean %+.3e\n", names[v], mx, sqrt(sq / (16 * N)), mean / (16 * N));
        }
        cudaFree(A); cudaFree(B); cudaFree(C);
    }
    return 0;
}
