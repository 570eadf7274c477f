// How does the warp-level TF32 MMA (HMMA.1688.F32.TF32) round on B200?  Single-instruction probes.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/mb_mma_round profiles/microbench_mma_rounding.cu
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// out[0] = C + sum_k a[k]*b[k] for row 0 / column 0 of the tile
__global__ void probe(const float* a8, const float* b8, float cin, float* out) {
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    uint32_t a[4] = {0, 0, 0, 0}, b[2] = {0, 0};
    if (g == 0) { a[0] = __float_as_uint(a8[t]); a[2] = __float_as_uint(a8[t + 4]); }   // row 0
    if (g == 0) { b[0] = __float_as_uint(b8[t]); b[1] = __float_as_uint(b8[t + 4]); }   // col 0
    float c[4] = {0, 0, 0, 0};
    if (lane == 0) c[0] = cin;
    mma_tf32(c, a, b);
    if (lane == 0) out[0] = c[0];
}
int main() {
    float *da, *db, *dout; cudaMalloc(&da, 32); cudaMalloc(&db, 32); cudaMalloc(&dout, 4);
    auto run = [&](const char* name, const float (&a)[8], const float (&b)[8], float c, double exact) {
        cudaMemcpy(da, a, 32, cudaMemcpyHostToDevice); cudaMemcpy(db, b, 32, cudaMemcpyHostToDevice);
        probe<<<1, 32>>>(da, db, c, dout); float r; cudaMemcpy(&r, dout, 4, cudaMemcpyDeviceToHost);
        const float rn = (float)exact; const float dn = nextafterf(rn, rn > exact ? -INFINITY : INFINITY);
        printf("%-44s exact %.10e  got %.10e  (err %+.3f ulp; RN would give %.10e)\n", name, exact, r,
               (r - exact) / (double)fabsf(rn - dn), rn);
    };
    const float u = ldexpf(1.f, -23);     // ulp of [1,2)
    const float ones[8] = {1, 1, 1, 1, 1, 1, 1, 1};
    { float a[8] = {1.f, 1.5f * u * 1024, 0, 0, 0, 0, 0, 0}; float b[8] = {1.f, 1.f / 1024, 0, 0, 0, 0, 0, 0};
      run("C=0: 1 + 1.5ulp", a, b, 0.f, 1.0 + 1.5 * u); }
    { float a[8] = {-1.f, -1.5f * u * 1024, 0, 0, 0, 0, 0, 0}; float b[8] = {1.f, 1.f / 1024, 0, 0, 0, 0, 0, 0};
      run("C=0: -1 - 1.5ulp", a, b, 0.f, -1.0 - 1.5 * u); }
    { float a[8] = {1.f, 1.75f * u * 1024, 0, 0, 0, 0, 0, 0}; float b[8] = {1.f, 1.f / 1024, 0, 0, 0, 0, 0, 0};
      run("C=0: 1 + 1.75ulp", a, b, 0.f, 1.0 + 1.75 * u); }
    { float a[8] = {-1.f, -1.75f * u * 1024, 0, 0, 0, 0, 0, 0}; float b[8] = {1.f, 1.f / 1024, 0, 0, 0, 0, 0, 0};
      run("C=0: -1 - 1.75ulp", a, b, 0.f, -1.0 - 1.75 * u); }
    { float a[8] = {1.75f * u * 1024, 0, 0, 0, 0, 0, 0, 0}; float b[8] = {1.f / 1024, 0, 0, 0, 0, 0, 0, 0};
      run("C=1: 1 + 1.75ulp (C + product)", a, b, 1.f, 1.0 + 1.75 * u);
      run("C=-1: -1 + 1.75ulp", a, b, -1.f, -1.0 + 1.75 * u); }
    { float a[8] = {-1.75f * u * 1024, 0, 0, 0, 0, 0, 0, 0}; float b[8] = {1.f / 1024, 0, 0, 0, 0, 0, 0, 0};
      run("C=-1: -1 - 1.75ulp", a, b, -1.f, -1.0 - 1.75 * u);
      run("C=1: 1 - 1.75ulp", a, b, 1.f, 1.0 - 1.75 * u); }
    { float a[8] = {1.f, -1.f, ldexpf(1.f, -30), 0, 0, 0, 0, 0};
      run("C=0: 1 - 1 + 2^-30 (exact inner sum?)", a, ones, 0.f, ldexp(1.0, -30)); }
    { float a[8] = {1.f, ldexpf(1.f, -26), ldexpf(1.f, -26), ldexpf(1.f, -26), ldexpf(1.f, -26), ldexpf(1.f, -26), ldexpf(1.f, -26), ldexpf(1.f, -26)};
      run("C=0: 1 + 7*2^-26 (=1+0.875ulp): kept bits?", a, ones, 0.f, 1.0 + 7 * ldexp(1.0, -26)); }
    { float a[8] = {1.f, ldexpf(1.f, -24), ldexpf(1.f, -24), ldexpf(1.f, -24), ldexpf(1.f, -24), 0, 0, 0};
      run("C=0: 1 + 4*2^-24 (=1+2ulp, each below ulp)", a, ones, 0.f, 1.0 + 4 * ldexp(1.0, -24)); }
    { float a[8] = {ldexpf(1.f, -24), ldexpf(1.f, -24), ldexpf(1.f, -24), ldexpf(1.f, -24), 0, 0, 0, 0};
      run("C=1: 1 + 4*2^-24 (products below ulp(C))", a, ones, 1.f, 1.0 + 4 * ldexp(1.0, -24)); }
    return 0;
}
