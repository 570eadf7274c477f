// tcgen05 / TMEM microbenchmark for the row-tile engine (round 2).  One generic kernel executes a
// host-built list of tcgen05.mma operations on a host-built shared-memory image (operands already
// in the no-swizzle K-major core-matrix layout the engine uses) and dumps the accumulator columns
// of tensor memory.  The host side uses it for
//   1. descriptor / layout validation (cta_group::1 M=128, cta_group::2 M=128 -> the "2x2" layout),
//   2. issue-rate measurements (cycles per MMA for the shapes the engine can use),
//   3. the accumulate-rounding behaviour of the tensor core (round-to-nearest or truncation),
//   4. accuracy of the three-product splits (f16 hi/lo with a per-row power-of-two scale, tf32
//      hi/lo) against fp64 and against plain fp32 FMA arithmetic.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/mb_umma profiles/microbench_umma.cu
#define FAB_UMMA_WATCHDOG
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <random>
#include "../fab_torch_b200/csrc/umma.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct Op { uint32_t a_off, b_off, d_col, acc; };

struct Params {
    const uint8_t* image;      // per CTA: image + rank * image_bytes
    uint32_t image_bytes;
    const Op* ops;
    uint32_t n_ops, reps;
    uint32_t idesc, lbo_a, lbo_b;
    float* dump;               // per CTA [128][ncols]
    uint32_t ncols;
    unsigned long long* cycles;
    uint32_t ld16_off8;        // 1: dump columns [8, ncols-8) with x16 loads whose column is 8 mod 16
};

extern __shared__ __align__(1024) uint8_t smem_raw[];

template <int CG, bool KIND16>
__global__ void __launch_bounds__(128, 1) k_run(Params p) {
    __shared__ __align__(8) uint64_t bar_load, bar_done;
    __shared__ uint32_t tmem_slot;
    __shared__ Op s_ops[256];
    for (uint32_t i = threadIdx.x; i < p.n_ops && i < 256; i += blockDim.x) s_ops[i] = p.ops[i];
    const uint32_t rank = CG == 2 ? umma::cluster_ctarank() : 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        umma::mbar_init(&bar_load, 1);
        umma::mbar_init(&bar_done, 1);
        umma::mbar_fence_init();
    }
    if (warp == 0) umma::tmem_alloc<CG>(&tmem_slot, 512);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tbase = tmem_slot;
    if (threadIdx.x == 0) {
        const uint8_t* src = p.image + (size_t)rank * p.image_bytes;
        umma::mbar_expect_tx(&bar_load, p.image_bytes);
        const uint64_t pol = umma::l2_evict_last_policy();
        for (uint32_t o = 0; o < p.image_bytes; o += 32768) {
            const uint32_t n = p.image_bytes - o < 32768 ? p.image_bytes - o : 32768;
            umma::bulk_g2s(smem_raw + o, src + o, n, &bar_load, pol);
        }
    }
    umma::mbar_wait(&bar_load, 0);
    if (CG == 2) umma::cluster_sync_all(); else __syncthreads();
    long long t0 = 0, t1 = 0;
    if (rank == 0 && threadIdx.x == 0) {
        const uint32_t sbase = umma::smem_u32(smem_raw);
        t0 = clock64();
        for (uint32_t r = 0; r < p.reps; ++r)
            for (uint32_t i = 0; i < p.n_ops; ++i) {
                const Op op = s_ops[i];
                umma::mma_ss<CG, KIND16>(tbase + op.d_col, umma::smem_desc(sbase + op.a_off, p.lbo_a, 128),
                                         umma::smem_desc(sbase + op.b_off, p.lbo_b, 128), p.idesc,
                                         op.acc != 0 || r > 0);
            }
        umma::mma_commit<CG>(&bar_done, 3);
    }
    umma::mbar_wait(&bar_done, 0);
    if (rank == 0 && threadIdx.x == 0) { t1 = clock64(); if (p.cycles) *p.cycles = (unsigned long long)(t1 - t0); }
    umma::tc_fence_after();
    float* out = p.dump + (size_t)rank * 128 * p.ncols;
    for (uint32_t c = 0; c < p.ncols; c += 8) {
        uint32_t v[8];
        umma::tmem_ld8(tbase + ((uint32_t)(32 * warp) << 16) + c, v);
        umma::tmem_ld_wait();
        for (int j = 0; j < 8; ++j) out[(size_t)(32 * warp + lane) * p.ncols + c + j] = __uint_as_float(v[j]);
    }
    if (p.ld16_off8)
        for (uint32_t c = 8; c + 16 <= p.ncols; c += 16) {
            uint32_t v[16];
            umma::tmem_ld16(tbase + ((uint32_t)(32 * warp) << 16) + c, v);
            umma::tmem_ld_wait();
            for (int j = 0; j < 16; ++j) out[(size_t)(32 * warp + lane) * p.ncols + c + j] = __uint_as_float(v[j]);
        }
    umma::tc_fence_before();
    if (CG == 2) umma::cluster_sync_all(); else __syncthreads();
    if (warp == 0) umma::tmem_free<CG>(tbase, 512);
}

// tight issue loop: descriptors advance by a constant, 8 MMAs per unrolled iteration
template <int CG, bool KIND16>
__global__ void __launch_bounds__(128, 1) k_rate(uint32_t idesc, uint32_t lbo_a, uint32_t lbo_b, uint32_t iters,
                                                 unsigned long long* cycles, uint32_t nacc_mask, uint32_t acc_stride, uint32_t uniform) {
    __shared__ __align__(8) uint64_t bar_done;
    __shared__ uint32_t tmem_slot;
    const uint32_t rank = CG == 2 ? umma::cluster_ctarank() : 0;
    const int warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
    if (threadIdx.x == 0) { umma::mbar_init(&bar_done, 1); umma::mbar_fence_init(); }
    if (warp == 0) umma::tmem_alloc<CG>(&tmem_slot, 512);
    umma::fence_proxy_async();
    umma::tc_fence_before();
    if (CG == 2) umma::cluster_sync_all(); else __syncthreads();
    umma::tc_fence_after();
    const uint32_t tbase = tmem_slot;
    if (uniform && rank == 0 && warp == 0) {
        // whole warp in the loop, one elected lane issues: operands stay warp-uniform
        const uint32_t sbase = umma::smem_u32(smem_raw);
        const uint64_t a0 = umma::smem_desc(sbase, lbo_a, 128), b0 = umma::smem_desc(sbase + 16384, lbo_b, 128);
        const long long t0 = clock64();
        for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (umma::elect_one())
                    umma::mma_ss<CG, KIND16>(tbase + (j & nacc_mask) * acc_stride, a0 + (uint64_t)(j & 3) * 2, b0 + (uint64_t)(j & 3) * 2, idesc, true);
        }
        if (umma::elect_one()) umma::mma_commit<CG>(&bar_done, 3);
        __syncwarp();
        umma::mbar_wait(&bar_done, 0);
        if (threadIdx.x == 0) *cycles = (unsigned long long)(clock64() - t0);
    } else if (!uniform && rank == 0 && threadIdx.x == 0) {
        const uint32_t sbase = umma::smem_u32(smem_raw);
        const uint64_t a0 = umma::smem_desc(sbase, lbo_a, 128), b0 = umma::smem_desc(sbase + 16384, lbo_b, 128);
        const long long t0 = clock64();
        for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                umma::mma_ss<CG, KIND16>(tbase + (j & nacc_mask) * acc_stride, a0 + (uint64_t)(j & 3) * 2, b0 + (uint64_t)(j & 3) * 2, idesc, true);
        }
        umma::mma_commit<CG>(&bar_done, 3);
        umma::mbar_wait(&bar_done, 0);
        *cycles = (unsigned long long)(clock64() - t0);
    } else {
        umma::mbar_wait(&bar_done, 0);
    }
    umma::tc_fence_before();
    if (CG == 2) umma::cluster_sync_all(); else __syncthreads();
    if (warp == 0) umma::tmem_free<CG>(tbase, 512);
}

static void test_rate_tight(int CG, bool kind16, int M, int N, int nacc = 2, uint32_t uniform = 0) {
    unsigned long long* d_cyc; CK(cudaMalloc(&d_cyc, 8));
    void* fn = CG == 1 ? (kind16 ? (void*)k_rate<1, true> : (void*)k_rate<1, false>)
                       : (kind16 ? (void*)k_rate<2, true> : (void*)k_rate<2, false>);
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CG); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 65536;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    uint32_t idesc = umma::instr_desc(kind16 ? umma::FMT_F16 : umma::FMT_TF32, M, N), la = 64, lb = 64, iters = 256;
    uint32_t nmask = nacc - 1, astride = 512 / nacc;
    void* args[] = {&idesc, &la, &lb, &iters, &d_cyc, &nmask, &astride, &uniform};
    cudaError_t e = cudaLaunchKernelExC(&cfg, fn, args);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("tight CG=%d M=%d N=%d : %s\n", CG, M, N, cudaGetErrorString(e)); exit(1); }
    unsigned long long cyc; CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
    const double per = (double)cyc / (iters * 8);
    printf("tight%s nacc=%d CG=%d %s M=%3d N=%3d : %6.1f cycles/MMA -> %6.0f FLOP/cycle/SM (floor %d)\n", uniform ? "(uniform)" : "", nacc, CG, kind16 ? "f16 " : "tf32", M, N, per,
           2.0 * M * N * (kind16 ? 16 : 8) / per / CG, M * N / (256 * CG));
    cudaFree(d_cyc);
}

// ---- host helpers -------------------------------------------------------------------------------
// operand plane with R rows in the engine's layout: 16-byte k-chunk c of row r at c*R*16 + r*16
static size_t plane_off(int R, int r, int kbyte) { return (size_t)(kbyte / 16) * R * 16 + (size_t)r * 16 + kbyte % 16; }

struct Run {
    int CG = 1; bool kind16 = true; int M = 128, N = 64;   // M, N: instruction shape (whole pair for CG=2)
    int rowsA = 128, rowsB = 64;                            // rows held per CTA
    std::vector<uint8_t> image[2];
    std::vector<Op> ops;
    int ncols = 64, reps = 1, ld16_off8 = 0;
    std::vector<float> dump;                                // [CG][128][ncols]
    unsigned long long cycles = 0;
};

static void launch(Run& r) {
    const size_t ib = r.image[0].size();
    uint8_t* d_img; Op* d_ops; float* d_dump; unsigned long long* d_cyc;
    CK(cudaMalloc(&d_img, ib * r.CG)); CK(cudaMalloc(&d_ops, r.ops.size() * sizeof(Op)));
    CK(cudaMalloc(&d_dump, sizeof(float) * 128 * r.ncols * r.CG)); CK(cudaMalloc(&d_cyc, 8));
    for (int c = 0; c < r.CG; ++c) CK(cudaMemcpy(d_img + c * ib, r.image[c].data(), ib, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ops, r.ops.data(), r.ops.size() * sizeof(Op), cudaMemcpyHostToDevice));
    CK(cudaMemset(d_dump, 0xff, sizeof(float) * 128 * r.ncols * r.CG));
    Params p{d_img, (uint32_t)ib, d_ops, (uint32_t)r.ops.size(), (uint32_t)r.reps,
             umma::instr_desc(r.kind16 ? umma::FMT_F16 : umma::FMT_TF32, r.M, r.N), (uint32_t)r.rowsA * 16,
             (uint32_t)r.rowsB * 16, d_dump, (uint32_t)r.ncols, d_cyc, (uint32_t)r.ld16_off8};
    void* fn = r.CG == 1 ? (r.kind16 ? (void*)k_run<1, true> : (void*)k_run<1, false>)
                         : (r.kind16 ? (void*)k_run<2, true> : (void*)k_run<2, false>);
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((ib + 1023) / 1024 * 1024)));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(r.CG); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = (ib + 1023) / 1024 * 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = r.CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    void* args[] = {&p};
    CK(cudaLaunchKernelExC(&cfg, fn, args));
    CK(cudaDeviceSynchronize());
    r.dump.resize((size_t)128 * r.ncols * r.CG);
    CK(cudaMemcpy(r.dump.data(), d_dump, r.dump.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&r.cycles, d_cyc, 8, cudaMemcpyDeviceToHost));
    cudaFree(d_img); cudaFree(d_ops); cudaFree(d_dump); cudaFree(d_cyc);
}

static void put16(std::vector<uint8_t>& img, size_t base, int R, int r, int k, float v) {
    const __half h = __float2half_rn(v);
    memcpy(&img[base + plane_off(R, r, 2 * k)], &h, 2);
}
static void put32(std::vector<uint8_t>& img, size_t base, int R, int r, int k, float v) {
    memcpy(&img[base + plane_off(R, r, 4 * k)], &v, 4);
}

// ---- test 1: layout / descriptor validation with small integers (exact) ----------------------------
static int test_layout(int CG, bool kind16, int N, int K, int ld16_off8 = 0) {
    Run r; r.ld16_off8 = ld16_off8; r.CG = CG; r.kind16 = kind16; r.M = 128; r.N = N;
    r.rowsA = 128 / CG; r.rowsB = N / CG; r.ncols = N / CG; r.reps = 1;
    const int es = kind16 ? 2 : 4, kstep = kind16 ? 16 : 8;
    const size_t offA = 0, szA = (size_t)r.rowsA * K * es, offB = szA, szB = (size_t)r.rowsB * K * es;
    auto Af = [](int m, int k) { return (float)((m * 7 + k * 3) % 5 - 2); };
    auto Bf = [](int n, int k) { return (float)((n * 5 + k) % 7 - 3); };
    for (int c = 0; c < CG; ++c) {
        r.image[c].assign(szA + szB, 0);
        for (int m = 0; m < r.rowsA; ++m) for (int k = 0; k < K; ++k)
            (kind16 ? put16 : put32)(r.image[c], offA, r.rowsA, m, k, Af(c * r.rowsA + m, k));
        for (int n = 0; n < r.rowsB; ++n) for (int k = 0; k < K; ++k)
            (kind16 ? put16 : put32)(r.image[c], offB, r.rowsB, n, k, Bf(c * r.rowsB + n, k));
    }
    for (int ks = 0; ks < K / kstep; ++ks)
        r.ops.push_back({(uint32_t)(offA + (size_t)ks * 2 * r.rowsA * 16), (uint32_t)(offB + (size_t)ks * 2 * r.rowsB * 16), 0u, ks > 0 ? 1u : 0u});
    launch(r);
    int bad = 0, shown = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
        double e = 0; for (int k = 0; k < K; ++k) e += (double)Af(m, k) * Bf(n, k);
        int cta, ln, col;
        if (CG == 1) { cta = 0; ln = m; col = n; }
        else { cta = m / 64; ln = m % 64 + 64 * (n >= N / 2); col = n % (N / 2); }
        const float g = r.dump[((size_t)cta * 128 + ln) * r.ncols + col];
        if (g != (float)e) { ++bad; if (shown++ < 4) printf("    mismatch m=%d n=%d expect %g got %g\n", m, n, e, g); }
    }
    printf("layout CG=%d %s N=%d K=%d : %s (%d mismatches of %d)\n", CG, kind16 ? "f16 " : "tf32", N, K, bad ? "FAIL" : "ok", bad, 128 * N);
    if (bad && CG == 2) {   // discover where a few outputs actually landed
        for (int probe = 0; probe < 6; ++probe) {
            const int m = probe * 23 % 128, n = (probe * 37 + 5) % N;
            double e = 0; for (int k = 0; k < K; ++k) e += (double)Af(m, k) * Bf(n, k);
            printf("    D[%d][%d]=%g found at:", m, n, e);
            int cnt = 0;
            for (int c = 0; c < CG; ++c) for (int ln = 0; ln < 128; ++ln) for (int col = 0; col < r.ncols; ++col)
                if (r.dump[((size_t)c * 128 + ln) * r.ncols + col] == (float)e && cnt++ < 6) printf(" (cta%d,l%d,c%d)", c, ln, col);
            printf("\n");
        }
    }
    return bad;
}

// ---- test 2: issue rate ---------------------------------------------------------------------------
static void test_rate(int CG, bool kind16, int M, int N) {
    Run r; r.CG = CG; r.kind16 = kind16; r.M = M; r.N = N;
    r.rowsA = M / CG; r.rowsB = N / CG; r.ncols = 8; r.reps = 64;
    const int K32 = 8;                                      // eight k-steps of 32 bytes
    const size_t szA = (size_t)r.rowsA * K32 * 32, szB = (size_t)r.rowsB * K32 * 32;
    for (int c = 0; c < CG; ++c) r.image[c].assign(szA + szB, 0);
    for (int ks = 0; ks < K32; ++ks)
        r.ops.push_back({(uint32_t)((size_t)ks * 2 * r.rowsA * 16), (uint32_t)(szA + (size_t)ks * 2 * r.rowsB * 16), 0u, 1u});
    launch(r);
    const double per = (double)r.cycles / (r.reps * K32);
    const double flop = 2.0 * M * N * (kind16 ? 16 : 8);
    printf("rate CG=%d %s M=%3d N=%3d : %7.1f cycles/MMA  -> %6.0f FLOP/cycle/SM (floor M*N/(256*CG)=%d)\n", CG,
           kind16 ? "f16 " : "tf32", M, N, per, flop / per / CG, M * N / (256 * CG));
}

// ---- test 3: accumulate rounding ------------------------------------------------------------------
// acc = 1.0, then `n` times acc += x with x = xa*xb; prints the result in ulps of 1.0 (2^-23)
static void test_round(bool kind16, float xa, float xb, int n, const char* what) {
    Run r; r.CG = 1; r.kind16 = kind16; r.M = 128; r.N = 16; r.rowsA = 128; r.rowsB = 16; r.ncols = 16; r.reps = 1;
    const size_t szA = 128 * 32, szB = 16 * 32;     // one k-step (32 bytes of K) per plane
    r.image[0].assign(2 * szA + 2 * szB, 0);
    for (int m = 0; m < 128; ++m) {
        (kind16 ? put16 : put32)(r.image[0], 0, 128, m, 0, 1.0f);
        (kind16 ? put16 : put32)(r.image[0], szA, 128, m, 0, xa);
    }
    for (int q = 0; q < 16; ++q) {
        (kind16 ? put16 : put32)(r.image[0], 2 * szA, 16, q, 0, 1.0f);
        (kind16 ? put16 : put32)(r.image[0], 2 * szA + szB, 16, q, 0, xb);
    }
    r.ops.push_back({0u, (uint32_t)(2 * szA), 0u, 0u});
    for (int i = 0; i < n; ++i) r.ops.push_back({(uint32_t)szA, (uint32_t)(2 * szA + szB), 0u, 1u});
    launch(r);
    const double got = r.dump[0], exact = 1.0 + (double)n * xa * xb;
    printf("round %s %-34s n=%3d : got 1%+.3f ulp, exact 1%+.3f ulp\n", kind16 ? "f16 " : "tf32", what, n,
           (got - 1.0) / ldexp(1.0, -23), (exact - 1.0) / ldexp(1.0, -23));
}

// ---- test 4: accuracy of the three-product splits ---------------------------------------------------
static float tf32_hi(float v) { uint32_t u; memcpy(&u, &v, 4); u &= 0xffffe000u; float h; memcpy(&h, &u, 4); return h; }
static float tf32_rn(float v) { uint32_t u; memcpy(&u, &v, 4); u += 0x1000u; u &= 0xffffe000u; float h; memcpy(&h, &u, 4); return h; }

struct Acc { double bias = 0, rms = 0; };
static Acc score(const std::vector<float>& got, const std::vector<double>& exact) {
    double num = 0, den = 0, sb = 0, sa = 0;
    for (size_t i = 0; i < got.size(); ++i) {
        const double d = got[i] - exact[i];
        num += d * d; den += exact[i] * exact[i];
        sb += d * (exact[i] >= 0 ? 1 : -1); sa += fabs(exact[i]);
    }
    return {sb / sa, sqrt(num / den)};
}

// mode 0: one accumulator, order per k-step lo*hi, hi*lo, hi*hi; mode 1: cross terms in a second
// accumulator; mode 2: as 0 but K split in two halves with their own accumulators
static void test_split(bool kind16, int mode, int data, int K) {
    const int M = 128, N = 32;
    std::mt19937 rng(1234 + data);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<float> A((size_t)M * K), B((size_t)N * K);
    for (auto& v : A) { float g = nd(rng); v = data == 0 ? fmaxf(g, 0.f) : (data == 1 ? g : fabsf(g) + 0.5f); }
    for (auto& v : B) { float g = nd(rng) * 0.056f; v = data == 2 ? fabsf(g) + 0.01f : g; }
    // per-row scale of A for the f16 split (power of two, row max -> [2^13, 2^14)); B: global scale
    std::vector<float> sa(M, 1.f); float sb = 1.f;
    if (kind16) {
        for (int m = 0; m < M; ++m) {
            float mx = 0; for (int k = 0; k < K; ++k) mx = fmaxf(mx, fabsf(A[(size_t)m * K + k]));
            int e; frexpf(mx, &e); sa[m] = mx > 0 ? ldexpf(1.f, 14 - e) : 1.f;
        }
        float mx = 0; for (auto v : B) mx = fmaxf(mx, fabsf(v));
        int e; frexpf(mx, &e); sb = ldexpf(1.f, 14 - e);
    }
    Run r; r.CG = 1; r.kind16 = kind16; r.M = M; r.N = N; r.rowsA = M; r.rowsB = N; r.reps = 1;
    r.ncols = mode == 0 ? N : 2 * N;
    const int es = kind16 ? 2 : 4, kstep = kind16 ? 16 : 8;
    const size_t szA = (size_t)M * K * es, szB = (size_t)N * K * es;
    const size_t oAh = 0, oAl = szA, oBh = 2 * szA, oBl = 2 * szA + szB;
    r.image[0].assign(2 * szA + 2 * szB, 0);
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) {
        const float v = A[(size_t)m * K + k] * sa[m];
        if (kind16) { const float h = __half2float(__float2half_rn(v)); put16(r.image[0], oAh, M, m, k, h); put16(r.image[0], oAl, M, m, k, v - h); }
        else { const float h = tf32_hi(v); put32(r.image[0], oAh, M, m, k, h); put32(r.image[0], oAl, M, m, k, tf32_rn(v - h)); }
    }
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) {
        const float v = B[(size_t)n * K + k] * sb;
        if (kind16) { const float h = __half2float(__float2half_rn(v)); put16(r.image[0], oBh, N, n, k, h); put16(r.image[0], oBl, N, n, k, v - h); }
        else { const float h = tf32_hi(v); put32(r.image[0], oBh, N, n, k, h); put32(r.image[0], oBl, N, n, k, tf32_rn(v - h)); }
    }
    const int nk = K / kstep;
    for (int ks = 0; ks < nk; ++ks) {
        const uint32_t ka = (uint32_t)((size_t)ks * 2 * M * 16), kb = (uint32_t)((size_t)ks * 2 * N * 16);
        const uint32_t half = (mode == 2 && ks >= nk / 2) ? N : 0;
        const bool first = mode == 2 ? (ks == 0 || ks == nk / 2) : ks == 0;
        const uint32_t cx = mode == 1 ? N : half;            // accumulator of the cross terms
        r.ops.push_back({(uint32_t)oAl + ka, (uint32_t)oBh + kb, cx, first ? 0u : 1u});
        r.ops.push_back({(uint32_t)oAh + ka, (uint32_t)oBl + kb, cx, 1u});
        r.ops.push_back({(uint32_t)oAh + ka, (uint32_t)oBh + kb, half, (mode == 1 ? !first : true) ? 1u : 0u});
    }
    launch(r);
    std::vector<double> exact((size_t)M * N);
    std::vector<float> got((size_t)M * N), f32((size_t)M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
        double e = 0; float f = 0.f;
        for (int k = 0; k < K; ++k) { e += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k]; f = fmaf(A[(size_t)m * K + k], B[(size_t)n * K + k], f); }
        exact[(size_t)m * N + n] = e; f32[(size_t)m * N + n] = f;
        float g = r.dump[(size_t)m * r.ncols + n];
        if (mode != 0) g += r.dump[(size_t)m * r.ncols + N + n];
        got[(size_t)m * N + n] = g / (sa[m] * sb);
    }
    const Acc a = score(got, exact), b = score(f32, exact);
    static const char* dn[] = {"relu(g) x g", "g x g", "pos x pos"};
    printf("split %s mode=%d K=%3d %-11s : bias %+.3e rms %.3e   (fp32 fma: bias %+.3e rms %.3e)\n", kind16 ? "f16 " : "tf32",
           mode, K, dn[data], a.bias, a.rms, b.bias, b.rms);
}

int main(int argc, char** argv) {
    const char* grp = argc > 1 ? argv[1] : "all";
    auto on = [&](const char* g) { return !strcmp(grp, "all") || !strcmp(grp, g); };
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("[%s] device %s sm_%d%d, %d SMs\n", grp, prop.name, prop.major, prop.minor, prop.multiProcessorCount);
    if (on("layout1")) {
        test_layout(1, true, 64, 64);
        test_layout(1, false, 64, 32);
        test_layout(1, true, 160, 32);
    }
    if (on("layout2")) {
        test_layout(2, true, 64, 64);
        test_layout(2, false, 160, 32);
        test_layout(2, true, 256, 32);
    }
    if (on("rate1"))
        for (int k16 = 1; k16 >= 0; --k16) {
            test_rate(1, k16, 128, 32); test_rate(1, k16, 128, 64); test_rate(1, k16, 128, 160); test_rate(1, k16, 128, 256);
            test_rate(1, k16, 64, 160); test_rate(1, k16, 64, 256);
        }
    if (on("rate2"))
        for (int k16 = 1; k16 >= 0; --k16) {
            test_rate(2, k16, 128, 32); test_rate(2, k16, 128, 64); test_rate(2, k16, 128, 160); test_rate(2, k16, 128, 256);
            test_rate(2, k16, 256, 160); test_rate(2, k16, 256, 256);
        }
    if (on("tight")) {
        const int ns[] = {16, 32, 64, 96, 128, 160, 192, 256};
        for (int n : ns) test_rate_tight(1, true, 128, n);
        for (int n : ns) if (n >= 32) test_rate_tight(2, true, 128, n);
        test_rate_tight(2, true, 256, 256);
        test_rate_tight(1, false, 128, 160);
        test_rate_tight(2, false, 128, 160);
    }
    if (on("nacc")) {
        for (int na : {1, 2, 4, 8}) { test_rate_tight(2, true, 128, 32, na); test_rate_tight(2, true, 128, 64, na); test_rate_tight(1, true, 128, 32, na); }
        test_rate_tight(2, true, 128, 160, 1); test_rate_tight(2, true, 128, 160, 2); test_rate_tight(2, true, 128, 256, 1);
    }
    if (on("ldalign")) {       // x16 tensor-memory loads at a column that is 8 mod 16: same data?
        test_layout(1, true, 160, 32, 1);
        test_layout(2, true, 256, 32, 1);
    }
    if (on("uniform")) {
        for (int n : {32, 64, 96, 128, 160, 192, 256}) test_rate_tight(2, true, 128, n, 2, 1);
        for (int n : {16, 32, 64, 128, 256}) test_rate_tight(1, true, 128, n, 2, 1);
    }
    if (on("tightodd")) { test_rate_tight(2, true, 128, 176); test_rate_tight(2, true, 128, 48); test_rate_tight(2, true, 128, 16); }
    if (on("round")) {
        test_round(true, 1.5f * ldexpf(1.f, -12), ldexpf(1.f, -12), 8, "x=+0.75ulp (RN:+1/step, RZ:0)");
        test_round(false, 1.5f * ldexpf(1.f, -12), ldexpf(1.f, -12), 8, "x=+0.75ulp (RN:+1/step, RZ:0)");
        test_round(true, -ldexpf(1.f, -12), ldexpf(1.f, -13), 8, "x=-0.25ulp (RN:0, RZ:-0.5/step)");
        test_round(false, -ldexpf(1.f, -12), ldexpf(1.f, -13), 8, "x=-0.25ulp (RN:0, RZ:-0.5/step)");
        test_round(true, ldexpf(1.f, -12), ldexpf(1.f, -12), 8, "x=+0.5ulp (tie)");
        test_round(true, 1.25f * ldexpf(1.f, -11), ldexpf(1.f, -12), 8, "x=+1.25ulp");
        test_round(false, 1.25f * ldexpf(1.f, -11), ldexpf(1.f, -12), 8, "x=+1.25ulp");
        test_round(true, 1.5f * ldexpf(1.f, -12), ldexpf(1.f, -12), 1, "x=+0.75ulp single");
        test_round(true, -1.5f * ldexpf(1.f, -12), ldexpf(1.f, -12), 1, "x=-0.75ulp single");
    }
    if (on("split"))
        for (int data = 0; data < 3; ++data) {
            for (int mode = 0; mode < 3; ++mode) test_split(true, mode, data, 320);
            for (int mode = 0; mode < 3; ++mode) test_split(false, mode, data, 160);
            test_split(true, 0, data, 160);
        }
    return 0;
}
