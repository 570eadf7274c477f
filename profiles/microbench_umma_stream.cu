// What bounds the tensor rate of the row-tile engine's wide GEMM (cta_group::2, M 128 x N 160 x K 16,
// three MMAs per k-step: cross = lo*hi, cross += hi*lo, main = hi*hi)?  One CTA pair, operands in the
// engine's layouts (A: hi / lo planes of 64 rows x 336 k, k-step stride 2 KB; B: ring of 4 stages x
// 4 k-steps x 5 KB), zero data, no epilogue, no TMA traffic.  Modes:
//   same    : every MMA reads the same A / B addresses (the "tight" case of microbench_umma.cu)
//   stream  : addresses advance like the engine's k loop (A over 21 k-steps, B over the ring)
//   streamA : only A advances          streamB : only B advances
//   two     : stream, with TWO issuing warps (warp 0: main, warp 2: the two cross products)
//   n80/n256: stream with instruction N = 80 / 256 (smaller / larger B operand per MMA)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fab_torch_b200/csrc -o profiles/mb_umma_stream profiles/microbench_umma_stream.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "umma.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

extern __shared__ __align__(1024) uint8_t smem_raw[];

constexpr uint32_t A_PLANE = 64 * 336 * 2, RING = 4 * 4 * 8192, A_OFF = 0, B_OFF = 2 * A_PLANE;

// mode bits: 1 = A advances, 2 = B advances, 4 = two issuers
__global__ void __launch_bounds__(128, 1) k_stream(uint32_t idesc, uint32_t nrows_b, uint32_t mode, uint32_t iters,
                                                   unsigned long long* cycles) {
    __shared__ __align__(8) uint64_t bar_done;
    __shared__ uint32_t tmem_slot;
    const uint32_t rank = umma::cluster_ctarank();
    const int warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < (B_OFF + RING) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
    if (threadIdx.x == 0) { umma::mbar_init(&bar_done, (mode & 4) ? 2 : 1); umma::mbar_fence_init(); }
    if (warp == 0) umma::tmem_alloc<2>(&tmem_slot, 512);
    umma::fence_proxy_async();
    umma::tc_fence_before();
    umma::cluster_sync_all();
    umma::tc_fence_after();
    const uint32_t tbase = tmem_slot;
    const bool two = (mode & 4) != 0;
    if (rank == 0 && (warp == 0 || (two && warp == 2))) {
        const uint32_t sbase4 = umma::smem_u32(smem_raw) >> 4;
        const uint32_t dhi = (128u >> 4) | (1u << 14);
        auto desc = [&](uint32_t lo) { return ((uint64_t)dhi << 32) | lo; };
        const uint32_t lead = umma::elect_one() ? 1u : 0u;
        const uint32_t a_base = sbase4 + (A_OFF >> 4) + ((1024u >> 4) << 16), al = A_PLANE >> 4;
        const uint32_t b_base = sbase4 + (B_OFF >> 4) + (((nrows_b * 16u) >> 4) << 16), kstr = nrows_b * 2, nbh = nrows_b / 2;
        const uint32_t d0 = tbase, d1 = tbase + 128;
        const long long t0 = clock64();
        for (uint32_t it = 0; it < iters; ++it) {
            // one block of a W x W GEMM: 20 k-steps in 5 stages of 4
            for (uint32_t st = 0; st < 5; ++st) {
                uint32_t ah = a_base + ((mode & 1) ? st * 4 * 128 : 0);
                uint32_t bk = b_base + ((mode & 2) ? (st & 3) * (4 * kstr) : 0);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (!two || warp == 2) {
                        umma::mma_ss2_pred(d1, desc(ah + al), desc(bk), idesc, 1u, lead);
                        umma::mma_ss2_pred(d1, desc(ah), desc(bk + nbh), idesc, 1u, lead);
                    }
                    if (!two || warp == 0) umma::mma_ss2_pred(d0, desc(ah), desc(bk), idesc, 1u, lead);
                    if (mode & 1) ah += 128;
                    if (mode & 2) bk += kstr;
                }
            }
        }
        if (lead) umma::mma_commit<2>(&bar_done, 3);
        __syncwarp();
        umma::mbar_wait(&bar_done, 0);
        if (threadIdx.x == 0) *cycles = (unsigned long long)(clock64() - t0);
    } else {
        umma::mbar_wait(&bar_done, 0);
    }
    umma::tc_fence_before();
    umma::cluster_sync_all();
    if (warp == 0) umma::tmem_free<2>(tbase, 512);
}

// queue depth: cycles until the issuing thread gets past K back-to-back MMAs (N = 160) on an idle pipe
__global__ void __launch_bounds__(128, 1) k_queue(uint32_t idesc, unsigned long long* out /* [33] */) {
    __shared__ __align__(8) uint64_t bar_done;
    __shared__ uint32_t tmem_slot;
    const uint32_t rank = umma::cluster_ctarank();
    const int warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < (B_OFF + RING) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
    if (threadIdx.x == 0) { umma::mbar_init(&bar_done, 1); umma::mbar_fence_init(); }
    if (warp == 0) umma::tmem_alloc<2>(&tmem_slot, 512);
    umma::fence_proxy_async();
    umma::tc_fence_before();
    umma::cluster_sync_all();
    umma::tc_fence_after();
    const uint32_t tbase = tmem_slot;
    if (rank == 0 && warp == 0) {
        const uint32_t sbase4 = umma::smem_u32(smem_raw) >> 4;
        const uint32_t dhi = (128u >> 4) | (1u << 14);
        auto desc = [&](uint32_t lo) { return ((uint64_t)dhi << 32) | lo; };
        const uint32_t lead = umma::elect_one() ? 1u : 0u;
        const uint32_t a = sbase4 + ((1024u >> 4) << 16), b = sbase4 + (B_OFF >> 4) + ((2560u >> 4) << 16);
        uint32_t par = 0;
        for (int K = 1; K <= 32; ++K) {
            const long long t0 = clock64();
            for (int j = 0; j < K; ++j) umma::mma_ss2_pred(tbase, desc(a), desc(b), idesc, 1u, lead);
            const long long t1 = clock64();
            if (lead) umma::mma_commit<2>(&bar_done, 1);
            __syncwarp();
            umma::mbar_wait(&bar_done, par);
            par ^= 1;
            const long long t2 = clock64();
            if (threadIdx.x == 0) out[K] = ((unsigned long long)(t1 - t0) << 32) | (unsigned long long)(t2 - t0);
        }
    }
    umma::tc_fence_before();
    umma::cluster_sync_all();
    if (warp == 0) umma::tmem_free<2>(tbase, 512);
}

static void run_queue() {
    unsigned long long* d; CK(cudaMalloc(&d, 8 * 33));
    const int smem = B_OFF + RING;
    CK(cudaFuncSetAttribute(k_queue, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    uint32_t idesc = umma::instr_desc(umma::FMT_F16, 128, 160);
    void* args[] = {&idesc, &d};
    for (int rep = 0; rep < 2; ++rep) { CK(cudaLaunchKernelExC(&cfg, (const void*)k_queue, args)); CK(cudaDeviceSynchronize()); }
    unsigned long long h[33]; CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
    printf("queue: K back-to-back MMAs (N = 160) on an idle pipe: cycles until the issuer is past them / until they are complete\n");
    for (int K = 1; K <= 32; ++K) printf("  K=%2d  issued after %5llu   complete after %5llu\n", K, h[K] >> 32, h[K] & 0xffffffffull);
    cudaFree(d);
}

static void run(const char* name, int N, uint32_t mode) {
    unsigned long long* d_cyc; CK(cudaMalloc(&d_cyc, 8));
    const int smem = B_OFF + RING;
    CK(cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    uint32_t idesc = umma::instr_desc(umma::FMT_F16, 128, N), nrows = N, iters = 200;
    void* args[] = {&idesc, &nrows, &mode, &iters, &d_cyc};
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaLaunchKernelExC(&cfg, (const void*)k_stream, args));
        CK(cudaDeviceSynchronize());
    }
    unsigned long long cyc; CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
    const double per = (double)cyc / (iters * 60.0);
    printf("%-8s N=%3d : %6.1f cycles/MMA   (math floor %d, smem operand bytes per MMA and SM: A 2048 + B %d)\n", name, N, per,
           N / 4, N / 2 * 32);
    cudaFree(d_cyc);
}

int main() {
    run_queue();
    run("same", 160, 0); run("streamA", 160, 1); run("streamB", 160, 2); run("stream", 160, 3); run("two", 160, 7);
    run("two-same", 160, 4);
    run("stream", 80, 3); run("stream", 128, 3); run("stream", 256, 3); run("two", 256, 7); run("two", 80, 7);
    return 0;
}
