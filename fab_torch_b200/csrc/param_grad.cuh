// Parameter gradient of sum_i g_i log q_theta(x_i) for the RealNVP flow (the
// theta-gradient of the FAB loss, fab/core.py:112-118, minibatch loop
// fab/train_with_prioritised_buffer.py:158-186).
//
// Two steps:
//   1. k_flow_tape (tile engine, flow_tile.cuh with TAPE): log q, d log q / dx and the activation
//      tape -- per layer and particle the MLP activations and the UNIT-SEED gradients of the reverse
//      sweep (the sweep is linear in its seed, so the gradients of the weighted loss are g_i times
//      the unit-seed ones).
//   2. k_wgrad: the weight gradients as GEMMs with the BATCH as the contraction dimension,
//         C[m][n] = sum_i g_i A[i][m] B[i][n],        A, B = column ranges of the tape rows,
//      four per layer (the ones columns of the tape turn bias gradients into extra rows/columns):
//         Ga = [z_in | 1]^T gh1    (d+1) x W     rows 0..d-1: dM1 = d/d(Wmix[:, :d1] W1^T), row d: db1
//         Gb = [z_in | 1]^T gv     (d+1) x d     direct part of dWmix; row d: gradient of the bias of v
//         Gc =  gh2^T      [h1 | 1]  W x (W+1)   dW2 (torch layout [out][in]) | db2
//         Gd =  gparam^T   [h2 | 1]  2d2 x (W+1) dW3 (rows: shifts, then scales) | db3
//      64 x 64 tiles per CTA on the warp-level tensor path (mma.sync.m16n8k8, 3xTF32 = fp32-grade),
//      the batch split over blockIdx.y into fixed slices whose partial tiles are added in slice
//      order by k_wgrad_reduce: deterministic, no atomics.  The remaining chain rule (dW1 = dM1^T Wmix[:, :d1], dWmix += dM1 W1, LU
//      parameters of Wmix) acts on [W x d]-sized matrices in parameter space and is done by the
//      host (fab_torch_b200/flow.py).
//   k_base_grad: d/d loc, d/d log_scale of the base Gaussian and sum_i g_i (= d/d sum(log_S_k)).
#pragma once
#include "flow_tile.cuh"

template <int TP>
__global__ void __launch_bounds__(FAB_NT, FAB_MIN_CTAS)
k_flow_tape(TileLayout L, fab_flow_desc f, const float* __restrict__ blob, const float* __restrict__ x,
            float* __restrict__ log_q, float* __restrict__ grad, FabTape tape, long long n) {
    float* const smem = fab_smem;
    const TileBufs b = tile_bufs(L);
    float* lq = smem + L.o_state;
    const long long row0 = (long long)blockIdx.x * L.T;
    const int np = (int)min((long long)L.T, n - row0);
    tile_init(L, f, blob);
    load_rows_act<TP>(zsel(L, 0), x, L.d, row0, np);
    __syncthreads();
    flow_inverse<TP, true, true>(L, f, blob, 0, lq, tape, row0, np);
    flow_backward<TP, true>(L, f, blob, tape, row0, np);
    if (grad) store_rows_act<TP>(grad, L.d, b.gs, row0, np);
    for (int p = threadIdx.x; p < np; p += FAB_NT) log_q[row0 + p] = lq[p];
}

#define WG_BM 64
#define WG_BN 64
#define WG_KC 32
#define WG_LD 72          // smem row stride: fragment loads (4 k rows x 8 columns) hit 32 distinct banks

// grid: x = tiles (M tiles x N tiles), y = batch slices, z = layers.  256 threads = 8 warps in a 2 x 4
// grid, each warp a 32 x 16 piece of the 64 x 64 tile = 2 x 2 mma.sync.m16n8k8 tiles; the batch index is
// the k dimension, so the A operand (g_i x tape column m) is read k-major from shared memory.  3xTF32:
// both operands split into tf32 hi + lo, products lo*hi, hi*lo, hi*hi into the fp32 accumulators (the
// same split as the flow GEMMs, mma_gemm.cuh; in registers -- splitting once at the store into shared
// memory doubles the fragment loads and was slower, 206 vs 168 us for the W x W gradient).  The next chunk of tape rows is fetched into registers
// while the current one is multiplied.
__global__ void __launch_bounds__(256)
k_wgrad(FabTape tape, int offA, int M, int offB, int N, const float* __restrict__ g, int rows_per_split,
        float* __restrict__ part) {
    __shared__ float As[WG_KC][WG_LD], Bs[WG_KC][WG_LD];
    const int tilesN = (N + WG_BN - 1) / WG_BN;
    const int m0 = (blockIdx.x / tilesN) * WG_BM, n0 = (blockIdx.x % tilesN) * WG_BN;
    const float* T = tape.layer(blockIdx.z);
    const long long i0 = (long long)blockIdx.y * rows_per_split;
    const long long i1 = min(tape.n, i0 + rows_per_split);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
    const int wm = (warp >> 2) * 32, wn = (warp & 3) * 16;
    constexpr int NLD = (WG_KC * WG_BM) / 256;            // elements of each operand per thread and chunk
    const int mm = threadIdx.x & 63, k0 = threadIdx.x >> 6;   // element r of the thread: row k0 + 4 r, column mm
    const bool okA = m0 + mm < M, okB = n0 + mm < N;
    float ra[NLD], rb[NLD];
    auto fetch = [&](long long ic) {
#pragma unroll
        for (int r = 0; r < NLD; ++r) {
            const long long i = ic + k0 + 4 * r;
            const bool ok = i < i1;
            const float* row = T + (size_t)i * tape.RS;
            ra[r] = (ok && okA) ? __ldg(g + i) * __ldg(row + offA + m0 + mm) : 0.f;
            rb[r] = (ok && okB) ? __ldg(row + offB + n0 + mm) : 0.f;
        }
    };
    float acc[2][2][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[a][c][e] = 0.f;
    fetch(i0);
    for (long long ic = i0; ic < i1; ic += WG_KC) {
#pragma unroll
        for (int r = 0; r < NLD; ++r) { As[k0 + 4 * r][mm] = ra[r]; Bs[k0 + 4 * r][mm] = rb[r]; }
        __syncthreads();
        if (ic + WG_KC < i1) fetch(ic + WG_KC);
        // the tensor core truncates every accumulation (mma_gemm.cuh): a chunk's 12 MMAs go into a fresh
        // accumulator that is added to the running sum with a round-to-nearest FADD
        float cp[2][2][4];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int e = 0; e < 4; ++e) cp[a][c][e] = 0.f;
#pragma unroll
        for (int kb = 0; kb < WG_KC; kb += 8) {
            uint32_t ah[2][4], al[2][4], bh[2][2], bl[2][2];
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const float* ap = &As[kb + t][wm + 16 * a + gq];
                split_tf32(ap[0], ah[a][0], al[a][0]);
                split_tf32(ap[8], ah[a][1], al[a][1]);
                split_tf32(ap[4 * WG_LD], ah[a][2], al[a][2]);
                split_tf32(ap[4 * WG_LD + 8], ah[a][3], al[a][3]);
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const float* bp = &Bs[kb + t][wn + 8 * c + gq];
                split_tf32(bp[0], bh[c][0], bl[c][0]);
                split_tf32(bp[4 * WG_LD], bh[c][1], bl[c][1]);
            }
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    mma_tf32(cp[a][c], al[a], bh[c][0], bh[c][1]);
                    mma_tf32(cp[a][c], ah[a], bl[c][0], bl[c][1]);
                    mma_tf32(cp[a][c], ah[a], bh[c][0], bh[c][1]);
                }
        }
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[a][c][e] += cp[a][c][e];
        __syncthreads();
    }
    float* P = part + ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * M * N;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int m = m0 + wm + 16 * a + gq + (e >> 1) * 8, nn = n0 + wn + 8 * c + 2 * t + (e & 1);
                if (m < M && nn < N) P[(size_t)m * N + nn] = acc[a][c][e];
            }
}

// out[layer * out_stride + out_off + e] = sum over the batch slices (in slice order) of part[layer][s][e]
__global__ void k_wgrad_reduce(const float* __restrict__ part, int splits, int MN, float* __restrict__ out,
                               long long out_stride, long long out_off) {
    const int layer = blockIdx.y;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < MN; e += gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int sp = 0; sp < splits; ++sp) s += part[((size_t)layer * splits + sp) * MN + e];
        out[(size_t)layer * out_stride + out_off + e] = s;
    }
}

// block j < d: d/d loc_j = sum_i g_i u_ij / scale_j,  d/d log_scale_j = sum_i g_i (u_ij^2 - 1);
// block d: sum_i g_i.   tail = [dloc (d) | dlog_scale (d) | sum g]
__global__ void __launch_bounds__(256)
k_base_grad(const float* __restrict__ zf, int DP, int d, const float* __restrict__ loc, const float* __restrict__ lsc,
            const float* __restrict__ g, long long n, float* __restrict__ tail) {
    __shared__ float r1[256], r2[256];
    const int j = blockIdx.x;
    float s1 = 0.f, s2 = 0.f;
    if (j < d) {
        const float inv = expf(-__ldg(lsc + j)), lo = __ldg(loc + j);
        for (long long i = threadIdx.x; i < n; i += 256) {
            const float u = (zf[(size_t)i * DP + j] - lo) * inv, gi = g[i];
            s1 += gi * u * inv;
            s2 += gi * (u * u - 1.0f);
        }
    } else {
        for (long long i = threadIdx.x; i < n; i += 256) s1 += g[i];
    }
    r1[threadIdx.x] = s1; r2[threadIdx.x] = s2;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { r1[threadIdx.x] += r1[threadIdx.x + o]; r2[threadIdx.x] += r2[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (j < d) { tail[j] = r1[0]; tail[d + j] = r2[0]; }
        else tail[2 * d] = r1[0];
    }
}
