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
//      fp32 FMA tiles (64 x 64 per CTA, 4 x 4 per thread), the batch split over blockIdx.y into
//      fixed slices whose partial tiles are added in slice order by k_wgrad_reduce: deterministic,
//      no atomics.  The remaining chain rule (dW1 = dM1^T Wmix[:, :d1], dWmix += dM1 W1, LU
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
#define WG_KC 16

// grid: x = tiles (M tiles x N tiles), y = batch slices, z = layers
__global__ void __launch_bounds__(256)
k_wgrad(FabTape tape, int offA, int M, int offB, int N, const float* __restrict__ g, int rows_per_split,
        float* __restrict__ part) {
    __shared__ float As[WG_KC][WG_BM + 4], Bs[WG_KC][WG_BN + 4];
    const int tilesN = (N + WG_BN - 1) / WG_BN;
    const int m0 = (blockIdx.x / tilesN) * WG_BM, n0 = (blockIdx.x % tilesN) * WG_BN;
    const float* T = tape.layer(blockIdx.z);
    const long long i0 = (long long)blockIdx.y * rows_per_split;
    const long long i1 = min(tape.n, i0 + rows_per_split);
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
    for (long long ic = i0; ic < i1; ic += WG_KC) {
#pragma unroll
        for (int r = 0; r < (WG_KC * WG_BM) / 256; ++r) {
            const int idx = r * 256 + threadIdx.x, kk = idx / WG_BM, mm = idx % WG_BM;
            const long long i = ic + kk;
            const bool ok = i < i1;
            const float* row = T + (size_t)i * tape.RS;
            As[kk][mm] = (ok && m0 + mm < M) ? __ldg(g + i) * __ldg(row + offA + m0 + mm) : 0.f;
            Bs[kk][mm] = (ok && n0 + mm < N) ? __ldg(row + offB + n0 + mm) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < WG_KC; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][4 * ty]);
            const float4 c = *reinterpret_cast<const float4*>(&Bs[kk][4 * tx]);
            const float av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(av[u], cv[v], acc[u][v]);
        }
        __syncthreads();
    }
    float* P = part + ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * M * N;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int m = m0 + 4 * ty + u;
        if (m >= M) continue;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int nn = n0 + 4 * tx + v;
            if (nn < N) P[(size_t)m * N + nn] = acc[u][v];
        }
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
