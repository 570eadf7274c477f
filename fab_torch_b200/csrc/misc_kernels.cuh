// Small HBM-bound kernels around the tile kernels: standalone target evaluation, log-weight
// update, NaN/inf filter (stable compaction), ESS / log Z reduction, systematic resampling.
#pragma once
#include "target_tile.cuh"

// ------------------------------------------------------------------ K5/K15 standalone target
__global__ void k_target(fab_target_desc t, const float* __restrict__ x, float* __restrict__ lp,
                         float* __restrict__ g, long long n) {
    const int lane = threadIdx.x & 31;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    const float* xr = x + row * t.dim;
    float* gr = g ? g + row * t.dim : nullptr;
    float v = (t.kind == FAB_TARGET_MANYWELL) ? manywell_row(t, xr, gr, t.dim, lane)
            : (t.kind == FAB_TARGET_ALDP_SURROGATE) ? aldp_row(t, xr, gr, t.dim, lane)
                                                    : gmm_row(t, xr, gr, t.dim, lane);
    if (lane == 0) lp[row] = v;
}

// Streaming form of the many-well target for d = 32*C (C = 1, 2, 4): 8*C threads per row, one
// float4 each, U rows per thread in flight (64 B/thread: enough outstanding bytes to fill HBM;
// the warp-per-row form above keeps 4 B/thread in flight and stops at ~1.4 TB/s).  The per-row
// sum is taken in EXACTLY the order of manywell_row + warp_sum -- "lane" l = (4*tr + c) % 32 first
// accumulates its C chunks in sequence, then the xor-16/8/4/2/1 butterfly -- so both kernels return
// the same bits (tests/test_gpu_misc.py::test_streaming_kernels_match_scalar_forms).
template <int C, int U>
__global__ void __launch_bounds__(256)
k_target_manywell_v4(fab_target_desc t, const float4* __restrict__ x, float* __restrict__ lp,
                     float4* __restrict__ g, long long n) {
    constexpr int TPR = 8 * C;          // threads (float4s) per row
    constexpr int RPW = 32 / TPR;       // rows per warp and load instruction
    const int lane = threadIdx.x & 31;
    const int tr = lane % TPR, grp = lane - tr;
    const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long row0 = warp_id * (RPW * U) + lane / TPR;
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long long row = row0 + u * RPW;
        v[u] = row < n ? __ldcs(x + row * TPR + tr) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long long row = row0 + u * RPW;
        const float xv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
        float e[4], gj[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) manywell_elem(t, xv[c], c, e[c], gj[c]);   // parity of 4*tr+c = parity of c
        if (g && row < n) __stcs(g + row * TPR + tr, make_float4(gj[0], gj[1], gj[2], gj[3]));
        float a[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            a[c] = 0.f;
#pragma unroll
            for (int k = 0; k < C; ++k)         // chunk k of lane l lives in thread (tr % 8) + 8k
                a[c] += __shfl_sync(FAB_FULL, e[c], grp + (tr & 7) + 8 * k);
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1)         // lane xor 16, 8, 4  ==  thread xor 4, 2, 1
#pragma unroll
            for (int c = 0; c < 4; ++c) a[c] += __shfl_xor_sync(FAB_FULL, a[c], o);
        const float b0 = a[0] + a[2], b1 = a[1] + a[3];     // lane xor 2
        const float s = b0 + b1;                            // lane xor 1
        if (tr == 0 && row < n) lp[row] = -s - t.log_norm;
    }
}

// ------------------------------------------------------------------ K11 log-weight update
__global__ void k_logw_update(fab_gamma g, fab_gamma gn, const float* __restrict__ lq,
                              const float* __restrict__ lp, float* __restrict__ lw, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float inc = __fsub_rn(gamma_of(gn, lq[i], lp[i]), gamma_of(g, lq[i], lp[i]));
    lw[i] = __fadd_rn(lw[i], inc);
}
// 16-byte form (all three pointers 16-byte aligned): thread i updates elements 4i..4i+3; the
// n % 4 tail goes to the threads behind the last full float4.
__global__ void __launch_bounds__(256)
k_logw_update_v4(fab_gamma g, fab_gamma gn, const float* __restrict__ lq,
                 const float* __restrict__ lp, float* __restrict__ lw, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n4 = n >> 2;
    if (i < n4) {
        const float4 q = __ldcs((const float4*)lq + i), p = __ldcs((const float4*)lp + i);
        float4 w = __ldcs((const float4*)lw + i);
        w.x = __fadd_rn(w.x, __fsub_rn(gamma_of(gn, q.x, p.x), gamma_of(g, q.x, p.x)));
        w.y = __fadd_rn(w.y, __fsub_rn(gamma_of(gn, q.y, p.y), gamma_of(g, q.y, p.y)));
        w.z = __fadd_rn(w.z, __fsub_rn(gamma_of(gn, q.z, p.z), gamma_of(g, q.z, p.z)));
        w.w = __fadd_rn(w.w, __fsub_rn(gamma_of(gn, q.w, p.w), gamma_of(g, q.w, p.w)));
        __stcs((float4*)lw + i, w);
    } else {
        const long long j = 4 * n4 + (i - n4);
        if (j < n) {
            const float inc = __fsub_rn(gamma_of(gn, lq[j], lp[j]), gamma_of(g, lq[j], lp[j]));
            lw[j] = __fadd_rn(lw[j], inc);
        }
    }
}

// ------------------------------------------------------------------ K12 NaN/inf filter
// ws layout: int32 pos[n] | int32 meta[4] (meta[0]=n_in, meta[1]=n_out) | float staging[...]
#define FAB_SCAN_NT 1024
__global__ void __launch_bounds__(FAB_SCAN_NT)
k_filter_scan(const float* __restrict__ lq, const float* __restrict__ lp, long long n,
              const int* __restrict__ n_in_p, int* __restrict__ n_out_p, int* __restrict__ pos,
              int* __restrict__ meta) {
    __shared__ int warp_tot[FAB_SCAN_NT / 32];
    __shared__ int carry, chunk_total;
    const int n_in = n_in_p ? *n_in_p : (int)n;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n_in; base += FAB_SCAN_NT) {
        const int i = base + threadIdx.x;
        int v = 0;
        if (i < n_in) v = (fab_isfinite(lq[i]) && fab_isfinite(lp[i])) ? 1 : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FAB_FULL, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_tot[lane];
            int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FAB_FULL, wi, o);
                if (lane >= o) wi += t;
            }
            warp_tot[lane] = wi - w;        // exclusive prefix of warp totals
            if (lane == 31) chunk_total = wi;
        }
        __syncthreads();
        const int excl = carry + warp_tot[warp] + incl - v;
        if (i < n_in) pos[i] = v ? excl : -1;
        __syncthreads();
        if (threadIdx.x == 0) carry += chunk_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) { meta[0] = n_in; meta[1] = carry; *n_out_p = carry; }
}

// rows: x, gq, gp are [n,d]; lq, lp, lw are [n].  Stage valid rows at their compact position.
__global__ void k_filter_stage(fab_point pt, const float* __restrict__ lw, int d,
                               const int* __restrict__ pos, const int* __restrict__ meta,
                               float* __restrict__ stage, long long n) {
    const int n_in = meta[0], n_out = meta[1];
    if (n_out == n_in) return;
    const int rowf = 3 * d + 3;
    const int lane = threadIdx.x & 31;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n_in) return;
    const int dst = pos[row];
    if (dst < 0) return;
    float* o = stage + (size_t)dst * rowf;
    for (int j = lane; j < d; j += 32) {
        o[j] = pt.d_x[row * d + j];
        o[d + j] = pt.d_grad_log_q ? pt.d_grad_log_q[row * d + j] : 0.f;
        o[2 * d + j] = pt.d_grad_log_p ? pt.d_grad_log_p[row * d + j] : 0.f;
    }
    if (lane == 0) { o[3 * d] = pt.d_log_q[row]; o[3 * d + 1] = pt.d_log_p[row]; o[3 * d + 2] = lw[row]; }
}
__global__ void k_filter_unstage(fab_point pt, float* __restrict__ lw, int d,
                                 const int* __restrict__ meta, const float* __restrict__ stage,
                                 long long n) {
    const int n_in = meta[0], n_out = meta[1];
    if (n_out == n_in) return;
    const int rowf = 3 * d + 3;
    const int lane = threadIdx.x & 31;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n_out) return;
    const float* o = stage + (size_t)row * rowf;
    for (int j = lane; j < d; j += 32) {
        pt.d_x[row * d + j] = o[j];
        if (pt.d_grad_log_q) pt.d_grad_log_q[row * d + j] = o[d + j];
        if (pt.d_grad_log_p) pt.d_grad_log_p[row * d + j] = o[2 * d + j];
    }
    if (lane == 0) { pt.d_log_q[row] = o[3 * d]; pt.d_log_p[row] = o[3 * d + 1]; lw[row] = o[3 * d + 2]; }
}

// ------------------------------------------------------------------ K13 ESS / log Z
__device__ __forceinline__ float block_reduce(float v, bool is_max, float* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = is_max ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        float w = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : (is_max ? -CUDART_INF_F : 0.f);
        w = is_max ? warp_max(w) : warp_sum(w);
        if (lane == 0) sh[32] = w;
    }
    __syncthreads();
    return sh[32];
}

// partial4 = (max, sum e^(v-max), sum e^(2(v-max)), count), v = lw[i] - (sub ? sub[i] : 0)
__global__ void __launch_bounds__(FAB_SCAN_NT)
k_ess_partial(const float* __restrict__ lw, const float* __restrict__ sub, long long n,
              const int* __restrict__ n_active, float* __restrict__ out4) {
    __shared__ float sh[33];
    const long long m_n = n_active ? (long long)*n_active : n;
    float mx = -CUDART_INF_F, has_nan = 0.f;
    for (long long i = threadIdx.x; i < m_n; i += FAB_SCAN_NT) {
        const float v = sub ? __fsub_rn(lw[i], sub[i]) : lw[i];
        if (v != v) has_nan = 1.f;
        mx = fmaxf(mx, v);
    }
    mx = block_reduce(mx, true, sh);
    has_nan = block_reduce(has_nan, false, sh);
    float s1 = 0.f, s2 = 0.f;
    for (long long i = threadIdx.x; i < m_n; i += FAB_SCAN_NT) {
        const float v = sub ? __fsub_rn(lw[i], sub[i]) : lw[i];
        const float e = (mx == -CUDART_INF_F) ? 0.f : expf(v - mx);
        s1 += e; s2 += e * e;
    }
    s1 = block_reduce(s1, false, sh);
    s2 = block_reduce(s2, false, sh);
    if (threadIdx.x == 0) {
        const float nanv = __int_as_float(0x7fc00000);
        out4[0] = has_nan > 0.f ? nanv : mx;
        out4[1] = s1; out4[2] = s2; out4[3] = (float)m_n;
    }
}

// out3 = (ESS, logsumexp, count) from n_parts quadruples (one per rank)
__global__ void k_ess_finalize(const float* __restrict__ parts, int n_parts, float* __restrict__ out3) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float M = -CUDART_INF_F; bool nan = false;
    for (int r = 0; r < n_parts; ++r) {
        const float m = parts[4 * r];
        if (m != m) nan = true;
        M = fmaxf(M, m);
    }
    float S1 = 0.f, S2 = 0.f, N = 0.f;
    for (int r = 0; r < n_parts; ++r) {
        const float m = parts[4 * r];
        if (parts[4 * r + 3] > 0.f && m > -CUDART_INF_F) {
            const float sc = expf(m - M);
            S1 += parts[4 * r + 1] * sc;
            S2 += parts[4 * r + 2] * sc * sc;
        }
        N += parts[4 * r + 3];
    }
    const float nanv = __int_as_float(0x7fc00000);
    out3[0] = nan ? nanv : (S1 * S1) / (S2 * N);      // 1 / (N * sum softmax^2)
    out3[1] = nan ? nanv : M + logf(S1);
    out3[2] = N;
}

// ------------------------------------------------------------------ R systematic resampling
// Deterministic double exp for t <= 0: only IEEE add/mul/rint/ldexp, mirrored op by op in
// oracle/resample.py:exp_det.
__device__ __forceinline__ double exp_det(double t) {
    const double LOG2E = 0x1.71547652b82fep+0, LN2_HI = 0x1.62e42fee00000p-1,
                 LN2_LO = 0x1.a39ef35793c76p-33;
    const double C[14] = {1.0, 1.0, 0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040,
                          1.0 / 40320, 1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800,
                          1.0 / 479001600, 1.0 / 6227020800};
    if (t < -60.0) t = -60.0;
    const double k = rint(__dmul_rn(t, LOG2E));
    const double r = __dsub_rn(__dsub_rn(t, __dmul_rn(k, LN2_HI)), __dmul_rn(k, LN2_LO));
    double p = C[13];
#pragma unroll
    for (int i = 12; i >= 0; --i) p = __dadd_rn(__dmul_rn(p, r), C[i]);
    return ldexp(p, (int)k);
}

// ws: uint64 cdf[n] | float meta[2]
__global__ void __launch_bounds__(FAB_SCAN_NT)
k_resample_cdf(const float* __restrict__ lw, long long n, unsigned long long* __restrict__ cdf) {
    __shared__ float sh[33];
    __shared__ unsigned long long wtot[32];
    __shared__ unsigned long long carry, chunk_total;
    float mx = -CUDART_INF_F;
    for (long long i = threadIdx.x; i < n; i += FAB_SCAN_NT) {
        const float v = lw[i];
        if (fab_isfinite(v)) mx = fmaxf(mx, v);
    }
    mx = block_reduce(mx, true, sh);
    if (threadIdx.x == 0) carry = 0ull;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long base = 0; base < n; base += FAB_SCAN_NT) {
        const long long i = base + threadIdx.x;
        unsigned long long q = 0ull;
        if (i < n) {
            const float v = lw[i];
            if (fab_isfinite(v) && mx > -CUDART_INF_F) {
                const double t = (double)__fsub_rn(v, mx);
                if (!(t < -60.0)) q = (unsigned long long)floor(__dmul_rn(exp_det(t), 1073741824.0));
            }
        }
        unsigned long long incl = q;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(FAB_FULL, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const unsigned long long w = wtot[lane];
            unsigned long long wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(FAB_FULL, wi, o);
                if (lane >= o) wi += t;
            }
            wtot[lane] = wi - w;
            if (lane == 31) chunk_total = wi;
        }
        __syncthreads();
        if (i < n) cdf[i] = carry + wtot[warp] + incl;
        __syncthreads();
        if (threadIdx.x == 0) carry += chunk_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) cdf[n] = carry;    // total S
}

__global__ void k_resample_search(const unsigned long long* __restrict__ cdf, long long n,
                                  unsigned int u0, long long* __restrict__ anc) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const unsigned __int128 S = cdf[n];
    const unsigned __int128 thr = (((unsigned __int128)(unsigned long long)k << 32) + u0) * S;
    const unsigned __int128 scale = (unsigned __int128)(unsigned long long)n << 32;
    long long lo = 0, hi = n;          // first i with cdf[i]*scale > thr
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if ((unsigned __int128)cdf[mid] * scale > thr) hi = mid; else lo = mid + 1;
    }
    // S == 0 (no finite log-weight at all): the search runs off the end; keep the identity map
    // instead of an out-of-range ancestor (the host raises on this state, resample.py)
    anc[k] = S == 0 ? k : (lo < n ? lo : n - 1);
}

__global__ void k_gather_rows(const float* __restrict__ src, float* __restrict__ dst,
                              const long long* __restrict__ anc, long long n, int rowf) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * rowf) return;
    const long long k = i / rowf;
    const int j = (int)(i - k * rowf);
    dst[i] = src[anc[k] * rowf + j];
}
// 16-byte form (row_floats % 4 == 0, both bases 16-byte aligned): a CTA copies a contiguous run
// of 256*U float4s of `dst`; one 64-bit division per thread finds the run's first row, the rest
// is 32-bit arithmetic.  All U loads are issued before the first store.
template <int U>
__global__ void __launch_bounds__(256)
k_gather_rows_v4(const float4* __restrict__ src, float4* __restrict__ dst,
                 const long long* __restrict__ anc, long long n, int rv) {
    const long long tot = n * rv;
    const long long base = (long long)blockIdx.x * (256 * U);
    const long long k0 = base / rv;
    const unsigned off0 = (unsigned)(base - k0 * rv) + threadIdx.x;
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const unsigned off = off0 + u * 256;
        if (base + threadIdx.x + u * 256 < tot) {
            const unsigned kr = off / (unsigned)rv, q = off - kr * (unsigned)rv;
            v[u] = __ldg(src + __ldg(anc + k0 + kr) * rv + q);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long long i = base + threadIdx.x + u * 256;
        if (i < tot) __stcs(dst + i, v[u]);
    }
}
