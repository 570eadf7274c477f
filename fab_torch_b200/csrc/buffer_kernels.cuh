// Prioritised replay buffer on the device (fab/utils/prioritised_replay_buffer.py:71-131; SURVEY
// §8f row 1, BASELINE config 5).  HBM-bound integer / copy work:
//   k_buffer_add     ring write of a batch (x, log_w, log_q_old) at (current_index + i) % max_length
//   k_buffer_keys    order-preserving uint32 key of (gumbel + log_w)           [:88-100, :10-14]
//   k_buffer_select  exact top-k of the keys by 4-pass radix select (one CTA, deterministic):
//                    the Gumbel-top-k "sample without replacement" index set
//   k_buffer_adjust  log_w[idx] += adjustment, log_q_old[idx] = log_q for finite entries,
//                    log_w[idx] = -inf for the others                            [:117-131]
#pragma once
#include "common.cuh"

__global__ void k_buffer_add(float* __restrict__ bx, float* __restrict__ blw, float* __restrict__ blq,
                             long long max_length, int d, long long current_index,
                             const float* __restrict__ x, const float* __restrict__ lw,
                             const float* __restrict__ lq, long long batch) {
    const long long tot = batch * d;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < tot;
         e += (long long)gridDim.x * blockDim.x) {
        const long long i = e / d;
        const int j = (int)(e - i * d);
        const long long row = (current_index + i) % max_length;
        bx[row * d + j] = x[e];
        if (j == 0) { blw[row] = lw[i]; blq[row] = lq[i]; }
    }
}

// larger float <-> larger key; NaN sorts above +inf like torch.topk
__device__ __forceinline__ unsigned int fab_float_key(float v) {
    if (v != v) return 0xffffffffu;
    const unsigned int u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void k_buffer_keys(const float* __restrict__ logits, const float* __restrict__ gumbel,
                              long long n, unsigned int* __restrict__ keys) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        keys[i] = fab_float_key(__fadd_rn(gumbel[i], logits[i]));       // z + logits  (:13)
}

#define FAB_SEL_NT 1024
// out[0..k): indices of the k largest keys, ascending index order; ties at the threshold are
// resolved towards the lower index.
__global__ void __launch_bounds__(FAB_SEL_NT)
k_buffer_select(const unsigned int* __restrict__ keys, long long n, long long k,
                long long* __restrict__ out) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned int s_prefix, s_mask;
    __shared__ long long s_need;
    __shared__ int warp_tot[FAB_SEL_NT / 32];
    __shared__ long long s_base, s_ties_left;
    __shared__ int s_chunk_tot;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_prefix = 0u; s_mask = 0u; s_need = k; }
    __syncthreads();
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int b = threadIdx.x; b < 256; b += FAB_SEL_NT) hist[b] = 0u;
        __syncthreads();
        const unsigned int prefix = s_prefix, mask = s_mask;
        for (long long i = threadIdx.x; i < n; i += FAB_SEL_NT) {
            const unsigned int key = keys[i];
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            long long need = s_need;
            int b = 255;
            for (; b > 0; --b) {                   // walk from the largest digit down
                if ((long long)hist[b] >= need) break;
                need -= hist[b];
            }
            s_need = need;                          // still to take inside bucket b
            s_prefix = prefix | ((unsigned int)b << shift);
            s_mask = mask | (255u << shift);
        }
        __syncthreads();
    }
    // threshold key = s_prefix; take every key > thr and the first s_need keys == thr
    const unsigned int thr = s_prefix;
    if (threadIdx.x == 0) { s_base = 0; s_ties_left = s_need; }
    __syncthreads();
    for (long long base = 0; base < n; base += FAB_SEL_NT) {
        const long long i = base + threadIdx.x;
        const unsigned int key = i < n ? keys[i] : 0u;
        const int gt = (i < n && key > thr) ? 1 : 0;
        const int eq = (i < n && key == thr) ? 1 : 0;
        // block-wide exclusive scans of gt and eq (two rounds of the same routine)
        int vals[2] = {gt, eq}, excl[2], tot[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            int incl = vals[q];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FAB_FULL, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) warp_tot[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                const int w = warp_tot[lane];
                int wi = w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(FAB_FULL, wi, o);
                    if (lane >= o) wi += t;
                }
                warp_tot[lane] = wi - w;                  // exclusive prefix of the warp totals
                if (lane == 31) s_chunk_tot = wi;         // total of the chunk
            }
            __syncthreads();
            excl[q] = warp_tot[warp] + incl - vals[q];
            tot[q] = s_chunk_tot;
            __syncthreads();                              // warp_tot / s_chunk_tot are reused
        }
        const long long ties_left = s_ties_left, obase = s_base;
        const int eq_taken_before = excl[1] < ties_left ? excl[1] : (int)ties_left;
        const int take_eq = eq && excl[1] < ties_left;
        if (gt || take_eq) out[obase + excl[0] + eq_taken_before] = i;
        __syncthreads();
        if (threadIdx.x == 0) {
            const long long eq_taken = tot[1] < ties_left ? tot[1] : ties_left;
            s_base = obase + tot[0] + eq_taken;
            s_ties_left = ties_left - eq_taken;
        }
        __syncthreads();
    }
}

__global__ void k_buffer_adjust(float* __restrict__ blw, float* __restrict__ blq,
                                const long long* __restrict__ idx, const float* __restrict__ adj,
                                const float* __restrict__ lq, long long m) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const long long r = idx[i];
    if (fab_isfinite(adj[i]) && fab_isfinite(lq[i])) {
        blw[r] = __fadd_rn(blw[r], adj[i]);
        blq[r] = lq[i];
    } else {
        blw[r] = -CUDART_INF_F;
    }
}
