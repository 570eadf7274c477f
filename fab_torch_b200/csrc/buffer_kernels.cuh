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
        long long row = current_index + i;              // batch <= max_length: wraps at most once
        if (row >= max_length) row -= max_length;
        bx[row * d + j] = x[e];
        if (j == 0) { blw[row] = lw[i]; blq[row] = lq[i]; }
    }
}
// 16-byte form (d % 4 == 0, x and the ring 16-byte aligned): same run-per-CTA mapping as
// k_gather_rows_v4 (misc_kernels.cuh), reading the batch contiguously and writing ring rows.
template <int U>
__global__ void __launch_bounds__(256)
k_buffer_add_v4(float4* __restrict__ bx, float* __restrict__ blw, float* __restrict__ blq,
                long long max_length, int rv, long long current_index,
                const float4* __restrict__ x, const float* __restrict__ lw,
                const float* __restrict__ lq, long long batch) {
    const long long tot = batch * rv;
    const long long base = (long long)blockIdx.x * (256 * U);
    const long long k0 = base / rv;
    const unsigned off0 = (unsigned)(base - k0 * rv) + threadIdx.x;
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (base + threadIdx.x + u * 256 < tot) v[u] = __ldcs(x + base + threadIdx.x + u * 256);
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (base + threadIdx.x + u * 256 < tot) {
            const unsigned off = off0 + u * 256;
            const unsigned kr = off / (unsigned)rv, q = off - kr * (unsigned)rv;
            const long long i = k0 + kr;
            long long row = current_index + i;
            if (row >= max_length) row -= max_length;
            bx[row * rv + q] = v[u];
            if (q == 0) { blw[row] = lw[i]; blq[row] = lq[i]; }
        }
    }
}

// larger float <-> larger key; NaN sorts above +inf like torch.topk
__device__ __forceinline__ unsigned int fab_float_key(float v) {
    if (v != v) return 0xffffffffu;
    const unsigned int u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// ---- exact top-k of n keys at full-chip width (deterministic: integer histograms only) ----------
// Control block at the head of the workspace, zeroed by the host wrapper before every call.
#define FAB_SEL_NT 1024
#define FAB_SEL_MAXB 1024
struct fab_sel_ctl {
    unsigned int hist[4][256];                  // digit histograms of the four radix passes
    unsigned int prefix, mask;                  // digits fixed so far / their bit mask
    unsigned long long need;                    // keys still to take inside the current bucket
    unsigned int ticket[8];                     // last-CTA-done counters, one per launch
    unsigned int cnt[FAB_SEL_MAXB][2];          // per CTA: #(key > thr), #(key == thr)
    unsigned long long off[FAB_SEL_MAXB][2];    // exclusive prefix of cnt over the CTAs
};

// Warp-aggregated shared-memory histogram increment (the top digit of a float key takes a handful
// of values: plain atomics would serialise on one address).  digit >= 256: no vote.
__device__ __forceinline__ void sel_vote(unsigned int* sh_hist, unsigned int digit) {
    const unsigned int peers = __match_any_sync(FAB_FULL, digit);
    if (digit < 256u && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&sh_hist[digit], __popc(peers));
}

// Flush the CTA histogram; the last CTA to arrive picks the bucket that holds the k-th largest
// key: the digit b with above(b) < need <= above(b) + hist[b], above(b) = #keys in buckets > b.
__device__ __forceinline__ void sel_flush_and_pick(fab_sel_ctl* ctl, int pass, unsigned int* sh_hist,
                                                   unsigned int prefix, unsigned int mask, long long k) {
    __shared__ bool s_last;
    __shared__ unsigned long long s_wtot[8];
    __syncthreads();
    for (int b = threadIdx.x; b < 256; b += blockDim.x)
        if (sh_hist[b]) atomicAdd(&ctl->hist[pass][b], sh_hist[b]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&ctl->ticket[pass], 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long need = pass == 0 ? (unsigned long long)k : ctl->need;   // before any write
    unsigned long long h = 0ull, incl = 0ull;
    if (threadIdx.x < 256) {                    // thread r scans bucket 255 - r (largest digit first)
        h = ((volatile unsigned int*)ctl->hist[pass])[255 - threadIdx.x];
        incl = h;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(FAB_FULL, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_wtot[warp] = incl;
    }
    __syncthreads();
    if (threadIdx.x < 256) {
        unsigned long long above = incl - h;
        for (int w = 0; w < warp; ++w) above += s_wtot[w];
        if (above < need && need <= above + h) {
            const int shift = 24 - 8 * pass;
            ctl->need = need - above;
            ctl->prefix = prefix | ((255u - threadIdx.x) << shift);
            ctl->mask = mask | (255u << shift);
        }
    }
}

// keys + first radix pass (top digit).  CTA b owns keys [b*chunk, (b+1)*chunk).
__global__ void __launch_bounds__(FAB_SEL_NT)
k_buffer_keys(const float* __restrict__ logits, const float* __restrict__ gumbel, long long n,
              long long chunk, long long k, unsigned int* __restrict__ keys, fab_sel_ctl* ctl) {
    __shared__ unsigned int sh_hist[256];
    for (int b = threadIdx.x; b < 256; b += FAB_SEL_NT) sh_hist[b] = 0u;
    __syncthreads();
    const long long lo = (long long)blockIdx.x * chunk, hi = lo + chunk < n ? lo + chunk : n;
    for (long long base = lo; base < hi; base += FAB_SEL_NT) {
        const long long i = base + threadIdx.x;
        unsigned int digit = 256u;
        if (i < hi) {
            const unsigned int key = fab_float_key(__fadd_rn(gumbel[i], logits[i]));   // z + logits (:13)
            keys[i] = key;
            digit = key >> 24;
        }
        sel_vote(sh_hist, digit);
    }
    sel_flush_and_pick(ctl, 0, sh_hist, 0u, 0u, k);
}

// radix passes 1..3: histogram of the next digit over the keys that match the digits fixed so far
__global__ void __launch_bounds__(FAB_SEL_NT)
k_buffer_select_hist(const unsigned int* __restrict__ keys, long long n, long long chunk, int pass,
                     fab_sel_ctl* ctl) {
    __shared__ unsigned int sh_hist[256];
    for (int b = threadIdx.x; b < 256; b += FAB_SEL_NT) sh_hist[b] = 0u;
    __syncthreads();
    const unsigned int prefix = ctl->prefix, mask = ctl->mask;
    const int shift = 24 - 8 * pass;
    const long long lo = (long long)blockIdx.x * chunk, hi = lo + chunk < n ? lo + chunk : n;
    for (long long base = lo; base < hi; base += FAB_SEL_NT) {
        const long long i = base + threadIdx.x;
        unsigned int digit = 256u;
        if (i < hi) {
            const unsigned int key = keys[i];
            if ((key & mask) == prefix) digit = (key >> shift) & 255u;
        }
        sel_vote(sh_hist, digit);
    }
    sel_flush_and_pick(ctl, pass, sh_hist, prefix, mask, 0);
}

// threshold key = ctl->prefix.  Per-CTA counts of keys above / at the threshold; the last CTA
// turns them into exclusive prefixes over the CTAs.
__global__ void __launch_bounds__(FAB_SEL_NT)
k_buffer_select_count(const unsigned int* __restrict__ keys, long long n, long long chunk,
                      fab_sel_ctl* ctl) {
    __shared__ unsigned int s_gt, s_eq;
    __shared__ bool s_last;
    __shared__ unsigned long long s_w[2][32];
    if (threadIdx.x == 0) { s_gt = 0u; s_eq = 0u; }
    __syncthreads();
    const unsigned int thr = ctl->prefix;
    const long long lo = (long long)blockIdx.x * chunk, hi = lo + chunk < n ? lo + chunk : n;
    unsigned int gt = 0u, eq = 0u;
    for (long long i = lo + threadIdx.x; i < hi; i += FAB_SEL_NT) {
        const unsigned int key = keys[i];
        gt += key > thr; eq += key == thr;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        gt += __shfl_xor_sync(FAB_FULL, gt, o);
        eq += __shfl_xor_sync(FAB_FULL, eq, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_gt, gt); atomicAdd(&s_eq, eq); }
    __syncthreads();
    if (threadIdx.x == 0) {
        ctl->cnt[blockIdx.x][0] = s_gt; ctl->cnt[blockIdx.x][1] = s_eq;
        __threadfence();
        s_last = atomicAdd(&ctl->ticket[4], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long v[2], incl[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        v[q] = threadIdx.x < gridDim.x ? ((volatile unsigned int*)ctl->cnt[threadIdx.x])[q] : 0u;
        incl[q] = v[q];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(FAB_FULL, incl[q], o);
            if (lane >= o) incl[q] += t;
        }
        if (lane == 31) s_w[q][warp] = incl[q];
    }
    __syncthreads();
    if (threadIdx.x < gridDim.x) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            unsigned long long ex = incl[q] - v[q];
            for (int w = 0; w < warp; ++w) ex += s_w[q][w];
            ctl->off[threadIdx.x][q] = ex;
        }
    }
}

// out[0..k): indices of the k largest keys in ascending index order; ties at the threshold are
// resolved towards the lower index (the first `need` keys equal to it).
__global__ void __launch_bounds__(FAB_SEL_NT)
k_buffer_select_scatter(const unsigned int* __restrict__ keys, long long n, long long chunk,
                        const fab_sel_ctl* __restrict__ ctl, long long* __restrict__ out) {
    __shared__ unsigned int warp_tot[FAB_SEL_NT / 32];
    __shared__ unsigned int s_chunk_tot;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int thr = ctl->prefix;
    const unsigned long long need = ctl->need;
    unsigned long long gt_run = ctl->off[blockIdx.x][0], eq_run = ctl->off[blockIdx.x][1];
    const long long lo = (long long)blockIdx.x * chunk, hi = lo + chunk < n ? lo + chunk : n;
    for (long long base = lo; base < hi; base += FAB_SEL_NT) {
        const long long i = base + threadIdx.x;
        const unsigned int key = i < hi ? keys[i] : 0u;
        const unsigned int gt = (i < hi && key > thr) ? 1u : 0u;
        const unsigned int eq = (i < hi && key == thr) ? 1u : 0u;
        const unsigned int v = gt | (eq << 16);     // both counts (<= 1024) scanned in one word
        unsigned int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t = __shfl_up_sync(FAB_FULL, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const unsigned int w = warp_tot[lane];
            unsigned int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int t = __shfl_up_sync(FAB_FULL, wi, o);
                if (lane >= o) wi += t;
            }
            warp_tot[lane] = wi - w;
            if (lane == 31) s_chunk_tot = wi;
        }
        __syncthreads();
        const unsigned int excl = warp_tot[warp] + incl - v, tot = s_chunk_tot;
        const unsigned long long gt_before = gt_run + (excl & 0xffffu);
        const unsigned long long eq_before = eq_run + (excl >> 16);
        if (gt || (eq && eq_before < need))
            out[gt_before + (eq_before < need ? eq_before : need)] = i;
        gt_run += tot & 0xffffu;
        eq_run += tot >> 16;
        __syncthreads();                            // warp_tot / s_chunk_tot are reused
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
