// Tile GEMM on the warp-level tensor path: out[p][n] = sum_k act[p][k] * M[k][n] for the TP particle
// slots of one CTA, error-compensated "3xTF32" so that the result carries fp32 accuracy.
//
// Why this engine (measured on B200, profiles/microbench_hmma.cu): one CTA owns at
// most 16 particles, far below the 128-row tile a tcgen05.mma needs, and with the weights as the
// 128-row operand the UMMA is bound by reading that operand from shared memory.  The warp-level
// mma.sync.m16n8k8 (SASS HMMA.1688.F32.TF32) takes exactly the shape we have -- 16 particle
// slots x 8 output columns per instruction -- and issues once per 8 cycles per SM sub-partition
// = 297 dense TF32 TFLOP/s chip-wide.  With the three-product split (hi*hi + hi*lo + lo*hi, fp32
// accumulate) that is 99 TFLOP/s of fp32-grade work versus the 74.4 TFLOP/s FP32 FMA ceiling,
// and the FMA/ALU pipes stay free for the fused epilogues.
//
//   act : shared memory, k-major: element (slot p, row k) at act[k*S + p]; S = 24 for TP = 16
//         (fragment loads hit 32 distinct banks), 8 for TP = 8 (slots 8..15 of the MMA tile are
//         fed zeros).  Rows are padded to whole k-tile pairs (16) with finite values.
//   Wf  : global (L2-resident) weights in FRAGMENT ORDER, float4 Wf[KT2][NT][32]:
//         Wf[kp][nt][lane] = { M[16kp+t][8nt+g], M[16kp+t+4][8nt+g], M[16kp+8+t][8nt+g],
//         M[16kp+12+t][8nt+g] },  g = lane>>2, t = lane&3  -- the B fragments of two consecutive
//         k-tiles, so one coalesced LDG.128 per lane feeds six MMAs.  Each word is loaded once per
//         CTA, straight into registers, three k-tile pairs ahead of its use.  The host stores
//         the weights rounded to 22 significant bits (11 hi + 11 lo) so that the hi/lo split
//         below is exact for them.
//   split: v = hi + lo with hi = v & 0xffffe000 (the tf32 grid) and lo = v - hi (exact), rounded
//         to the tf32 grid (the tensor core itself would truncate the low 13 bits of lo).  Products kept: lo*hi, hi*lo, hi*hi (small
//         terms first).  A non-finite activation gives lo = NaN, i.e. inf degrades to NaN; both
//         are "not finite" to every consumer (ais.py:190-213, hmc.py:113-117).
//
// Two routines: mma_gemm_wide (N >= 64: the 8 warps split the n-tiles, epilogue straight from
// the accumulator fragments -- no partial sums, no second barrier) and mma_gemm_ksplit (narrow
// N: warps split K, partials meet in shared memory).
#pragma once
#include "common.cuh"

// Mean relative shrink of one k-tile partial sum (8 terms, truncating tensor-core adder), measured
// on B200 for signed, ReLU-like and positive operands alike: -1.5 * 2^-24
// (profiles/microbench_tf32_bias.cu).  Every GEMM result is scaled back by (1 + FAB_TRUNC_EPS).
#define FAB_TRUNC_EPS 8.94069671630859375e-8f

#define FAB_NTW 5     // n-tiles per warp and pass of the wide GEMM (8 warps x 5 x 8 = 320 columns)
#define FAB_NTK 4     // n-tiles per warp and pass of the k-split GEMM

template <int TP> struct ActL {
    static constexpr int S = TP == 16 ? 24 : 8;       // operand slot stride
    static constexpr int RS = TP == 16 ? 20 : 12;     // partial-sum slot stride (conflict-free stores)
};

// Weight words are read once per CTA (no L1 allocation) but by every CTA and every launch: they
// must stay in L2.  Plain loads insert at normal priority, and once L2 is full of other data at
// the same priority (e.g. after a 256 MB fill between calls) the streamed blob keeps missing:
// measured +15 % cycles per k_hmc_step launch, persisting long after the fill
// (profiles/ab_clock.py).  The evict_last policy keeps the blob resident.
#ifndef FAB_NO_L2_HINT
__device__ __forceinline__ uint64_t l2_keep_policy() {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(l2_keep_policy()));
    return r;
}
#else
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
#endif

__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(v) & 0xffffe000u;
    // + 0x1000: the tensor core truncates its operands to the tf32 grid, which would shrink
    // every product by ~2e-7 (a multiplicative bias that compounds over the layers); adding half
    // a tf32 ulp first turns that truncation into round-to-nearest
    lo = __float_as_uint(__fsub_rn(v, __uint_as_float(hi))) + 0x1000u;
}

// weights are stored pre-rounded to hi(11 bits) + lo(11 bits) (flow.py: _round22), so their lo
// part already sits on the tf32 grid and the split is exact
__device__ __forceinline__ void split_w(float v, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(v) & 0xffffe000u;
    lo = __float_as_uint(__fsub_rn(v, __uint_as_float(hi)));
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A fragments (hi, lo) of the k-tile whose first row is `kb`
template <int TP>
__device__ __forceinline__ void load_a(const float* act, int kb, int g, int t, uint32_t (&ah)[4],
                                       uint32_t (&al)[4]) {
    constexpr int S = ActL<TP>::S;
    const float* ap = act + (size_t)(kb + t) * S + g;
    split_tf32(ap[0], ah[0], al[0]);
    split_tf32(ap[4 * S], ah[2], al[2]);
    if (TP == 16) {
        split_tf32(ap[8], ah[1], al[1]);
        split_tf32(ap[4 * S + 8], ah[3], al[3]);
    } else {
        ah[1] = al[1] = ah[3] = al[3] = 0u;
    }
}

// one k-tile pair for NT_ n-tiles: c[i] += A(2 k-tiles) * w[i].
// The tensor core truncates toward zero -- every addend to ~2^-25 of the largest one and every
// result to fp32 (measured: profiles/microbench_mma_rounding.cu, microbench_tf32_accuracy.cu; a
// 120-MMA chain into one accumulator drifts by -7e-6 relative on same-sign data, and even a 6-MMA
// chain shrinks every output by ~2e-7, which compounds over the flow layers).  So the three
// products of ONE k-tile go into a fresh accumulator, small terms first (they meet a near-empty
// accumulator), and that 8-term partial sum is added to the running total with a round-to-nearest
// FADD: one truncation per k-tile at partial-sum magnitude, with signs that average out.
// CNT > 0: exactly the first CNT tiles are live (compile time, no predicates around the MMAs);
// CNT = 0: the first `cnt` tiles are live (run time, predicated).
template <int TP, int NT_, int CNT>
__device__ __forceinline__ void mma_pair(float (&c)[NT_][4], const float* act, int kb, int g, int t,
                                         const float4 (&w)[NT_], int cnt) {
    constexpr bool FULL = CNT > 0;
    constexpr int NL = CNT > 0 ? CNT : NT_;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint32_t ah[4], al[4];
        load_a<TP>(act, kb + 8 * h, g, t, ah, al);
        uint32_t bh[NT_][2], bl[NT_][2];
        float cp[NT_][4];
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            split_w(h == 0 ? w[i].x : w[i].z, bh[i][0], bl[i][0]);
            split_w(h == 0 ? w[i].y : w[i].w, bh[i][1], bl[i][1]);
            cp[i][0] = cp[i][1] = cp[i][2] = cp[i][3] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < NL; ++i) if (FULL || i < cnt) mma_tf32(cp[i], al, bh[i][0], bh[i][1]);
#pragma unroll
        for (int i = 0; i < NL; ++i) if (FULL || i < cnt) mma_tf32(cp[i], ah, bl[i][0], bl[i][1]);
#pragma unroll
        for (int i = 0; i < NL; ++i) if (FULL || i < cnt) mma_tf32(cp[i], ah, bh[i][0], bh[i][1]);
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            c[i][0] += cp[i][0]; c[i][1] += cp[i][1]; c[i][2] += cp[i][2]; c[i][3] += cp[i][3];
        }
    }
}

// Optional L1 prefetch of the first weight words a following GEMM will read (the first
// FAB_PF_DEPTH k-tile pairs of the warp's own tiles; KT2 = k-tile pairs of that GEMM).
// OFF by default: measured on B200 with the L2 evict_last policy in place, chain time at depth
// 0 / 1 / 2 / 3 = 24.24 / 24.77 / 25.74 / 25.81 ms (profiles/ab_chain.py) -- the no-allocate
// loads do not hit the prefetched lines, so a prefetch is a second L2 read of the same line.
// Two deeper variants were measured as well and are slower still: issuing the prefetch one GEMM
// earlier (-DFAB_PF_EARLY, 24.72 at depth 1) and carrying the register ring ACROSS GEMMs so that
// the next GEMM's first three stages load during the epilogue and barrier (28.3 ms: the ring
// registers stay live through every epilogue; profiles/experiments/cross_gemm_ring.patch).
#ifndef FAB_PF_DEPTH
#define FAB_PF_DEPTH 0
#endif
__device__ __forceinline__ void mma_prefetch(const float4* __restrict__ Wf, int NT, bool ksplit, int KT2) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int nt0, kp0, stride, cnt, kstep;
    if (ksplit) {
        int NGR, KS; fab_ksplit_plan(NT, NGR, KS);
        nt0 = warp % NGR; kp0 = warp / NGR; stride = NGR; cnt = FAB_NTK; kstep = KS < KT2 ? KS : KT2;
    } else {
        nt0 = warp; kp0 = 0; stride = 8; cnt = FAB_NTW; kstep = 1;
    }
    // (a k-split warp whose first k-pair does not exist prefetches a word of the next operand:
    // harmless, the blob is contiguous and padded)
#pragma unroll
    for (int dd = 0; dd < FAB_PF_DEPTH; ++dd) {
        const int kp = kp0 + dd * kstep;
        if (dd > 0 && kp >= KT2) break;
        const float4* p = Wf + ((size_t)kp * NT + nt0) * 32 + lane;
        for (int i = 0; i < cnt; ++i)
            if (nt0 + stride * i < NT) asm volatile("prefetch.global.L1 [%0];" ::"l"(p + (size_t)i * stride * 32));
    }
}

// Wide GEMM.  Warp w owns the n-tiles nt = base + w + 8*i (i < FAB_NTW) of every pass
// base = 0, 40, 80, ...; accumulators start from the bias (nullptr = 0).  epi(nt, c) is called
// warp-uniformly once per owned tile with the lane's fragment: c[0],c[1] = (slot g, columns
// 8nt+2t, +1), c[2],c[3] = (slot g+8, same columns).  The caller must __syncthreads() before the
// operand `act` is overwritten and before anything the epilogue wrote is read.
template <int TP, int NTW, class Epi>
__device__ __forceinline__ void mma_gemm_wide_n(const float* act, int KT2, const float4* __restrict__ Wf,
                                              int NT, const float* __restrict__ bias, Epi epi) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int npass = (NT + 8 * NTW - 1) / (8 * NTW);
    const int total = npass * KT2;
    // prefetch cursor (pass, kp) of the weight ring.  (A leaner cursor -- running pointer, per-pass
    // tile count, unpredicated loads for full passes -- executes fewer instructions but was
    // measured slower, 25.46 vs 24.24 ms per chain: its extra branches cost more than the
    // address arithmetic it saves.)
    int pf_base = 0, pf_kp = 0, pf_it = 0;
    auto fetch = [&](float4 (&dst)[NTW]) {
        if (pf_it < total) {
            const float4* p = Wf + ((size_t)pf_kp * NT + pf_base + warp) * 32 + lane;
#pragma unroll
            for (int i = 0; i < NTW; ++i)
                if (pf_base + warp + 8 * i < NT) dst[i] = ldg_stream(p + (size_t)i * 8 * 32);
            ++pf_it;
            if (++pf_kp == KT2) { pf_kp = 0; pf_base += 8 * NTW; }
        }
    };
    // three-deep register ring: the main loop is unrolled by three so that no register moves are
    // needed; each buffer is refilled (for the pair three steps ahead) right after its use.
    // (profiles/microbench_mma_variants.cu times the alternatives in isolation -- two-deep ring,
    // packed f32x2 adds/splits, k-tile software pipelining, 16 warps: none beats this form in
    // the full kernel, where the deeper prefetch also covers the short GEMMs' L2 latency.)
    float4 w0[NTW] = {}, w1[NTW] = {}, w2[NTW] = {};
    fetch(w0);
    fetch(w1);
    fetch(w2);
    for (int base = 0; base < NT; base += 8 * NTW) {
        int cnt = (NT - base - warp + 7) / 8;                 // tiles this warp owns in the pass
        if (cnt > NTW) cnt = NTW;
        if (cnt < 0) cnt = 0;
        float c[NTW][4];
#pragma unroll
        for (int i = 0; i < NTW; ++i) {
            float b0 = 0.f, b1 = 0.f;
            if (bias && i < cnt) {
                const float2 bv = *reinterpret_cast<const float2*>(bias + (base + warp + 8 * i) * 8 + 2 * t);
                b0 = bv.x; b1 = bv.y;
            }
            // (1 - eps) here and (1 + eps) on the way out leave the bias unscaled
            b0 = fmaf(-FAB_TRUNC_EPS, b0, b0); b1 = fmaf(-FAB_TRUNC_EPS, b1, b1);
            c[i][0] = b0; c[i][1] = b1; c[i][2] = b0; c[i][3] = b1;
        }
        int kp = 0;
#ifdef FAB_RING_PEEL
        // Experiment knob (default off, NOT yet measured): branch-free steady state for single-pass
        // GEMMs.  While three more stages exist the refills need no guard, no tile predicate and
        // no address arithmetic beyond one running pointer; the guarded loops below drain the
        // last 3..5 k-tile pairs.  (profiles/r01_hot_lines.md: the guarded cursor is 16 % of the
        // kernel's instructions; a leaner but branchy cursor was slower.)
        if (npass == 1 && cnt >= NTW - 1 && KT2 >= 6) {
            const size_t kstr = (size_t)NT * 32;
            const float4* q = Wf + (3 * kstr + (size_t)warp * 32 + lane);      // stage 3
#define FAB_PEEL3(CNT)                                                                        \
            for (; kp + 6 <= KT2; kp += 3) {                                                   \
                mma_pair<TP, NTW, CNT>(c, act, kp * 16, g, t, w0, cnt);                        \
                _Pragma("unroll") for (int i = 0; i < CNT; ++i) w0[i] = ldg_stream(q + (size_t)i * 8 * 32);            \
                mma_pair<TP, NTW, CNT>(c, act, kp * 16 + 16, g, t, w1, cnt);                   \
                _Pragma("unroll") for (int i = 0; i < CNT; ++i) w1[i] = ldg_stream(q + kstr + (size_t)i * 8 * 32);     \
                mma_pair<TP, NTW, CNT>(c, act, kp * 16 + 32, g, t, w2, cnt);                   \
                _Pragma("unroll") for (int i = 0; i < CNT; ++i) w2[i] = ldg_stream(q + 2 * kstr + (size_t)i * 8 * 32); \
                q += 3 * kstr;                                                                 \
            }
            if (cnt == NTW) { FAB_PEEL3(NTW) }
            else { FAB_PEEL3(NTW - 1) }
#undef FAB_PEEL3
            pf_kp = pf_it = kp + 3;             // stages fetched so far; the guarded cursor takes over
        }
#endif
#define FAB_RING3(CNT)                                                        \
        for (; kp + 3 <= KT2; kp += 3) {                                       \
            mma_pair<TP, NTW, CNT>(c, act, kp * 16, g, t, w0, cnt);            \
            fetch(w0);                                                         \
            mma_pair<TP, NTW, CNT>(c, act, kp * 16 + 16, g, t, w1, cnt);       \
            fetch(w1);                                                         \
            mma_pair<TP, NTW, CNT>(c, act, kp * 16 + 32, g, t, w2, cnt);       \
            fetch(w2);                                                         \
        }
        if (cnt == NTW) { FAB_RING3(NTW) }
        else if (cnt == NTW - 1) { FAB_RING3(NTW - 1) }
#undef FAB_RING3
        for (; kp < KT2; ++kp) {                 // tail / partially filled passes: rotate by moves
            float4 w[NTW];
#pragma unroll
            for (int i = 0; i < NTW; ++i) { w[i] = w0[i]; w0[i] = w1[i]; w1[i] = w2[i]; }
            fetch(w2);
            if (cnt == NTW) mma_pair<TP, NTW, NTW>(c, act, kp * 16, g, t, w, cnt);
            else if (cnt == NTW - 1) mma_pair<TP, NTW, NTW - 1>(c, act, kp * 16, g, t, w, cnt);
            else mma_pair<TP, NTW, 0>(c, act, kp * 16, g, t, w, cnt);
        }
#pragma unroll
        for (int i = 0; i < NTW; ++i) {
            if (i < cnt) {
#pragma unroll
                for (int q = 0; q < 4; ++q) c[i][q] = fmaf(c[i][q], FAB_TRUNC_EPS, c[i][q]);
                epi(base + warp + 8 * i, c[i]);
            }
        }
    }
}

// default tile count per warp
template <int TP, class Epi>
__device__ __forceinline__ void mma_gemm_wide(const float* act, int KT2, const float4* __restrict__ Wf,
                                              int NT, const float* __restrict__ bias, Epi epi) {
    mma_gemm_wide_n<TP, FAB_NTW>(act, KT2, Wf, NT, bias, epi);
}

// k-split GEMM for narrow outputs.  Warp = ks*NGR + ngr owns n-tiles nt = base + ngr + NGR*i
// (i < FAB_NTK) and the k-tile pairs kp = ks, ks+KSe, ...; partial sums go to
// red[(ks*N8 + n)*RS + slot], N8 = 8*NT.  Returns KSe, the number of partials per output.
// The caller must __syncthreads() before reading `red`.
template <int TP>
__device__ __forceinline__ int mma_gemm_ksplit(const float* act, int KT2, const float4* __restrict__ Wf,
                                               int NT, float* red) {
    constexpr int RS = ActL<TP>::RS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    int NGR, KS; fab_ksplit_plan(NT, NGR, KS);
    const int KSe = KS < KT2 ? KS : KT2;
    const int ngr = warp % NGR, ks = warp / NGR;
    if (ks >= KSe) return KSe;
    const int N8 = NT * 8;
    const int my_k = (KT2 - ks + KSe - 1) / KSe;              // k-pairs of this warp per pass
    const int pass_tiles = NGR * FAB_NTK;
    const int npass = (NT + pass_tiles - 1) / pass_tiles;
    const int total = npass * my_k;
    int pf_base = 0, pf_j = 0, pf_it = 0;
    auto fetch = [&](float4 (&dst)[FAB_NTK]) {
        if (pf_it < total) {
            const float4* p = Wf + ((size_t)(ks + pf_j * KSe) * NT + pf_base + ngr) * 32 + lane;
#pragma unroll
            for (int i = 0; i < FAB_NTK; ++i)
                if (pf_base + ngr + NGR * i < NT) dst[i] = ldg_stream(p + (size_t)i * NGR * 32);
            ++pf_it;
            if (++pf_j == my_k) { pf_j = 0; pf_base += pass_tiles; }
        }
    };
    float4 w0[FAB_NTK] = {}, w1[FAB_NTK] = {}, w2[FAB_NTK] = {};
    fetch(w0);
    fetch(w1);
    fetch(w2);
    for (int base = 0; base < NT; base += pass_tiles) {
        int cnt = (NT - base - ngr + NGR - 1) / NGR;
        if (cnt > FAB_NTK) cnt = FAB_NTK;
        if (cnt < 0) cnt = 0;
        float c[FAB_NTK][4];
#pragma unroll
        for (int i = 0; i < FAB_NTK; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
        for (int j = 0; j < my_k; ++j) {
            float4 w[FAB_NTK];
#pragma unroll
            for (int i = 0; i < FAB_NTK; ++i) { w[i] = w0[i]; w0[i] = w1[i]; w1[i] = w2[i]; }
            fetch(w2);
            if (cnt == FAB_NTK) mma_pair<TP, FAB_NTK, FAB_NTK>(c, act, (ks + j * KSe) * 16, g, t, w, cnt);
            else mma_pair<TP, FAB_NTK, 0>(c, act, (ks + j * KSe) * 16, g, t, w, cnt);
        }
#pragma unroll
        for (int i = 0; i < FAB_NTK; ++i) {
            if (i < cnt) {
                float* r = red + ((size_t)ks * N8 + (base + ngr + NGR * i) * 8 + 2 * t) * RS + g;
                r[0] = fmaf(c[i][0], FAB_TRUNC_EPS, c[i][0]);
                r[RS] = fmaf(c[i][1], FAB_TRUNC_EPS, c[i][1]);
                if (TP == 16) {
                    r[8] = fmaf(c[i][2], FAB_TRUNC_EPS, c[i][2]);
                    r[RS + 8] = fmaf(c[i][3], FAB_TRUNC_EPS, c[i][3]);
                }
            }
        }
    }
    return KSe;
}

// sum of the k-split partials of output (slot p, column n)
template <int TP>
__device__ __forceinline__ float red_sum(const float* red, int KSe, int N8, int p, int n) {
    constexpr int RS = ActL<TP>::RS;
    const float* r = red + (size_t)n * RS + p;
    float s = r[0];
    const int stride = N8 * RS;
#pragma unroll 4
    for (int ks = 1; ks < KSe; ++ks) { r += stride; s += r[0]; }
    return s;
}
