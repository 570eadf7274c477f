// RealNVP evaluation for the particles of one CTA, entirely in shared memory.
//
//   flow_inverse  : x -> z, log q(x)             normflows NormalizingFlow.log_prob  (K3)
//   flow_backward : d log q / d x                analytic reverse sweep, replaces
//                                                torch.autograd.grad in base.py:50-56 (K4)
//   flow_sample   : eps -> x, log q(x)           normflows NormalizingFlow.sample     (K1,K2)
//
// Layer semantics follow SURVEY Appendix B / oracle/realnvp.py (the restated normflows):
//   sampling, layer k:  [v1,v2] = u;  (shift,scale) = MLP(v1);  y2 = v2*exp(scale)+shift;
//                       u' = [v1,y2] @ Wmix^-1;  log q -= sum(scale) - sum(log_S)
//   inverse,  layer k:  v = u @ Wmix; (shift,scale) = MLP(v1);  y2 = (v2-shift)*exp(-scale);
//                       u' = [v1,y2];           log q += sum(log_S) - sum(scale)
// Per layer and direction this is three GEMMs (include/fab_b200.h: o_mw1/o_w2/o_w3 and
// o_w3t/o_w2t/o_w1mt): the mixing matrix is merged into the neighbouring MLP GEMM.  The two wide
// GEMMs of a direction finish in their accumulator fragments (bias = accumulator init, ReLU +
// mask ballots / mask select in registers, results stored as the next GEMM's operand); the narrow
// third GEMM splits K over the warps and its partials meet in the coupling step.
//
// zs0/zs1, gs, par, h1, h2 are MMA operands in the k-major layout of mma_gemm.cuh: element
// (slot p, row n) of any of them sits at float index n*S + p.
#pragma once
#include "mma_gemm.cuh"

// FAB_PF_EARLY (experiment knob): issue the L1 prefetch of a GEMM's first weight words before the
// PREVIOUS GEMM starts instead of behind it.
#ifdef FAB_PF_EARLY
#define PF_E(x) x
#define PF_L(x)
#else
#define PF_E(x)
#define PF_L(x) x
#endif

struct TileBufs {
    float *gs, *par, *h1, *h2, *red, *sy2, *ses, *ld, *scl;
    float *loc, *lsc, *inv, *logs;          // staged constants
    uint32_t *m1, *m2;
};

// The one dynamic shared-memory array of every tile kernel.  Each device function re-derives its
// buffer pointers from it (instead of receiving a struct of generic pointers) so that the
// compiler knows the address space and emits LDS/STS, and nothing is parked in local memory.
extern __shared__ __align__(16) float fab_smem[];

// the two ping-pong operand buffers of the running latent (offset select keeps the address space)
__device__ __forceinline__ float* zsel(const TileLayout& L, int which) {
    return fab_smem + (which ? L.o_zs1 : L.o_zs0);
}

__device__ __forceinline__ TileBufs tile_bufs(const TileLayout& L) {
    float* const smem = fab_smem;
    TileBufs b;
    b.gs = smem + L.o_gs;
    b.par = smem + L.o_par; b.h1 = smem + L.o_h1;   b.h2 = smem + L.o_h2;
    b.red = smem + L.o_red; b.sy2 = smem + L.o_sy2; b.ses = smem + L.o_ses;
    b.ld = smem + L.o_ld;   b.scl = smem + L.o_scl;
    b.loc = smem + L.o_const; b.lsc = b.loc + L.DP; b.inv = b.lsc + L.DP; b.logs = b.inv + L.DP;
    b.m1 = reinterpret_cast<uint32_t*>(smem + L.o_m1);
    b.m2 = reinterpret_cast<uint32_t*>(smem + L.o_m2);
    return b;
}

// Once per kernel: zero the operand buffers (pad rows/slots must stay finite because the MMAs
// multiply them by zero weights / feed them to unused accumulator rows) and stage the small
// per-flow constants (base loc / log_scale, per-layer sum(log_S)) in shared memory so that no
// serial code path waits on an L2 round trip.
__device__ __forceinline__ void tile_init(const TileLayout& L, const fab_flow_desc& f, const float* __restrict__ blob) {
    const TileBufs b = tile_bufs(L);
    float* const smem = fab_smem;
    for (int i = threadIdx.x; i < L.o_red; i += FAB_NT) smem[i] = 0.f;     // zs0..h2 are contiguous
    for (int j = threadIdx.x; j < L.DP; j += FAB_NT) {
        const float loc = j < L.d ? __ldg(blob + f.off_base_loc + j) : 0.f;
        const float ls = j < L.d ? __ldg(blob + f.off_base_log_scale + j) : 0.f;
        b.loc[j] = loc; b.lsc[j] = ls; b.inv[j] = expf(-ls);
    }
    for (int k = threadIdx.x; k < L.K; k += FAB_NT)
        b.logs[k] = __ldg(blob + f.off_layers + (size_t)k * f.layer_stride + f.o_logs);
    if (threadIdx.x == 0) {                      // sum_k sum(log_S_k), added to every log-det at once
        float tot = 0.f;
        for (int k = 0; k < L.K; ++k)
            tot += __ldg(blob + f.off_layers + (size_t)k * f.layer_stride + f.o_logs);
        b.logs[L.K] = tot;
    }
    __syncthreads();
}

// ---- accumulator-fragment epilogues of the wide GEMMs ------------------------------------------
// store fragment (tile column block n0 = 8*tile) as rows n0+2t, n0+2t+1 of an operand buffer
template <int TP>
__device__ __forceinline__ void frag_store(float* dst, int n0, int g, int t, float c0, float c1,
                                           float c2, float c3) {
    constexpr int S = ActL<TP>::S;
    float* r = dst + (size_t)(n0 + 2 * t) * S + g;
    r[0] = c0; r[S] = c1;
    if (TP == 16) { r[8] = c2; r[S + 8] = c3; }
}

// h = relu(c) -> dst tile `ht`; with SAVE the ReLU masks are kept as one ballot per fragment
// register (word layout [ht][TP==16 ? 4 : 2])
template <int TP, bool SAVE>
__device__ __forceinline__ void hidden_fwd(float* dst, uint32_t* mask, int ht, int g, int t,
                                           const float (&c)[4]) {
    const bool p0 = c[0] > 0.f, p1 = c[1] > 0.f, p2 = c[2] > 0.f, p3 = c[3] > 0.f;
    if (SAVE) {
        const uint32_t b0 = __ballot_sync(FAB_FULL, p0), b1 = __ballot_sync(FAB_FULL, p1);
        if (TP == 16) {
            const uint32_t b2 = __ballot_sync(FAB_FULL, p2), b3 = __ballot_sync(FAB_FULL, p3);
            if ((threadIdx.x & 31) == 0) *reinterpret_cast<uint4*>(mask + 4 * ht) = make_uint4(b0, b1, b2, b3);
        } else {
            if ((threadIdx.x & 31) == 0) *reinterpret_cast<uint2*>(mask + 2 * ht) = make_uint2(b0, b1);
        }
    }
    frag_store<TP>(dst, 8 * ht, g, t, p0 ? c[0] : 0.f, p1 ? c[1] : 0.f, p2 ? c[2] : 0.f,
                   p3 ? c[3] : 0.f);
}

// gh = mask ? c : 0 -> dst tile `ht`
template <int TP>
__device__ __forceinline__ void hidden_bwd(float* dst, const uint32_t* mask, int ht, int g, int t,
                                           const float (&c)[4]) {
    const int lane = threadIdx.x & 31;
    if (TP == 16) {
        const uint4 m = *reinterpret_cast<const uint4*>(mask + 4 * ht);
        frag_store<TP>(dst, 8 * ht, g, t, (m.x >> lane) & 1u ? c[0] : 0.f, (m.y >> lane) & 1u ? c[1] : 0.f,
                       (m.z >> lane) & 1u ? c[2] : 0.f, (m.w >> lane) & 1u ? c[3] : 0.f);
    } else {
        const uint2 m = *reinterpret_cast<const uint2*>(mask + 2 * ht);
        frag_store<TP>(dst, 8 * ht, g, t, (m.x >> lane) & 1u ? c[0] : 0.f, (m.y >> lane) & 1u ? c[1] : 0.f,
                       0.f, 0.f);
    }
}

// MLP stages 2 and 3 (h1 -> h2 -> [shift|scale] partials in red); returns the partial count KSe.
// `next_wf` describes the GEMM that follows the coupling step (prefetched behind the last barrier).
template <int TP, bool SAVE>
__device__ __forceinline__ int mlp_tail(const TileLayout& L, const float* __restrict__ lay, const fab_flow_desc& f,
                                        int k, const float* next_wf, int next_NT, bool next_ksplit, int next_KT2) {
    const TileBufs b = tile_bufs(L);
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    uint32_t* m2 = SAVE ? b.m2 + (size_t)k * L.MW : nullptr;
    float* h2 = b.h2;
    PF_E(mma_prefetch(reinterpret_cast<const float4*>(lay + f.o_w3), L.P8 / 8, true, L.W16 / 16);)
    mma_gemm_wide<TP>(b.h1, L.W16 / 16, reinterpret_cast<const float4*>(lay + f.o_w2), L.NTH,
                      lay + f.o_b2, [&](int nt, const float (&c)[4]) {
                          hidden_fwd<TP, SAVE>(h2, m2, nt, g, t, c);
                      });
    PF_L(mma_prefetch(reinterpret_cast<const float4*>(lay + f.o_w3), L.P8 / 8, true, L.W16 / 16);)
    __syncthreads();
    prof_mark(5);
    PF_E(if (next_wf) mma_prefetch(reinterpret_cast<const float4*>(next_wf), next_NT, next_ksplit, next_KT2);)
    const int KSe = mma_gemm_ksplit<TP>(b.h2, L.W16 / 16, reinterpret_cast<const float4*>(lay + f.o_w3),
                                        L.P8 / 8, b.red);
    PF_L(if (next_wf) mma_prefetch(reinterpret_cast<const float4*>(next_wf), next_NT, next_ksplit, next_KT2);)
    __syncthreads();
    prof_mark(7);
    return KSe;
}

// Per-particle sums over a row index j in a CANONICAL order.  The log-density of a particle is
//   -d/2 log 2pi - sum_j gauss_j - sum_j (sum_layers scale_kj) + sum_k sum(log_S_k):
// rows [0, d) of scl hold the Gaussian terms, rows [d, d + d2) the coupling scales ACCUMULATED
// over the layers by the thread that owns (j, slot) in every layer (no per-layer reduction);
// one reduction at the end: half-warp (2*warp + lane/16) owns particle p, lane j%16 adds its
// rows in order, a fixed xor tree joins the 16 lanes.  The order depends neither on the thread
// mapping of the producers nor on the tile configuration, so a particle's result does not depend
// on the batch it is evaluated in (sharded == single-device runs bit for bit,
// tests/multi_gpu_check.py).  Returns the sum in every lane of the half-warp; `p` is that
// half-warp's particle (may be >= TP).
template <int TP>
__device__ __noinline__ float slot_column_sum(const float* scl, int nj, int& p) {
    const int lane = threadIdx.x & 31;
    p = 2 * (threadIdx.x >> 5) + (lane >> 4);
    float s = 0.f;
    if (p < TP)
        for (int j = lane & 15; j < nj; j += 16) s += scl[j * TP + p];
    s += __shfl_xor_sync(FAB_FULL, s, 8);
    s += __shfl_xor_sync(FAB_FULL, s, 4);
    s += __shfl_xor_sync(FAB_FULL, s, 2);
    s += __shfl_xor_sync(FAB_FULL, s, 1);
    return s;
}

// ---- activation tape for the parameter gradient (param_grad.cuh) -------------
// With TAPE the evaluation writes, per coupling layer and particle, one row
//   [z_in (d) | 1 | h1 (W) | 1 | h2 (W) | 1 | gparam (2 d2) | gh2 (W) | gh1 (W) | gv (d)]      (segments padded to 4 floats)
// to global memory: the operands of the weight-gradient GEMMs (batch = the contraction dimension;
// the ones columns turn the bias gradients into GEMM rows / columns).  Layer k's rows start at
// base + k n RS, the final latent z_K at base + K n RS with row stride DP.
struct FabTape {
    float* base;
    long long n;
    int RS, o_z, o_h1, o_h2, o_gpar, o_gh2, o_gh1, o_gv;
    __host__ __device__ float* layer(int k) const { return base + (size_t)k * n * RS; }
};
__host__ inline FabTape make_tape(const fab_flow_desc& f, float* base, long long n) {
    FabTape t{};
    t.base = base; t.n = n;
    int o = 0;
    auto take = [&](int c) { int r = o; o += fab_round4(c); return r; };
    t.o_z = take(f.dim + 1); t.o_h1 = take(f.width + 1); t.o_h2 = take(f.width + 1);
    t.o_gpar = take(2 * f.d2); t.o_gh2 = take(f.width); t.o_gh1 = take(f.width); t.o_gv = take(f.dim);
    t.RS = o;
    return t;
}
// operand buffer (k-major, slot stride S) -> rows of the tape; with ONE a trailing 1 per row
template <int TP, bool ONE>
__device__ __forceinline__ void tape_dump(float* dst, int RS, const float* src, int nfeat, int np) {
    constexpr int S = ActL<TP>::S;
    const int nf1 = nfeat + (ONE ? 1 : 0);
    for (int e = threadIdx.x; e < nf1 * np; e += FAB_NT) {
        const int p = e / nf1, j = e - p * nf1;
        dst[(size_t)p * RS + j] = (ONE && j == nfeat) ? 1.0f : src[(size_t)j * S + p];
    }
}

// x in zsel(L, cur) -> z in zsel(L, cur') (cur' returned), log q in lq_out[p] (shared, p < TP).  With
// SAVE the per-layer (y2, exp(-scale), ReLU masks) needed by flow_backward are kept and b.gs is
// set to d log N(z) / dz.
template <int TP, bool SAVE, bool TAPE = false>
__device__ int flow_inverse(const TileLayout& L, const fab_flow_desc& f,
                            const float* __restrict__ blob, int cur, float* lq_out,
                            FabTape tape = FabTape{}, long long row0 = 0, int np = 0) {
    const TileBufs b = tile_bufs(L);
    constexpr int S = ActL<TP>::S;
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    float* sacc = b.scl + (size_t)L.d * TP;                 // accumulated scales, rows [d, d + d2)
    for (int e = threadIdx.x; e < L.d2 * TP; e += FAB_NT) sacc[e] = 0.f;   // (each e has one owner)
    const int NT1 = L.D8 / 8 + L.NTH, NTV = L.D8 / 8;
    for (int k = L.K - 1; k >= 0; --k) {
        const float* lay = blob + f.off_layers + (size_t)k * f.layer_stride;
        // [v | h1pre] = z @ [Wmix | Wmix[:, :d1] W1^T] + [0 | b1]
        float* zn = zsel(L, cur ^ 1);
        float* h1 = b.h1;
        uint32_t* m1 = SAVE ? b.m1 + (size_t)k * L.MW : nullptr;
        auto epi1 = [&](int nt, const float (&c)[4]) {
            if (nt < NTV) frag_store<TP>(zn, 8 * nt, g, t, c[0], c[1], c[2], c[3]);
            else hidden_fwd<TP, SAVE>(h1, m1, nt - NTV, g, t, c);
        };
        PF_E(mma_prefetch(reinterpret_cast<const float4*>(lay + f.o_w2), L.NTH, false, L.W16 / 16);)
        // 41..48 tiles (config 2: 4 + 40): six per warp in ONE pass instead of a nearly empty second
        if (NT1 > 8 * FAB_NTW && NT1 <= 8 * (FAB_NTW + 1))
            mma_gemm_wide_n<TP, FAB_NTW + 1>(zsel(L, cur), L.D16 / 16,
                                             reinterpret_cast<const float4*>(lay + f.o_mw1), NT1, lay + f.o_b1, epi1);
        else
            mma_gemm_wide<TP>(zsel(L, cur), L.D16 / 16, reinterpret_cast<const float4*>(lay + f.o_mw1), NT1,
                              lay + f.o_b1, epi1);
        PF_L(mma_prefetch(reinterpret_cast<const float4*>(lay + f.o_w2), L.NTH, false, L.W16 / 16);)
        __syncthreads();
        prof_mark(3);
        const int KSe = mlp_tail<TP, SAVE>(L, lay, f, k,
                                           k > 0 ? lay - f.layer_stride + f.o_mw1 : nullptr, NT1, false, L.D16 / 16);
        if (TAPE) {      // operands of dM1 / dWmix (z_in), dW2 (h1), dW3 (h2); all three stay valid until here
            float* row = tape.layer(k) + (size_t)row0 * tape.RS;
            tape_dump<TP, true>(row + tape.o_z, tape.RS, zsel(L, cur), L.d, np);
            tape_dump<TP, true>(row + tape.o_h1, tape.RS, b.h1, f.width, np);
            tape_dump<TP, true>(row + tape.o_h2, tape.RS, b.h2, f.width, np);
        }
        // coupling inverse: y2 = (v2 - shift) * exp(-scale)
        {
            const float* b3 = lay + f.o_b3;
            for (int e = threadIdx.x; e < L.d2 * TP; e += FAB_NT) {
                const int j = e / TP, p = e - j * TP;
                const float shift = red_sum<TP>(b.red, KSe, L.P8, p, j) + __ldg(b3 + j);
                const float scale = red_sum<TP>(b.red, KSe, L.P8, p, L.d2 + j) + __ldg(b3 + L.d2 + j);
                const float es = expf(-scale);
                float* zp = zn + (size_t)(L.d1 + j) * S + p;
                const float y2 = (*zp - shift) * es;
                *zp = y2;
                sacc[e] += scale;
                if (SAVE) {
                    b.sy2[((size_t)k * L.d2 + j) * TP + p] = y2;
                    b.ses[((size_t)k * L.d2 + j) * TP + p] = es;
                }
            }
        }
        __syncthreads();
        prof_mark(8);
        cur ^= 1;
    }
    // base Gaussian: log N(z; loc, exp(log_scale)) and its z-gradient
    {
        const float* z = zsel(L, cur);
        if (TAPE) {
            constexpr int S_ = ActL<TP>::S;
            float* zf = tape.layer(L.K) + (size_t)row0 * L.DP;
            for (int e = threadIdx.x; e < L.d * np; e += FAB_NT) {
                const int p = e / L.d, j = e - p * L.d;
                zf[(size_t)p * L.DP + j] = z[(size_t)j * S_ + p];
            }
        }
        for (int e = threadIdx.x; e < L.d * TP; e += FAB_NT) {
            const int j = e / TP, p = e - j * TP;
            const float inv = b.inv[j];
            const float u = (z[(size_t)j * S + p] - b.loc[j]) * inv;
            b.scl[e] = b.lsc[j] + 0.5f * u * u;
            if (SAVE) b.gs[(size_t)j * S + p] = -u * inv;
        }
        __syncthreads();
        {
            int p;
            const float tot = slot_column_sum<TP>(b.scl, L.d + L.d2, p);
            if ((threadIdx.x & 15) == 0 && p < TP)
                lq_out[p] = b.logs[L.K] + (-0.5f * (float)L.d * 1.8378770664093453f - tot);
        }
    }
    __syncthreads();
    prof_mark(9);
    return cur;
}

// b.gs holds d log q / d z on entry (set by flow_inverse<SAVE=true>) and d log q / d x on exit.
template <int TP, bool TAPE = false>
__device__ void flow_backward(const TileLayout& L, const fab_flow_desc& f,
                              const float* __restrict__ blob,
                              FabTape tape = FabTape{}, long long row0 = 0, int np = 0) {
    const TileBufs b = tile_bufs(L);
    constexpr int S = ActL<TP>::S;
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    float* gs = b.gs;
    float* gv = b.h1 + (size_t)L.W16 * S;               // [gv] rows behind gh1
    for (int k = 0; k < L.K; ++k) {
        const float* lay = blob + f.off_layers + (size_t)k * f.layer_stride;
        // coupling backward: gv2 = g2*es, gshift = -gv2, gscale = -g2*y2 - 1
        for (int e = threadIdx.x; e < L.d2 * TP; e += FAB_NT) {
            const int j = e / TP, p = e - j * TP;
            const float es = b.ses[((size_t)k * L.d2 + j) * TP + p];
            const float y2 = b.sy2[((size_t)k * L.d2 + j) * TP + p];
            const float g2 = gs[(size_t)(L.d1 + j) * S + p];
            const float gv2 = g2 * es;
            b.par[(size_t)j * S + p] = -gv2;
            b.par[(size_t)(L.d2 + j) * S + p] = -g2 * y2 - 1.0f;
            gv[(size_t)(L.d1 + j) * S + p] = gv2;
        }
        for (int e = threadIdx.x; e < L.d1 * TP; e += FAB_NT) {
            const int j = e / TP, p = e - j * TP;
            gv[(size_t)j * S + p] = gs[(size_t)j * S + p];
        }
        __syncthreads();
        prof_mark(10);
        // gh2 = (gparam @ W3) * m2
        {
            float* h2 = b.h2;
            const uint32_t* m2 = b.m2 + (size_t)k * L.MW;
            PF_E(mma_prefetch(reinterpret_cast<const float4*>(lay + f.o_w2t), L.NTH, false, L.W16 / 16);)
            mma_gemm_wide<TP>(b.par, L.P16 / 16, reinterpret_cast<const float4*>(lay + f.o_w3t), L.NTH,
                              nullptr, [&](int nt, const float (&c)[4]) {
                                  hidden_bwd<TP>(h2, m2, nt, g, t, c);
                              });
        }
        PF_L(mma_prefetch(reinterpret_cast<const float4*>(lay + f.o_w2t), L.NTH, false, L.W16 / 16);)
        __syncthreads();
        prof_mark(11);
        // gh1 = (gh2 @ W2) * m1
        {
            float* h1 = b.h1;
            const uint32_t* m1 = b.m1 + (size_t)k * L.MW;
            PF_E(mma_prefetch(reinterpret_cast<const float4*>(lay + f.o_w1mt), L.D8 / 8, true, (L.W16 + L.D16) / 16);)
            mma_gemm_wide<TP>(b.h2, L.W16 / 16, reinterpret_cast<const float4*>(lay + f.o_w2t), L.NTH,
                              nullptr, [&](int nt, const float (&c)[4]) {
                                  hidden_bwd<TP>(h1, m1, nt, g, t, c);
                              });
        }
        PF_L(mma_prefetch(reinterpret_cast<const float4*>(lay + f.o_w1mt), L.D8 / 8, true, (L.W16 + L.D16) / 16);)
        __syncthreads();
        prof_mark(13);
        if (TAPE) {      // unit-seed gradients of this layer: gparam, gh2, gh1, gv (all still in place)
            float* row = tape.layer(k) + (size_t)row0 * tape.RS;
            tape_dump<TP, false>(row + tape.o_gpar, tape.RS, b.par, 2 * L.d2, np);
            tape_dump<TP, false>(row + tape.o_gh2, tape.RS, b.h2, f.width, np);
            tape_dump<TP, false>(row + tape.o_gh1, tape.RS, b.h1, f.width, np);
            tape_dump<TP, false>(row + tape.o_gv, tape.RS, gv, L.d, np);
        }
        PF_E(if (k + 1 < L.K)
            mma_prefetch(reinterpret_cast<const float4*>(lay + f.layer_stride + f.o_w3t), L.NTH, false, L.P16 / 16);)
        // g_u = [gh1 | gv] @ [W1 Wmix[:, :d1]^T ; Wmix^T]
        const int KSe = mma_gemm_ksplit<TP>(b.h1, (L.W16 + L.D16) / 16,
                                            reinterpret_cast<const float4*>(lay + f.o_w1mt), L.D8 / 8, b.red);
        PF_L(if (k + 1 < L.K)
            mma_prefetch(reinterpret_cast<const float4*>(lay + f.layer_stride + f.o_w3t), L.NTH, false, L.P16 / 16);)
        __syncthreads();
        prof_mark(15);
        for (int e = threadIdx.x; e < L.d * TP; e += FAB_NT) {
            const int j = e / TP, p = e - j * TP;
            gs[(size_t)j * S + p] = red_sum<TP>(b.red, KSe, L.D8, p, j);
        }
        __syncthreads();
        prof_mark(16);
    }
}

// eps in zsel(L, cur) -> x in zsel(L, cur) (same buffer), forward-pass log q in lq_out[p].
template <int TP>
__device__ void flow_sample(const TileLayout& L, const fab_flow_desc& f,
                            const float* __restrict__ blob, int cur, float* lq_out) {
    const TileBufs b = tile_bufs(L);
    constexpr int S = ActL<TP>::S;
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    float* z = zsel(L, cur);
    {   // base: z = loc + exp(log_scale)*eps ; log p0 = -d/2 log 2pi - sum(log_scale + eps^2/2)
        for (int e = threadIdx.x; e < L.d * TP; e += FAB_NT) {
            const int j = e / TP, p = e - j * TP;
            const float ls = b.lsc[j];
            const float ev = z[(size_t)j * S + p];
            b.scl[e] = ls + 0.5f * ev * ev;
            z[(size_t)j * S + p] = b.loc[j] + expf(ls) * ev;
        }
        __syncthreads();
    }
    float* sacc = b.scl + (size_t)L.d * TP;                 // accumulated scales, rows [d, d + d2)
    for (int e = threadIdx.x; e < L.d2 * TP; e += FAB_NT) sacc[e] = 0.f;
    __syncthreads();            // (a flow without coupling layers goes straight to the final sum)
    for (int k = 0; k < L.K; ++k) {
        const float* lay = blob + f.off_layers + (size_t)k * f.layer_stride;
        // h1 = relu(z1 @ W1^T + b1): rows >= d1 of the operand meet zero weight rows
        {
            float* h1 = b.h1;
            mma_gemm_wide<TP>(z, L.D1K / 16, reinterpret_cast<const float4*>(lay + f.o_w1), L.NTH,
                              lay + f.o_b1s, [&](int nt, const float (&c)[4]) {
                                  hidden_fwd<TP, false>(h1, nullptr, nt, g, t, c);
                              });
        }
        mma_prefetch(reinterpret_cast<const float4*>(lay + f.o_w2), L.NTH, false, L.W16 / 16);
        __syncthreads();
        const int KSe = mlp_tail<TP, false>(L, lay, f, k, lay + f.o_mix_inv, L.D8 / 8, true, L.D16 / 16);
        {
            const float* b3 = lay + f.o_b3;
            for (int e = threadIdx.x; e < L.d2 * TP; e += FAB_NT) {
                const int j = e / TP, p = e - j * TP;
                const float shift = red_sum<TP>(b.red, KSe, L.P8, p, j) + __ldg(b3 + j);
                const float scale = red_sum<TP>(b.red, KSe, L.P8, p, L.d2 + j) + __ldg(b3 + L.d2 + j);
                float* zp = z + (size_t)(L.d1 + j) * S + p;
                *zp = *zp * expf(scale) + shift;
                sacc[e] += scale;               // log q -= sum(scale) - sum(log_S), reduced at the end
            }
        }
        __syncthreads();
        // u' = [v1,y2] @ Wmix^-1 + t   (t: the folded ActNorm shift, zero without ActNorm)
        const int KS2 = mma_gemm_ksplit<TP>(z, L.D16 / 16, reinterpret_cast<const float4*>(lay + f.o_mix_inv),
                                            L.D8 / 8, b.red);
        if (k + 1 < L.K)
            mma_prefetch(reinterpret_cast<const float4*>(lay + f.layer_stride + f.o_w1), L.NTH, false, L.D1K / 16);
        __syncthreads();
        for (int e = threadIdx.x; e < L.d * TP; e += FAB_NT) {
            const int j = e / TP, p = e - j * TP;
            z[(size_t)j * S + p] = red_sum<TP>(b.red, KS2, L.D8, p, j) + __ldg(lay + f.o_tmix + j);
        }
        __syncthreads();
    }
    {
        int p;
        const float tot = slot_column_sum<TP>(b.scl, L.d + L.d2, p);
        if ((threadIdx.x & 15) == 0 && p < TP)
            lq_out[p] = b.logs[L.K] + (-0.5f * (float)L.d * 1.8378770664093453f - tot);
    }
    __syncthreads();
}
