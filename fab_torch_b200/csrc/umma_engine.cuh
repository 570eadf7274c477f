// Row-tile engine: RealNVP log-density + input-gradient and the fused HMC outer step on the
// 5th-generation tensor cores (tcgen05.mma, accumulators in tensor memory, weights staged into
// shared memory by the TMA engine).  Used for batches that fill 128-row tiles; smaller batches
// and shapes this engine does not cover stay on the warp-level engine (tile_kernels.cuh).
//
// Decomposition (measured choices: profiles/r02_mb_umma_*.log)
//   * A CTA PAIR (cluster of 2, tcgen05 cta_group::2, instruction M = 128) carries 128 particles,
//     64 per CTA.  cta_group::2 runs the 64-row half tile of each SM at the full tensor rate
//     (8172 of 8192 FLOP/cycle/SM measured) and each CTA stages only half of every weight matrix.
//   * fp32-grade GEMMs from f16 tensor-core products: every operand is split as v*s = hi + lo
//     (f16 each, s a power of two -- per particle row for activations, per matrix for weights);
//     D = hi*hi + (lo*hi + hi*lo), the cross terms in their own accumulator, fp32 accumulation in
//     tensor memory.  The tensor core truncates each accumulation toward zero (measured): the mean
//     shrink, 1.67e-8 per accumulation of the hi*hi chain for mixed-sign data, is folded into the
//     un-scaling factor.
//   * The row scale of a hidden operand is known BEFORE its GEMM has finished: it comes from the
//     bound |h_n| <= ||h_in||_2 max_n ||W[:,n]||_2 + max|b| (the weight norms are packed with the
//     images, ||h_in||_2 is accumulated by the epilogue that produced h_in).  The bound is a few
//     powers of two above the true row maximum -- f16 hi/lo keeps 22 significant bits for every
//     element within 2^-9 of the largest representable scaled value and 2^-25 absolute below that,
//     far inside fp32 -- and it removes both the extra pass over the accumulators and the
//     dependency of one column block's epilogue on the other block.
//   * MMA / epilogue overlap, in place: a wide GEMM is issued as two column blocks of N/2 outputs, each
//     by its own issuer warp, with the k loop in halves (the operand of the NEXT GEMM is written half
//     by half into the buffer this GEMM still reads):
//         issuer 0:  wait A0 | (blk0, k-half 0 + bias) | wait A1 | (blk0, k-half 1)        -> D0, D1
//         issuer 1:  wait A1 | (blk1, k-half 0 + bias) -> D0 | (blk1, k-half 1)            -> D1
//     A0 / A1 = "operand half 0 / 1 is written" (16 warp arrivals each), D0 / D1 = accumulator barriers
//     (one commit per issuer each).  D0 says "block 0 is complete AND nobody reads the first half of
//     the A operand any more", so the epilogue of block 0 -- which overwrites exactly that half with
//     the next GEMM's operand -- can run under the tail of block 1's MMAs; the epilogue of block 1
//     runs under (blk0, k-half 0) of the NEXT GEMM, which issuer 0 starts as soon as block 0's
//     columns have been written (A0).  Block 1's accumulators are still being read by that second
//     epilogue, hence issuer 1 waits for A1.  All 8 compute warps work on each block.
//   * Thread (row r, lane half h, warp half u) owns, in every block, columns [u WQ/2, (u+1) WQ/2)
//     of lane half h, and columns 8 (2u+h) .. +7 of every d-wide vector for the whole kernel: the
//     running latent, the momentum and the gradients live in its registers; only MMA operands pass
//     through shared memory.
//   * Bias vectors ride in the GEMMs: each biased operand has one extra k-step whose first column
//     holds the row scale s ("ones" column after un-scaling) and meets the bias row of the weights.
//   * Saved-for-backward state: (y2, exp(-scale)) of every coupling layer in the 160 tensor-memory
//     columns the accumulators do not use; ReLU masks (80 bits per thread and GEMM) in a caller-
//     owned scratch buffer (L2 resident).
//
// Warp roles (352 threads): warps 0-7 compute (epilogues, coupling, leapfrog, accept); warps 8 and
// 10 of the leader CTA issue the MMAs -- TWO issuers, because walking the issue program costs one
// thread ~350 cycles per weight stage on top of the MMAs (~73 cycles per MMA measured) while these
// small MMAs (M 128 x N 160 x K 16 per pair) execute in 40: warp 8 owns column block 0 of the wide
// GEMMs and the narrow GEMMs, warp 10 column block 1 (separate accumulators, so the two instruction
// streams never order against each other); the weight stages of the two blocks alternate in the
// ring, each issuer walks its own program and frees its own stages, both commit the accumulator
// barriers.  Warp 8 of the peer CTA relays "my half of the weights has landed" to the leader;
// warp 9 lane 0 streams weight stages global -> shared with cp.async.bulk.
#pragma once
#include <cuda_fp16.h>
#include "umma.cuh"
#include "common.cuh"
#include "target_tile.cuh"

#define UE_ROWS 64
#define UE_THREADS 352
#define UE_CTHREADS 256
#define UE_STAGE_BYTES 30720      // six k-steps of a W x W column block (160 rows x 32 bytes each); 3 x 30 KB measured best (4 x 22.5 KB: 0.682 ms, 2 x 45 KB: 0.702)
#define UE_NSTAGE 3
#define UE_ACC_COLS 352          // accumulator columns; [352, 512) hold the saved coupling state
#define UE_DQ 8                  // columns of a d-wide vector per thread (d = 32)
#define UE_TRUNC_PER_ACC 1.67e-8f

// One operand-matrix type of a layer (host-built, include/fab_b200.h documents the blob).
struct UType {
    int KS;            // k-steps of 16 (incl. the bias step)
    int KSr;           // k-steps of real weight rows
    int kfirst;        // k-steps of the first segment in consumption order: k-half 0 (+ bias step); = KS for d-wide inputs
    int hin;           // 1: the A operand is a hidden activation (written in halves by the previous epilogue)
    int ksps;          // k-steps per pipeline stage
    int R;             // weight rows per CTA and 16-byte k-chunk ( = N: hi and lo rows of N/2 outputs)
    int Rb;            // rows of one block image per CTA and k-chunk: wide N/2, narrow N
    int nblk;          // column blocks: wide 2, narrow 1
    int nbh;           // accumulator columns per block: wide N/4, narrow N/2
    int wide;          // 1: two column blocks x three MMAs per k-step; 0: narrow "concat" form
    int a_off, a_lo;   // shared-memory byte offset of the A operand (hi plane), distance to the lo plane
    int dcol;          // first accumulator column
    uint32_t idesc_a, idesc_b;
    long long blob_off;    // byte offset inside a layer block (rank 0; rank 1 follows at KS*R*32)
    int Kreal, N;      // logical K (without bias step) and N
    int bias;          // 1: a bias row follows the Kreal weight rows
    long long plain_off, plain_bias_off;   // float offsets inside a layer block of the plain buffer
};

#define UE_SCAL 16               // floats per layer in the scalar table (see k_umma_scales)

// One entry of the MMA issue program of a layer (host-built by make_ulayout, the same for every layer):
// `count` consecutive k-steps of one weight stage whose logical k-steps are consecutive too.
enum : uint32_t {
    UOP_WAIT_A0 = 1u, UOP_WAIT_A1 = 2u,     // before: wait for the operand halves of the group
    UOP_NEWSTAGE = 4u,                      // before: the next weight stage (wait full[slot])
    UOP_WIDE = 8u,                          // three MMAs per k-step (else the narrow two-MMA form)
    UOP_ACC = 16u,                          // the first k-step accumulates (else it overwrites)
    UOP_FREE = 32u,                         // after: the stage is consumed (commit empty[slot], next slot)
    UOP_D0 = 64u, UOP_D1 = 128u             // after: commit dfull[0] / dfull[1] (D1 ends the group)
};
struct UOp {
    uint32_t a;        // A descriptor low word of the first k-step, relative to the shared-memory base >> 4 (LBO included)
    uint32_t b;        // B descriptor low word relative to the stage base >> 4 (LBO included)
    uint32_t a_lo;     // distance hi -> lo plane of the A operand >> 4
    uint32_t kstr;     // B step per k-step >> 4
    uint32_t d0, d1;   // accumulator columns: wide main / cross, narrow [hi*hi | hi*lo] / += lo*hi
    uint32_t nbh;      // wide: distance from the hi to the lo rows of the B block >> 4 (rows x 16 B)
    uint32_t idesc0, idesc1;
    uint32_t count, flags;
    uint32_t it_rel;   // weight stage of this entry, counted from the first stage of the layer's (forward / gradient) program
};
#define UE_MAX_OPS 40

struct ULayout {
    int d, W, K, WQ;
    UType t[7];
    long long blob_bytes, off_layers, layer_bytes;     // blob: [scalars][layer blocks]
    int o_loc, o_lsc, o_scal;                           // float offsets in the scalar block; scal[K][UE_SCAL]
    long long plain_floats, plain_layer_floats, plain_off_layers;
    long long plain_logs_off;                           // float offset of sum(log_S) inside a plain layer block
    int s_h, s_z, s_par, s_gv, s_ring, s_ex, s_ssq, s_bar, smem_bytes;
    int hplane, zplane, pplane;
    float dl[8];     // truncation compensation: [0] v columns of type 0, [1] h1pre of type 0, [2..7] types 1..6
    // issue programs of the two issuer warps (issuer g owns column block g of the wide GEMMs, issuer 0 the
    // narrow ones): ops[r][0, n_ops_fwd[r]) forward groups of a layer, [n_ops_fwd[r], n_ops[r]) gradient groups
    int n_ops_fwd[2], n_ops[2];
    int n_stages_fwd, n_stages_bwd;      // weight stages per layer in the forward / gradient program
    UOp ops[2][UE_MAX_OPS];
};

__host__ inline bool umma_supported(int d, int W, int K) { return d == 32 && W % 64 == 0 && W >= 64 && W <= 320 && K >= 1 && K <= 10; }

// logical k-step of the j-th k-step in consumption order: [k-half 0 | bias step | k-half 1]
__host__ __device__ __forceinline__ int ue_kstep(const UType& t, int j) {
    if (!t.hin) return j;
    const int half = t.KSr / 2;
    if (j < half) return j;
    if (t.bias) return j == half ? t.KSr : j - 1;
    return j;
}

__host__ inline ULayout make_ulayout(int d, int W, int K) {
    ULayout L{};
    L.d = d; L.W = W; L.K = K; L.WQ = W / 4;
    const int hplane = UE_ROWS * (W + 16) * 2, zplane = UE_ROWS * (d + 16) * 2, pplane = UE_ROWS * d * 2;
    L.hplane = hplane; L.zplane = zplane; L.pplane = pplane;
    int o = 0;
    L.s_h = o; o += 2 * hplane;
    L.s_z = o; o += 2 * zplane;
    L.s_par = o; o += 2 * pplane;
    L.s_gv = o; o += 2 * pplane;
    o = (o + 127) & ~127;
    L.s_ring = o; o += UE_NSTAGE * UE_STAGE_BYTES;
    L.s_ex = o; o += 4 * 4 * UE_ROWS * 4 * 2;     // exchange buffers: 4 rotating x [2 values][4 groups][64 rows]
    L.s_ssq = o; o += 2 * 4 * UE_ROWS * 4;        // per-row partial sums of squares: 2 alternating x [4 threads][64 rows]
    L.s_bar = o; o += (2 * UE_NSTAGE + 4) * 8 + 16;
    L.smem_bytes = o;
    auto set = [&](int i, int Kreal, int N, int bias, int wide, int hin, int a_off, int a_lo, int dcol) {
        UType& t = L.t[i];
        t.Kreal = Kreal; t.N = N; t.bias = bias; t.wide = wide; t.hin = hin;
        t.KSr = Kreal / 16;
        t.KS = t.KSr + (bias ? 1 : 0);
        t.kfirst = hin ? t.KSr / 2 + (bias ? 1 : 0) : t.KS;
        t.R = N;
        t.nblk = wide ? 2 : 1;
        t.Rb = wide ? N / 2 : N;
        t.nbh = wide ? N / 4 : N / 2;
        t.ksps = UE_STAGE_BYTES / (t.Rb * 32);
        if (t.ksps > t.KS) t.ksps = t.KS;
        t.a_off = a_off; t.a_lo = a_lo; t.dcol = dcol;
        t.idesc_a = umma::instr_desc(umma::FMT_F16, 128, wide ? N / 2 : 2 * N);
        t.idesc_b = umma::instr_desc(umma::FMT_F16, 128, wide ? N / 2 : N);
    };
    set(0, d, d + W, 1, 1, 0, L.s_z, zplane, 0);        // z -> [v | h1pre]
    set(1, W, W, 1, 1, 1, L.s_h, hplane, 0);            // h1 -> h2pre
    set(2, W, d, 1, 0, 1, L.s_h, hplane, 0);            // h2 -> [shift | scale]
    set(3, d, W, 0, 1, 0, L.s_par, pplane, 0);          // gparam -> gh2
    set(4, W, W, 0, 1, 1, L.s_h, hplane, 0);            // gh2 -> gh1
    set(5, W, d, 0, 0, 1, L.s_h, hplane, 0);            // gh1 -> g (first part)
    set(6, d, d, 0, 0, 0, L.s_gv, pplane, d);           // gv  -> g (second part)
    long long bo = 0, po = 0;
    for (int i = 0; i < 7; ++i) {
        UType& t = L.t[i];
        t.blob_off = bo; bo += 2LL * t.KS * t.R * 32;
        t.plain_off = po; po += (long long)t.Kreal * t.N;
        t.plain_bias_off = po; if (t.bias) po += t.N;
    }
    L.plain_logs_off = po; po += 4;
    L.layer_bytes = bo; L.plain_layer_floats = po;
    L.o_loc = 0; L.o_lsc = d; L.o_scal = 2 * d;
    L.off_layers = ((2 * d + UE_SCAL * K) * 4 + 127) & ~127;
    L.blob_bytes = L.off_layers + L.layer_bytes * K;
    L.plain_off_layers = 2 * d;
    L.plain_floats = L.plain_off_layers + L.plain_layer_floats * K;
    // mean relative shrink of a chain of n truncating accumulations at ~final magnitude: measured
    // 1.67e-8 n for long mixed-sign chains (profiles/r02_mb_umma_layout_rounding_split.log); the
    // v = z @ Wmix columns see no bias row, i.e. one accumulation less
    L.dl[0] = 4.5e-8f;     // calibrated on log q of a 10-layer flow (one FMA rounds up for ~60 % of the mantissas)
    L.dl[1] = UE_TRUNC_PER_ACC * (float)L.t[0].KS;
    for (int i = 1; i < 7; ++i) L.dl[1 + i] = UE_TRUNC_PER_ACC * (float)L.t[i].KS;
    // ---- MMA issue programs.  Stream order of the weight stages (ue_for_each_stage): per operand type
    // stage by stage, for a wide type block 0's stage then block 1's; issuer g consumes block g's. ----
    int no[2] = {0, 0}, stage = 0;
    auto group = [&](int t0, int t1) {
        bool first[2] = {true, true}, a1[2] = {false, false};
        const bool wide_group = L.t[t0].wide != 0;
        for (int ti = t0; ti <= t1; ++ti) {
            const UType& t = L.t[ti];
            // boundaries inside the consumption order where a run must end
            const int kb0 = t.hin && t.bias ? t.KSr / 2 : t.KS, kb1 = t.hin && t.bias ? kb0 + 1 : t.KS;
            for (int j0 = 0; j0 < t.KS; j0 += t.ksps)
                for (int g = 0; g < t.nblk; ++g) {
                    const int r = g;                           // the issuer of this block
                    const int j1 = j0 + t.ksps < t.KS ? j0 + t.ksps : t.KS;
                    int ja = j0;
                    while (ja < j1) {
                        int jb = j1;
                        if (ja < kb0 && kb0 < jb) jb = kb0;
                        if (ja < kb1 && kb1 < jb) jb = kb1;
                        if (ja < t.kfirst && t.kfirst < jb) jb = t.kfirst;
                        UOp& o = L.ops[r][no[r]++];
                        o = UOp{};
                        o.a = (uint32_t)(t.a_off >> 4) + (uint32_t)ue_kstep(t, ja) * 128u + ((1024u >> 4) << 16);
                        o.b = (uint32_t)(ja - j0) * (uint32_t)t.Rb * 2u + (((uint32_t)t.Rb * 16u >> 4) << 16);
                        o.a_lo = (uint32_t)t.a_lo >> 4;
                        o.kstr = (uint32_t)t.Rb * 2u;
                        o.nbh = (uint32_t)t.nbh;
                        o.d0 = t.wide ? (uint32_t)(t.dcol + g * 2 * t.nbh) : (uint32_t)t.dcol;
                        o.d1 = o.d0 + (uint32_t)t.nbh;
                        o.idesc0 = t.idesc_a; o.idesc1 = t.wide ? t.idesc_a : t.idesc_b;
                        o.count = (uint32_t)(jb - ja);
                        o.it_rel = (uint32_t)stage;
                        uint32_t f = (t.wide ? UOP_WIDE : 0u) | (ja > 0 ? UOP_ACC : 0u);
                        // block 1 overwrites accumulators that the previous group's second epilogue reads: it waits
                        // for both operand halves before its first MMA; block 0 / narrow types wait for the second
                        // half only where they reach it
                        // (waiting for the second half implies the first: every warp arrives on A0 before A1)
                        if (first[r]) {
                            if (!t.hin || g == 1) { f |= UOP_WAIT_A1; a1[r] = true; }
                            else f |= UOP_WAIT_A0;
                            first[r] = false;
                        }
                        if (!a1[r] && ja >= t.kfirst) { f |= UOP_WAIT_A1; a1[r] = true; }
                        if (ja == j0) f |= UOP_NEWSTAGE;
                        if (jb == j1) f |= UOP_FREE;
                        // D0 (block 0 complete and the first operand half no longer read; both issuers commit it):
                        // issuer 0 after its last k-step, issuer 1 after (blk1, k-half 0 + bias) -- for a d-wide input
                        // after its only stage
                        if (t.wide && g == 0 && jb == t.KS) f |= UOP_D0;
                        if (t.wide && g == 1 && jb == (t.hin ? t.kfirst : t.KS)) f |= UOP_D0;
                        // D1 (group complete; both issuers commit it)
                        if (ti == t1 && jb == t.KS) f |= UOP_D1;
                        o.flags = f;
                        ja = jb;
                    }
                    ++stage;
                }
        }
        if (!wide_group) {
            // issuer 1 has no MMAs in a narrow group: an empty entry commits D1 for it -- AFTER the group's
            // operand barriers (they complete after the previous group's D1, so the two arrivals of one
            // D1 phase always come from the two issuers)
            UOp& o = L.ops[1][no[1]++];
            o = UOp{};
            o.flags = UOP_WAIT_A1 | UOP_D1;
        }
    };
    group(0, 0); group(1, 1); group(2, 2);
    L.n_ops_fwd[0] = no[0]; L.n_ops_fwd[1] = no[1]; L.n_stages_fwd = stage;
    stage = 0;
    group(3, 3); group(4, 4); group(5, 6);
    L.n_ops[0] = no[0]; L.n_ops[1] = no[1]; L.n_stages_bwd = stage;
    return L;
}

// ---------------------------------------------------------------------------------------------
// weight packing: plain fp32 matrices M[k][n] (+ bias rows) -> per-rank f16 hi/lo stage images
// ---------------------------------------------------------------------------------------------
// one block per (layer, type): s_w = 2^e with max|M|,|bias| * s_w in [2^13, 2^14).  Scalar table of a
// layer (UE_SCAL floats): [0..6] 1 / s_w of the seven types, [7] sum(log_S), [8..11] the column-norm
// bound max_n ||M[:,n]||_2 of the wide types 0 (h1pre columns only), 1, 3, 4 (the a-priori operand
// scales of the hidden activations come from it), [12..13] max_n |bias_n| of types 0 and 1.
__global__ void k_umma_scales(ULayout L, const float* __restrict__ plain, float* __restrict__ blob_f) {
    const int layer = blockIdx.x / 7, ti = blockIdx.x % 7;
    const UType& t = L.t[ti];
    const float* M = plain + L.plain_off_layers + (size_t)layer * L.plain_layer_floats + t.plain_off;
    const long long cnt = (long long)t.Kreal * t.N + (t.bias ? t.N : 0);
    float m = 0.f;
    for (long long i = threadIdx.x; i < cnt; i += blockDim.x) m = fmaxf(m, fabsf(M[i]));
    // column norms (threads walk the columns: coalesced rows) and the largest bias
    float cn = 0.f, bm = 0.f;
    if (t.wide) {
        const int n0 = ti == 0 ? L.d : 0;
        for (int n = n0 + threadIdx.x; n < t.N; n += blockDim.x) {
            float ss = 0.f;
            for (int k = 0; k < t.Kreal; ++k) { const float w = M[(size_t)k * t.N + n]; ss = fmaf(w, w, ss); }
            cn = fmaxf(cn, ss);
            if (t.bias) bm = fmaxf(bm, fabsf(M[(size_t)t.Kreal * t.N + n]));
        }
    }
    __shared__ float red[3][32];
    m = warp_max(m); cn = warp_max(cn); bm = warp_max(bm);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = m; red[1][threadIdx.x >> 5] = cn; red[2][threadIdx.x >> 5] = bm; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { m = fmaxf(m, red[0][w]); cn = fmaxf(cn, red[1][w]); bm = fmaxf(bm, red[2][w]); }
        int e = 0;
        if (m > 0.f && m < CUDART_INF_F) { frexpf(m, &e); e = 14 - e; }
        if (e > 30) e = 30;
        if (e < -30) e = -30;
        float* scal = blob_f + L.o_scal + (size_t)layer * UE_SCAL;
        scal[ti] = ldexpf(1.f, -e);
        if (ti == 0) scal[7] = plain[L.plain_off_layers + (size_t)layer * L.plain_layer_floats + L.plain_logs_off];
        if (t.wide) {
            const int wi = ti == 0 ? 0 : ti == 1 ? 1 : ti == 3 ? 2 : 3;
            scal[8 + wi] = sqrtf(cn) * 1.0001f;        // (upper bound: fp32 rounding of the sum)
            if (t.bias) scal[12 + wi] = bm;
        }
    }
    if (blockIdx.x == 0)
        for (int j = threadIdx.x; j < 2 * L.d; j += blockDim.x) blob_f[j] = plain[j];
}

// one thread per 16-byte chunk of the images: grid.x covers (rank, block, k-step in consumption
// order, chunk of the k-step, row).  Image of (type, rank): for each column block the k-step slabs
// [2 chunks][Rb rows][16 bytes] in the order the MMA issuer consumes them (ue_kstep).
__global__ void k_umma_pack(ULayout L, const float* __restrict__ plain, uint8_t* __restrict__ blob) {
    const int layer = blockIdx.y, ti = blockIdx.z;
    const UType& t = L.t[ti];
    const long long per_rank = (long long)t.KS * 2 * t.R;          // chunks x rows
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 2 * per_rank) return;
    const int rank = (int)(idx / per_rank);
    long long rem = idx % per_rank;
    const long long per_blk = (long long)t.KS * 2 * t.Rb;
    const int g = (int)(rem / per_blk);
    rem %= per_blk;
    const int slab = (int)(rem / t.Rb), row = (int)(rem % t.Rb);   // slab = 2 j + chunk-of-k-step
    const int c = 2 * ue_kstep(t, slab >> 1) + (slab & 1);         // logical 16-byte k-chunk
    // row -> (part, output column n)
    const int part = row / t.nbh, i = row % t.nbh;
    int n;
    if (t.wide) {
        const int q = 2 * g + rank;
        if (ti == 0) n = i < L.WQ ? L.d + q * L.WQ + i : q * UE_DQ + (i - L.WQ);    // [h1pre cols | v cols]
        else n = q * L.WQ + i;
    } else {
        n = i < UE_DQ ? rank * UE_DQ + i : L.d / 2 + rank * UE_DQ + (i - UE_DQ);
    }
    const float* blob_f = reinterpret_cast<const float*>(blob);
    // s_w from the scale kernel's output scal = 1 / s_w
    const float sc = blob_f[L.o_scal + (size_t)layer * UE_SCAL + ti];
    const float sw = 1.f / sc;             // both powers of two
    const float* M = plain + L.plain_off_layers + (size_t)layer * L.plain_layer_floats + t.plain_off;
    const float* B = plain + L.plain_off_layers + (size_t)layer * L.plain_layer_floats + t.plain_bias_off;
    __half out[8];
    for (int j = 0; j < 8; ++j) {
        const int k = 8 * c + j;
        float v = 0.f;
        if (k < t.Kreal) v = M[(size_t)k * t.N + n];
        else if (k == t.Kreal && t.bias) v = B[n];
        v *= sw;
        const __half hi = __float2half_rn(v);
        out[j] = part == 0 ? hi : __float2half_rn(v - __half2float(hi));
    }
    uint8_t* dst = blob + L.off_layers + (size_t)layer * L.layer_bytes + t.blob_off +
                   (size_t)rank * t.KS * t.R * 32 + (size_t)g * t.KS * t.Rb * 32 + ((size_t)slab * t.Rb + row) * 16;
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(out);
}

// ---------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------
extern __shared__ __align__(1024) uint8_t ue_smem[];

// -DUE_PROF: cycle counters of CTA 0 (compute warp 0 lane 0, the MMA issuer, the producer);
// read back with fab_umma_prof_read.  0: compute waits for D0, 1: compute waits for D1,
// 2: issuer waits for operands (A0 / A1), 3: issuer waits for weight stages, 4: issuer issues,
// 5: producer waits for free slots, 6: compute barrier/exchange, 7: flow evaluations (compute warp 0),
// 8..13: compute waits (D0 + D1) per GEMM group G1, G2, G3, G3T, G2T, G1T
#ifdef UE_PROF
__device__ unsigned long long g_ue_prof[16];
#define UE_T0() const long long ue_t0_ = clock64()
#define UE_ACC(id, cond) if (cond) atomicAdd(&g_ue_prof[id], (unsigned long long)(clock64() - ue_t0_))
#else
#define UE_T0()
#define UE_ACC(id, cond)
#endif

// full[s]: stage s has landed -- in the leader CTA it counts two arrivals: its own producer's
// expect_tx and the peer's relay (so the MMA issuer polls ONE barrier per stage: a try_wait costs
// ~100 cycles of the single issuing thread even when the phase is already complete).
// aready[h] (leader CTA, 16 warp arrivals per GEMM group): half h of the group's A operand is written
// (A0: first half of the k-chunks + the bias chunk; operands that are not written in halves arrive on both).
// dfull[0]: accumulator block 0 complete and the first operand half no longer read; dfull[1]: group complete.
struct UBars { uint64_t *full, *empty, *dfull, *aready; };
__device__ __forceinline__ UBars ue_bars(const ULayout& L) {
    uint64_t* b = reinterpret_cast<uint64_t*>(ue_smem + L.s_bar);
    return {b, b + UE_NSTAGE, b + 2 * UE_NSTAGE, b + 2 * UE_NSTAGE + 2};
}

// The GEMM groups of one flow evaluation, in execution order.  f(layer, first type, last type).
template <class F>
__device__ __forceinline__ void ue_for_each_group(const ULayout& L, bool grad, F f) {
    for (int k = L.K - 1; k >= 0; --k) { f(k, 0, 0); f(k, 1, 1); f(k, 2, 2); }
    if (grad) for (int k = 0; k < L.K; ++k) { f(k, 3, 3); f(k, 4, 4); f(k, 5, 6); }
}
// The weight stages of `n_evals` evaluations in stream order (per type stage by stage, block 0's stage then
// block 1's): f(type, layer, block, first k-step
// (consumption index), k-steps in the stage, running stage index)
template <class F>
__device__ __forceinline__ void ue_for_each_stage(const ULayout& L, int n_evals, bool grad, F f) {
    uint32_t it = 0;
    for (int ev = 0; ev < n_evals; ++ev)
        ue_for_each_group(L, grad, [&](int k, int t0, int t1) {
            for (int ti = t0; ti <= t1; ++ti) {
                const UType& t = L.t[ti];
                for (int j0 = 0; j0 < t.KS; j0 += t.ksps)
                    for (int g = 0; g < t.nblk; ++g) { f(t, k, g, j0, min(t.ksps, t.KS - j0), it); ++it; }
            }
        });
}

// warp 9 (warp-uniform loop, one elected lane issues): stream every weight stage of `n_evals` flow evaluations into the ring
__device__ void ue_producer(const ULayout& L, const uint8_t* __restrict__ blob, uint32_t rank, int n_evals, bool grad) {
    const UBars B = ue_bars(L);
    const uint64_t pol = umma::l2_evict_last_policy();
    ue_for_each_stage(L, n_evals, grad, [&](const UType& t, int k, int g, int j0, int cnt, uint32_t it) {
        const uint8_t* src = blob + L.off_layers + (size_t)k * L.layer_bytes + t.blob_off + (size_t)rank * t.KS * t.R * 32 +
                             ((size_t)g * t.KS + j0) * t.Rb * 32;
        const uint32_t slot = it % UE_NSTAGE, par = (it / UE_NSTAGE) & 1;
        { UE_T0(); umma::mbar_wait(B.empty + slot, par ^ 1); UE_ACC(5, blockIdx.x == 0 && (threadIdx.x & 31) == 0); }
        const uint32_t bytes = (uint32_t)cnt * t.Rb * 32;
        if (umma::elect_one()) {
            umma::mbar_expect_tx(B.full + slot, bytes);
            umma::bulk_g2s(ue_smem + L.s_ring + slot * UE_STAGE_BYTES, src, bytes, B.full + slot, pol);
        }
        __syncwarp();
    });
}

// warp 8 of the peer CTA: tell the leader when this CTA's half of a stage has landed
__device__ void ue_relay(const ULayout& L, int n_evals, bool grad) {
    const UBars B = ue_bars(L);
    ue_for_each_stage(L, n_evals, grad, [&](const UType&, int, int, int, int, uint32_t it) {
        const uint32_t slot = it % UE_NSTAGE, par = (it / UE_NSTAGE) & 1;
        umma::mbar_wait(B.full + slot, par);
        if (umma::elect_one()) umma::mbar_arrive_remote(B.full + slot, 0);
        __syncwarp();
    });
}

// warp 8 of the leader CTA: issue the MMAs of the pair by running the layer's issue program
// (ULayout::ops) -- a small table-driven loop: the issuer is a single thread's instruction stream and
// every instruction between two tcgen05.mma counts (the unrolled control flow of an earlier version
// was 16 k SASS lines and stalled on instruction fetch).  Every lane executes the loop, the MMAs and
// commits are predicated on one elected lane, so descriptors and addresses stay in uniform registers.
template <int ROLE>
__device__ void ue_mma(const ULayout& L, uint32_t tmem, int n_evals, bool grad) {
    const UBars B = ue_bars(L);
    const uint32_t sbase4 = umma::smem_u32(ue_smem) >> 4, ring4 = (umma::smem_u32(ue_smem) + L.s_ring) >> 4;
    const uint32_t dhi = (128u >> 4) | (1u << 14);        // descriptor high word: SBO = 128 B, version 1
    auto desc = [&](uint32_t lo) { return ((uint64_t)dhi << 32) | lo; };
    const uint32_t lead = umma::elect_one() ? 1u : 0u;      // the lane whose MMAs and commits are issued
    uint32_t it_base = 0, gi = 0, sb4 = 0, slot = 0;
    auto run = [&](int e0, int e1) {
        for (int e = e0; e < e1; ++e) {
            const UOp& o = L.ops[ROLE][e];
            const uint32_t f = o.flags;
            // the entry's fields are read BEFORE the waits: the constant-bank latency then overlaps the wait
            // instead of sitting between the barrier and the first MMA of every round trip
            const uint32_t o_a = o.a, o_b = o.b, al = o.a_lo, kstr = o.kstr, o_d0 = o.d0, o_d1 = o.d1, id0 = o.idesc0,
                           id1 = o.idesc1, nbh = o.nbh, o_count = o.count, o_it = o.it_rel;
            // the weight stage first (it is prefetched: the ~80-cycle poll of a complete barrier then
            // runs while the operands are still being written), then the operand barriers
            if (f & UOP_NEWSTAGE) {
                const uint32_t it = it_base + o_it;
                slot = it % UE_NSTAGE;
                UE_T0(); umma::mbar_wait_cluster(B.full + slot, (it / UE_NSTAGE) & 1); UE_ACC(3, blockIdx.x == 0 && (threadIdx.x & 31) == 0);
                sb4 = ring4 + slot * (UE_STAGE_BYTES >> 4);
            }
            if (f & UOP_WAIT_A0) {
                UE_T0(); umma::mbar_wait_cluster(B.aready, gi & 1); UE_ACC(2, blockIdx.x == 0 && (threadIdx.x & 31) == 0);
            }
            if (f & UOP_WAIT_A1) {
                UE_T0(); umma::mbar_wait_cluster(B.aready + 1, gi & 1); UE_ACC(2, blockIdx.x == 0 && (threadIdx.x & 31) == 0);
            }
            if (f & (UOP_WAIT_A0 | UOP_WAIT_A1 | UOP_NEWSTAGE)) umma::tc_fence_after();
            {
                UE_T0();
                uint32_t ah = sbase4 + o_a, bk = sb4 + o_b, acc = (f & UOP_ACC) ? 1u : 0u;
                const uint32_t d0 = tmem + o_d0, d1 = tmem + o_d1;
#ifdef UE_X_NOMMA      // (timing experiments of profiles/r02_rowtile_experiments.md: the kernel without its MMAs)
                if (false) {
#else
                if (f & UOP_WIDE) {
#endif
                    for (uint32_t n = o_count; n > 0; --n) {
                        umma::mma_ss2_pred(d1, desc(ah + al), desc(bk), id0, acc, lead);           // cross  = lo * hi
                        umma::mma_ss2_pred(d1, desc(ah), desc(bk + nbh), id0, 1u, lead);           // cross += hi * lo
                        umma::mma_ss2_pred(d0, desc(ah), desc(bk), id0, acc, lead);                // main   = hi * hi
                        ah += 128; bk += kstr; acc = 1u;
                    }
#ifdef UE_X_NOMMA
                } else if (false) {
#else
                } else {
#endif
                    for (uint32_t n = o_count; n > 0; --n) {
                        umma::mma_ss2_pred(d0, desc(ah), desc(bk), id0, acc, lead);                // [hi*hi | hi*lo]
                        umma::mma_ss2_pred(d1, desc(ah + al), desc(bk), id1, 1u, lead);            // += lo*hi
                        ah += 128; bk += kstr; acc = 1u;
                    }
                }
                UE_ACC(4, blockIdx.x == 0 && (threadIdx.x & 31) == 0);
            }
            if (f & UOP_FREE) { if (lead) umma::mma_commit<2>(B.empty + slot, 3); }
            if (f & UOP_D0) { if (lead) umma::mma_commit<2>(B.dfull, 3); }
            if (f & UOP_D1) { if (lead) umma::mma_commit<2>(B.dfull + 1, 3); ++gi; }
            __syncwarp();
        }
    };
    for (int ev = 0; ev < n_evals; ++ev) {
        for (int k = 0; k < L.K; ++k) { run(0, L.n_ops_fwd[ROLE]); it_base += L.n_stages_fwd; }
        if (grad) for (int k = 0; k < L.K; ++k) { run(L.n_ops_fwd[ROLE], L.n_ops[ROLE]); it_base += L.n_stages_bwd; }
    }
}

// ---- compute warps --------------------------------------------------------------------------------
struct UCw {
    int r, q, g, h;             // row, d-column group (2g + h), warp half, lane half
    uint32_t tl;                // tensor-memory lane base of the warp << 16
    uint32_t gi;                // GEMM groups consumed so far (phase of dfull[1])
    uint32_t n0;                // wide GEMM groups consumed so far (phase of dfull[0]: only wide groups commit it)
    uint32_t xb;                // rotating exchange buffer index
    uint32_t nq;                // sum-of-squares hand-offs so far (alternating buffer)
    uint32_t tmem;
    uint32_t rank;
};

__device__ __forceinline__ void ue_bar_compute() {
    UE_T0();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    UE_ACC(6, blockIdx.x == 0 && threadIdx.x == 0);
}

__device__ __forceinline__ float* ue_exbuf(const ULayout& L, UCw& c) {
    float* b = reinterpret_cast<float*>(ue_smem + L.s_ex) + (c.xb & 3) * (2 * 4 * UE_ROWS);
    ++c.xb;
    return b;
}
// max over the four threads of a row (two values at once)
__device__ __forceinline__ void ue_row_max2(const ULayout& L, UCw& c, float& a, float& b) {
    float* ex = ue_exbuf(L, c);
    ex[c.q * UE_ROWS + c.r] = a;
    ex[4 * UE_ROWS + c.q * UE_ROWS + c.r] = b;
    ue_bar_compute();
    a = fmaxf(fmaxf(ex[c.r], ex[UE_ROWS + c.r]), fmaxf(ex[2 * UE_ROWS + c.r], ex[3 * UE_ROWS + c.r]));
    const float* e2 = ex + 4 * UE_ROWS;
    b = fmaxf(fmaxf(e2[c.r], e2[UE_ROWS + c.r]), fmaxf(e2[2 * UE_ROWS + c.r], e2[3 * UE_ROWS + c.r]));
}
// sum over the four threads of a row in the canonical order q = 0, 1, 2, 3 (two values at once)
__device__ __forceinline__ void ue_row_sum2(const ULayout& L, UCw& c, float& a, float& b) {
    float* ex = ue_exbuf(L, c);
    ex[c.q * UE_ROWS + c.r] = a;
    ex[4 * UE_ROWS + c.q * UE_ROWS + c.r] = b;
    ue_bar_compute();
    a = ((ex[c.r] + ex[UE_ROWS + c.r]) + ex[2 * UE_ROWS + c.r]) + ex[3 * UE_ROWS + c.r];
    const float* e2 = ex + 4 * UE_ROWS;
    b = ((e2[c.r] + e2[UE_ROWS + c.r]) + e2[2 * UE_ROWS + c.r]) + e2[3 * UE_ROWS + c.r];
}
// Sum of squares of a hidden operand row, handed from the epilogue that wrote the operand to the
// epilogue of the GEMM that consumes it.  Every compute thread writes its part at the end of its
// epilogue; the reader first passes a CTA barrier of the 256 compute threads (so all four parts of
// its row are in place) -- this happens while the consuming GEMM's MMAs run, before the accumulator
// wait.  Two alternating buffers keep a fast thread's next hand-off away from a slow thread's read.
__device__ __forceinline__ void ue_ssq_put(const ULayout& L, UCw& c, float v) {
    reinterpret_cast<float*>(ue_smem + L.s_ssq)[(c.nq & 1) * (4 * UE_ROWS) + c.q * UE_ROWS + c.r] = v;
}
__device__ __forceinline__ float ue_ssq_get(const ULayout& L, UCw& c) {
    ue_bar_compute();
    const float* p = reinterpret_cast<const float*>(ue_smem + L.s_ssq) + (c.nq & 1) * (4 * UE_ROWS) + c.r;
    ++c.nq;
    return ((p[0] + p[UE_ROWS]) + p[2 * UE_ROWS]) + p[3 * UE_ROWS];
}
// power-of-two operand scale: m (the row maximum or an upper bound of it) -> [2^13, 2^14).  BIASED
// operands carry s itself as an f16 number (the "ones" column), so s <= 2^15 there (rows whose bound
// is below 2^-2 keep a smaller scaled maximum, still far above the f16 subnormals); a bound above
// 2^38 makes the ones column underflow, i.e. drops a bias that is < 1e-11 of the row.  Gradient
// operands are un-biased and take the full exponent range.
template <bool BIASED>
__device__ __forceinline__ void ue_scale_of(float m, float& s, float& inv_s) {
    int es = 0;
    if (m > 0.f) {
        const int E = (int)((__float_as_uint(m) >> 23) & 0xffu);      // m in [2^(E-127), 2^(E-126))
        es = 13 - (E - 127);
        if (BIASED && es > 15) es = 15;
        if (es > 120) es = 120;
        if (es < -120) es = -120;
    }
    s = __uint_as_float((uint32_t)(es + 127) << 23);
    inv_s = __uint_as_float((uint32_t)(127 - es) << 23);
}
// packed fp32 pairs (sm_100: add / mul / fma .f32x2 do two lanes per instruction; the epilogues are
// issue-bound).  -DUE_NO_F32X2 keeps the scalar forms.
__device__ __forceinline__ void ue_add2(float& x0, float& x1, float a0, float a1, float b0, float b1) {
#ifdef UE_NO_F32X2
    x0 = a0 + b0; x1 = a1 + b1;
#else
    asm("{\n\t.reg .b64 a, b, c;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tadd.rn.f32x2 c, a, b;\n\tmov.b64 {%0, %1}, c;\n\t}"
        : "=f"(x0), "=f"(x1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
#endif
}
__device__ __forceinline__ void ue_mul2(float& x0, float& x1, float a0, float a1, float b0, float b1) {
#ifdef UE_NO_F32X2
    x0 = a0 * b0; x1 = a1 * b1;
#else
    asm("{\n\t.reg .b64 a, b, c;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tmul.rn.f32x2 c, a, b;\n\tmov.b64 {%0, %1}, c;\n\t}"
        : "=f"(x0), "=f"(x1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
#endif
}
__device__ __forceinline__ void ue_fma2(float& x0, float& x1, float a0, float a1, float b0, float b1, float c0, float c1) {
#ifdef UE_NO_F32X2
    x0 = fmaf(a0, b0, c0); x1 = fmaf(a1, b1, c1);
#else
    asm("{\n\t.reg .b64 a, b, c, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tmov.b64 c, {%6, %7};\n\tfma.rn.f32x2 d, a, b, c;\n\tmov.b64 {%0, %1}, d;\n\t}"
        : "=f"(x0), "=f"(x1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
#endif
}
// split eight scaled values into f16 hi / lo and store them as one 16-byte k-chunk of row r
__device__ __forceinline__ void ue_store_chunk(uint8_t* hi_plane, int a_lo, int chunk, int r, const float (&v)[8]) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 hh = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        const float2 hf = __half22float2(hh);
        float l0, l1;
        ue_fma2(l0, l1, hf.x, hf.y, -1.0f, -1.0f, v[2 * i], v[2 * i + 1]);       // v - hi (exact: a product with -1)
        const __half2 ll = __floats2half2_rn(l0, l1);
        hw[i] = *reinterpret_cast<const uint32_t*>(&hh);
        lw[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    uint8_t* p = hi_plane + (size_t)chunk * (UE_ROWS * 16) + r * 16;
    *reinterpret_cast<uint4*>(p) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(p + a_lo) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}
// the "ones" column of a biased operand: first element of k-chunk `chunk` = s (rest of the chunk
// and the lo plane stay zero from initialisation)
__device__ __forceinline__ void ue_store_one(uint8_t* hi_plane, int chunk, int r, float s) {
    *reinterpret_cast<__half*>(hi_plane + (size_t)chunk * (UE_ROWS * 16) + r * 16) = __float2half_rn(s);
}
// operand (half) complete: make it visible to the tensor core and release the MMA issuer (one arrive
// per warp).  which = 0 / 1: that half; 2: both (operands written in one go)
__device__ __forceinline__ void ue_operand_ready(const ULayout& L, int which) {
    umma::fence_proxy_async();
    umma::tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        if (which != 1) umma::mbar_arrive_remote(ue_bars(L).aready, 0);
        if (which != 0) umma::mbar_arrive_remote(ue_bars(L).aready + 1, 0);
    }
}
// wait for accumulator barrier `which` of the current GEMM group
__device__ __forceinline__ void ue_wait_acc(const ULayout& L, UCw& c, int which) {
    UE_T0();
    umma::mbar_wait(ue_bars(L).dfull + which, (which ? c.gi : c.n0) & 1);
    UE_ACC(which, blockIdx.x == 0 && threadIdx.x == 0);
    // per GEMM group of a value+gradient evaluation (counters 8..13: G1, G2, G3, G3T, G2T, G1T)
    UE_ACC(8 + ((int)(c.gi % (6 * L.K)) < 3 * L.K ? (int)(c.gi % (6 * L.K)) % 3 : 3 + (int)(c.gi % (6 * L.K) - 3 * L.K) % 3),
           blockIdx.x == 0 && threadIdx.x == 0);
    umma::tc_fence_after();
}
// v = (main + cross) (1 + dl) cf: dl compensates the mean shrink of the truncating accumulation
// (one FMA: c + c*dl is rounded once, so a correction below one ulp is applied "on average");
// cf is a power of two (exact).
__device__ __forceinline__ void ue_acc16(uint32_t ta_main, uint32_t ta_cross, float cf, float dl, float (&v)[16]) {
    // (x8 loads: the thread's first column is a multiple of 8, not of 16)
    uint32_t a0[8], a1[8], b0[8], b1[8];
    umma::tmem_ld8(ta_main, a0);
    umma::tmem_ld8(ta_main + 8, a1);
    umma::tmem_ld8(ta_cross, b0);
    umma::tmem_ld8(ta_cross + 8, b1);
    umma::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float t0 = __uint_as_float(a0[i]) + __uint_as_float(b0[i]), t1 = __uint_as_float(a1[i]) + __uint_as_float(b1[i]);
        v[i] = fmaf(t0, dl, t0) * cf;
        v[8 + i] = fmaf(t1, dl, t1) * cf;
    }
}
__device__ __forceinline__ void ue_acc8(uint32_t ta_main, uint32_t ta_cross, float cf, float dl, float (&v)[8]) {
    uint32_t a[8], b[8];
    umma::tmem_ld8(ta_main, a);
    umma::tmem_ld8(ta_cross, b);
    umma::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float t = __uint_as_float(a[i]) + __uint_as_float(b[i]); v[i] = fmaf(t, dl, t) * cf; }
}

// Epilogue of column block G of a wide GEMM into the hidden operand: the thread's NC8 chunks of 8
// columns (one pass over tensor memory; the operand scale is known beforehand, see the header).
//   FWD:  h = relu(acc)   and the ReLU mask bits of these columns are produced (mask[] |=)
//   !FWD: h = mask ? acc : 0                                           (mask[] in)
// ta_main / ta_cross: tensor-memory addresses (lane base included) of the thread's first column of the
// block; sc = (un-scale of the accumulators) x (1 + dl) x (scale of the new operand); ssq += sum of
// squares of the stored (scaled) values.
template <bool FWD, int NC8, int G>
__device__ __forceinline__ void ue_epi_block(const ULayout& L, const UCw& c, uint32_t ta_main, uint32_t ta_cross, float sc,
                                             uint32_t (&mask)[3], float& ssq) {
#ifdef UE_X_NOEPI      // (timing experiments of profiles/r02_rowtile_experiments.md: the kernel without its wide epilogues)
    return;
#endif
    uint8_t* hp = ue_smem + L.s_h;
    const int chunk0 = ((2 * G + c.h) * L.WQ) / 8 + c.g * NC8;
    float ssq2 = 0.f;                  // second lane of the packed sum of squares
    // software-pipelined over the chunks: the loads of chunk j+1 are in flight while chunk j is scaled /
    // split / stored (tcgen05.wait::ld waits for everything outstanding, so the next loads are issued
    // right AFTER the wait).  Issuing all loads of the block up front was measured: no difference.
    uint32_t a[2][8], b[2][8];
    umma::tmem_ld8(ta_main, a[0]);
    umma::tmem_ld8(ta_cross, b[0]);
#pragma unroll
    for (int j = 0; j < NC8; ++j) {
        umma::tmem_ld_wait();
        if (j + 1 < NC8) {
            umma::tmem_ld8(ta_main + 8 * (j + 1), a[(j + 1) & 1]);
            umma::tmem_ld8(ta_cross + 8 * (j + 1), b[(j + 1) & 1]);
        }
        const int bit0 = (G * NC8 + j) * 8;
        const uint32_t mw = FWD ? 0u : mask[bit0 >> 5] >> (bit0 & 31);
        uint32_t bits = 0;
        float hv[8];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            float v0, v1;
            ue_add2(v0, v1, __uint_as_float(a[j & 1][i]), __uint_as_float(a[j & 1][i + 1]),
                    __uint_as_float(b[j & 1][i]), __uint_as_float(b[j & 1][i + 1]));
            ue_mul2(v0, v1, v0, v1, sc, sc);
            if (FWD) {
                const bool on0 = v0 > 0.f, on1 = v1 > 0.f;
                bits |= ((on0 ? 1u : 0u) << i) | ((on1 ? 1u : 0u) << (i + 1));
                hv[i] = on0 ? v0 : 0.f; hv[i + 1] = on1 ? v1 : 0.f;
            } else {
                hv[i] = ((mw >> i) & 1u) ? v0 : 0.f; hv[i + 1] = ((mw >> (i + 1)) & 1u) ? v1 : 0.f;
            }
            ue_fma2(ssq, ssq2, hv[i], hv[i + 1], hv[i], hv[i + 1], ssq, ssq2);
        }
        if (FWD) mask[bit0 >> 5] |= bits << (bit0 & 31);
        ue_store_chunk(hp, L.hplane, chunk0 + j, c.r, hv);
    }
    ssq += ssq2;
}
// Both blocks of a wide GEMM, with the barrier protocol around them.  bound_fn(): upper bound of the
// row maximum of the result (evaluated before the first accumulator wait; it may read the hand-off of
// the previous epilogue, which ue_ssq_get orders with a CTA barrier); cf: un-scale of the accumulators (operand scale x weight scale); returns the
// new operand's inverse scale and (ssq_out) the thread's part of its squared row norm in UNSCALED units.
// between(blk): work after the wait for block blk (type 0: the v columns behind block c.g).
template <bool FWD, bool BIASED, int NC8, class FB, class F>
__device__ __forceinline__ float ue_epi_wide(const ULayout& L, UCw& c, const UType& t, float cf, float dl,
                                             uint32_t (&mask)[3], float& ssq_out, FB bound_fn, F between) {
    const uint32_t t0 = c.tmem + c.tl + t.dcol + c.g * (L.WQ / 2);      // the thread's first column of block 0 (main)
    float ssq = 0.f;
    if (FWD) { mask[0] = 0u; mask[1] = 0u; mask[2] = 0u; }
    // the operand scale is computed BEFORE the accumulator wait (it depends on the previous epilogue's
    // hand-off at most, never on this GEMM): off the critical path of the round trip
    float s, inv_s;
    ue_scale_of<BIASED>(bound_fn(), s, inv_s);
    // un-scale, compensate and re-scale with one factor: s and cf are powers of two, (1 + dl) is a
    // few ulp for these long chains
    const float sc = cf * s * (1.0f + dl);
    ue_wait_acc(L, c, 0);
    ++c.n0;
    if (BIASED && c.q == 0) ue_store_one(ue_smem + L.s_h, L.W / 8, c.r, s);
    between(0);
    ue_epi_block<FWD, NC8, 0>(L, c, t0, t0 + t.nbh, sc, mask, ssq);
    ue_operand_ready(L, 0);
    ue_wait_acc(L, c, 1);
    ++c.gi;
    between(1);
    ue_epi_block<FWD, NC8, 1>(L, c, t0 + 2 * t.nbh, t0 + 3 * t.nbh, sc, mask, ssq);
    ssq_out = ssq * inv_s * inv_s;
    return inv_s;
}

// ReLU-mask scratch: word w of (layer, which, group q) of global row `grow`
__device__ __forceinline__ uint32_t* ue_mask_ptr(uint32_t* scratch, long long n_stride, int layer, int which, int q, long long grow) {
    return scratch + ((size_t)((layer * 2 + which) * 4 + q) * 3) * n_stride + grow;
}

// d log N(z)/dz seed, log q pieces etc. are produced here.  On entry zc[] holds the thread's 8
// columns of x; on exit (GRAD) gs[] holds d log q / dx of those columns; returns log q of the row
// (identical in the four threads of the row).
template <bool GRAD, int NCH>
__device__ __forceinline__ float ue_flow_eval(const ULayout& L, UCw& c, const float* __restrict__ blob_f,
                                              uint32_t* mscratch, long long n_stride, long long grow,
                                              float (&zc)[UE_DQ], float (&gs)[UE_DQ]) {
    uint8_t* const zp = ue_smem + L.s_z;
    const float* scal = blob_f + L.o_scal;
    const float sqrt_d = sqrtf((float)L.d);
    float inv_s_z, inv_s_h = 1.f, zmax;
    // x -> z operand
    {
        float m = 0.f, dummy = 0.f;
#pragma unroll
        for (int i = 0; i < UE_DQ; ++i) m = fmaxf(m, fabsf(zc[i]));
        ue_row_max2(L, c, m, dummy);
        zmax = m;
        float s;
        ue_scale_of<true>(m, s, inv_s_z);
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = zc[i] * s;
        ue_store_chunk(zp, L.zplane, c.q, c.r, v);
        if (c.q == 0) ue_store_one(zp, L.d / 8, c.r, s);
        ue_operand_ready(L, 2);
    }
    float sacc = 0.f;
    const uint32_t tl = c.tmem + c.tl;
    for (int k = L.K - 1; k >= 0; --k) {
        const float* sk = scal + k * UE_SCAL;
        float vreg[UE_DQ];
        uint32_t mask[3];
        float ssq;
        // ---- G1: [v | h1pre];  |h1pre_n| <= ||z||_2 ||W[:,n]||_2 + |b_n|,  ||z||_2 <= sqrt(d) max|z| ----
        {
            const UType& t = L.t[0];
            const float cf = inv_s_z * __ldg(sk + 0);
            const float bound = fmaf(zmax * sqrt_d, __ldg(sk + 8), __ldg(sk + 12));
            inv_s_h = ue_epi_wide<true, true, NCH>(L, c, t, cf, L.dl[1], mask, ssq, [&]() { return bound; }, [&](int blk) {
                if (blk == c.g) {        // the thread's v columns sit behind the hidden columns of block c.g
                    const uint32_t tv = tl + t.dcol + c.g * 2 * t.nbh + L.WQ;
                    ue_acc8(tv, tv + t.nbh, cf, L.dl[0], vreg);
                }
            });
            ue_ssq_put(L, c, ssq);
            if (GRAD)
#pragma unroll
                for (int w = 0; w < 3; ++w) ue_mask_ptr(mscratch, n_stride, k, 0, c.q, grow)[(size_t)w * n_stride] = mask[w];
            ue_operand_ready(L, 1);
        }
        // ---- G2: h2pre -----------------------------------------------------------------------
        {
            const UType& t = L.t[1];
            const float cf = inv_s_h * __ldg(sk + 1);
            const float c2 = __ldg(sk + 9), bm = __ldg(sk + 13);
            inv_s_h = ue_epi_wide<true, true, NCH>(L, c, t, cf, L.dl[2], mask, ssq,
                                                   [&]() { return fmaf(sqrtf(ue_ssq_get(L, c)), c2, bm); }, [](int) {});
            if (GRAD)
#pragma unroll
                for (int w = 0; w < 3; ++w) ue_mask_ptr(mscratch, n_stride, k, 1, c.q, grow)[(size_t)w * n_stride] = mask[w];
            ue_operand_ready(L, 1);
        }
        // ---- G3: [shift | scale] of the thread's transformed columns, coupling inverse -------------
        {
            const UType& t = L.t[2];
            const float cf = inv_s_h * __ldg(sk + 2);
            ue_wait_acc(L, c, 1);
            ++c.gi;
            if (c.g == 1) {
                float p[16];
                ue_acc16(tl + t.dcol, tl + t.dcol + t.nbh, cf, L.dl[3], p);
                uint32_t sv[16];
#pragma unroll
                for (int i = 0; i < UE_DQ; ++i) {
                    const float shift = p[i], scale = p[UE_DQ + i];
                    const float es = expf(-scale);
                    const float y2 = (vreg[i] - shift) * es;
                    zc[i] = y2;
                    sacc += scale;
                    sv[i] = __float_as_uint(y2);
                    sv[UE_DQ + i] = __float_as_uint(es);
                }
                if (GRAD) umma::tmem_st16(tl + UE_ACC_COLS + 16 * k, sv);       // (completion: one wait before the gradient sweep)
            } else {
#pragma unroll
                for (int i = 0; i < UE_DQ; ++i) zc[i] = vreg[i];
            }
            if (k > 0) {          // next layer's z operand
                float m = 0.f, dummy = 0.f;
#pragma unroll
                for (int i = 0; i < UE_DQ; ++i) m = fmaxf(m, fabsf(zc[i]));
                ue_row_max2(L, c, m, dummy);
                zmax = m;
                float s;
                ue_scale_of<true>(m, s, inv_s_z);
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = zc[i] * s;
                ue_store_chunk(zp, L.zplane, c.q, c.r, v);
                if (c.q == 0) ue_store_one(zp, L.d / 8, c.r, s);
                ue_operand_ready(L, 2);
            }
        }
    }
    // ---- base Gaussian -------------------------------------------------------------------------
    float gpart = 0.f;
#pragma unroll
    for (int i = 0; i < UE_DQ; ++i) {
        const int j = UE_DQ * c.q + i;
        const float lsc = __ldg(blob_f + L.o_lsc + j), inv = expf(-lsc);
        const float u = (zc[i] - __ldg(blob_f + L.o_loc + j)) * inv;
        gpart += lsc + 0.5f * u * u;
        gs[i] = -u * inv;
    }
    ue_row_sum2(L, c, gpart, sacc);
    float logs = 0.f;
    for (int k = 0; k < L.K; ++k) logs += __ldg(scal + k * UE_SCAL + 7);
    const float lq = logs + (-0.5f * (float)L.d * 1.8378770664093453f - (gpart + sacc));
    if (!GRAD) return lq;
    umma::tmem_st_wait();           // the saved (y2, exp(-scale)) of all layers are in tensor memory
    // ---- input-gradient sweep ---------------------------------------------------------------------
    uint8_t* const pp = ue_smem + L.s_par;
    uint8_t* const gp = ue_smem + L.s_gv;
    for (int k = 0; k < L.K; ++k) {
        const float* sk = scal + k * UE_SCAL;
        float inv_s_par, inv_s_gv, pmax;
        // coupling backward: gparam and gv operands
        {
            float gpar[16], gv[UE_DQ];
            float ma = 0.f, mb = 0.f;
            if (c.g == 1) {
                uint32_t sv[16];
                umma::tmem_ld16(tl + UE_ACC_COLS + 16 * k, sv);
                umma::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < UE_DQ; ++i) {
                    const float y2 = __uint_as_float(sv[i]), es = __uint_as_float(sv[UE_DQ + i]);
                    const float g2 = gs[i];
                    const float gv2 = g2 * es;
                    gpar[i] = -gv2;
                    gpar[UE_DQ + i] = -g2 * y2 - 1.0f;
                    gv[i] = gv2;
                    ma = fmaxf(ma, fmaxf(fabsf(gpar[i]), fabsf(gpar[UE_DQ + i])));
                }
            } else {
#pragma unroll
                for (int i = 0; i < UE_DQ; ++i) gv[i] = gs[i];
            }
#pragma unroll
            for (int i = 0; i < UE_DQ; ++i) mb = fmaxf(mb, fabsf(gv[i]));
            ue_row_max2(L, c, ma, mb);
            pmax = ma;
            float sa, sb;
            ue_scale_of<false>(ma, sa, inv_s_par);
            ue_scale_of<false>(mb, sb, inv_s_gv);
            if (c.g == 1) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = gpar[i] * sa;
                ue_store_chunk(pp, L.pplane, c.h, c.r, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = gpar[UE_DQ + i] * sa;
                ue_store_chunk(pp, L.pplane, 2 + c.h, c.r, v);
            }
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = gv[i] * sb;
            ue_store_chunk(gp, L.pplane, c.q, c.r, v);
            ue_operand_ready(L, 2);
        }
        uint32_t mask[3];
        float ssq;
        // ---- G3T: gh2 = (gparam @ W3) * m2;  ||gparam||_2 <= sqrt(d) max|gparam| ------------------
        {
            const UType& t = L.t[3];
            const float cf = inv_s_par * __ldg(sk + 3);
#pragma unroll
            for (int w = 0; w < 3; ++w) mask[w] = ue_mask_ptr(mscratch, n_stride, k, 1, c.q, grow)[(size_t)w * n_stride];
            const float bound = pmax * sqrt_d * __ldg(sk + 10);
            inv_s_h = ue_epi_wide<false, false, NCH>(L, c, t, cf, L.dl[4], mask, ssq, [&]() { return bound; }, [](int) {});
            ue_ssq_put(L, c, ssq);
            ue_operand_ready(L, 1);
        }
        // ---- G2T: gh1 = (gh2 @ W2) * m1 ------------------------------------------------------
        {
            const UType& t = L.t[4];
            const float cf = inv_s_h * __ldg(sk + 4);
#pragma unroll
            for (int w = 0; w < 3; ++w) mask[w] = ue_mask_ptr(mscratch, n_stride, k, 0, c.q, grow)[(size_t)w * n_stride];
            const float c2 = __ldg(sk + 11);
            inv_s_h = ue_epi_wide<false, false, NCH>(L, c, t, cf, L.dl[5], mask, ssq,
                                                    [&]() { return sqrtf(ue_ssq_get(L, c)) * c2; }, [](int) {});
            ue_operand_ready(L, 1);
        }
        // ---- G1T: g = gh1 @ (W1 Wmix1^T) + gv @ Wmix^T ---------------------------------------
        {
            const float ca = inv_s_h * __ldg(sk + 5), cb = inv_s_gv * __ldg(sk + 6);
            ue_wait_acc(L, c, 1);
            ++c.gi;
            float a[8], b[8];
            const UType& t5 = L.t[5];
            const UType& t6 = L.t[6];
            ue_acc8(tl + t5.dcol + UE_DQ * c.g, tl + t5.dcol + t5.nbh + UE_DQ * c.g, ca, L.dl[6], a);
            ue_acc8(tl + t6.dcol + UE_DQ * c.g, tl + t6.dcol + t6.nbh + UE_DQ * c.g, cb, L.dl[7], b);
#pragma unroll
            for (int i = 0; i < UE_DQ; ++i) gs[i] = a[i] + b[i];
        }
    }
    // the accumulators have been read: the next evaluation's first operand_ready releases the issuer
    return lq;
}

// ---- kernel prologue / epilogue shared by the kernels --------------------------------------------
struct UEnv { uint32_t rank, tmem; int warp, lane; };

__device__ __forceinline__ UEnv ue_setup(const ULayout& L) {
    UEnv e;
    e.rank = umma::cluster_ctarank();
    e.warp = threadIdx.x >> 5; e.lane = threadIdx.x & 31;
    uint32_t* slot = reinterpret_cast<uint32_t*>(ue_smem + L.s_bar + (2 * UE_NSTAGE + 4) * 8);
    // operand buffers start from zero: pad rows, bias chunks and lo planes of the "ones" columns
    for (int i = threadIdx.x; i < L.s_ring / 16; i += UE_THREADS) reinterpret_cast<uint4*>(ue_smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        const UBars B = ue_bars(L);
        for (int s = 0; s < UE_NSTAGE; ++s) { umma::mbar_init(B.full + s, e.rank == 0 ? 2 : 1); umma::mbar_init(B.empty + s, 1); }
        umma::mbar_init(B.dfull, 2); umma::mbar_init(B.dfull + 1, 2);      // one commit per issuer
        umma::mbar_init(B.aready, 16); umma::mbar_init(B.aready + 1, 16);
        umma::mbar_fence_init();
    }
    __syncthreads();
    if (e.warp == 8) umma::tmem_alloc<2>(slot, 512);
    umma::fence_proxy_async();
    umma::tc_fence_before();
    __syncthreads();                 // the allocating warp's write of *slot, CTA-wide (as in the guide)
    umma::cluster_sync_all();        // peer barriers initialised before any remote arrive / multicast commit
    umma::tc_fence_after();
    e.tmem = *slot;
    return e;
}
__device__ __forceinline__ void ue_teardown(const UEnv& e) {
    umma::tc_fence_before();
    umma::cluster_sync_all();
    if (e.warp == 8) umma::tmem_free<2>(e.tmem, 512);
}
__device__ __forceinline__ UCw ue_cw(const UEnv& e) {
    UCw c;
    const int q4 = e.warp & 3;
    const int l = 32 * q4 + e.lane;
    c.r = l & 63; c.h = l >> 6; c.g = e.warp >> 2; c.q = 2 * c.g + c.h;
    c.tl = (uint32_t)(32 * q4) << 16;
    c.gi = 0; c.n0 = 0; c.xb = 0; c.nq = 0; c.tmem = e.tmem; c.rank = e.rank;
    return c;
}

// ---------------------------------------------------------------------------------------------
// K3/K4 on the row-tile engine: x[n,d] -> log_q[n] (+ d log q / dx)
// ---------------------------------------------------------------------------------------------
template <bool GRAD, int NCH>
__global__ void __launch_bounds__(UE_THREADS, 1)
k_flow_logprob_u(ULayout L, const uint8_t* __restrict__ blob, const float* __restrict__ x, float* __restrict__ log_q,
                 float* __restrict__ grad, uint32_t* mscratch, long long n) {
    const UEnv e = ue_setup(L);
    const long long row0 = (long long)(blockIdx.x >> 1) * 128 + (long long)e.rank * UE_ROWS;
    const long long n_stride = (long long)gridDim.x * UE_ROWS;
    if (e.warp < 8) {
        UCw c = ue_cw(e);
        const long long grow = row0 + c.r;
        const bool live = grow < n;
        float zc[UE_DQ], gs[UE_DQ];
        if (live) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(x + grow * L.d + UE_DQ * c.q));
            const float4 b = __ldg(reinterpret_cast<const float4*>(x + grow * L.d + UE_DQ * c.q + 4));
            zc[0] = a.x; zc[1] = a.y; zc[2] = a.z; zc[3] = a.w; zc[4] = b.x; zc[5] = b.y; zc[6] = b.z; zc[7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < UE_DQ; ++i) zc[i] = 0.f;
        }
        UE_T0();
        const float lq = ue_flow_eval<GRAD, NCH>(L, c, reinterpret_cast<const float*>(blob), mscratch, n_stride,
                                            (long long)blockIdx.x * UE_ROWS + c.r, zc, gs);
        UE_ACC(7, blockIdx.x == 0 && threadIdx.x == 0);
        if (live) {
            if (c.q == 0) log_q[grow] = lq;
            if (GRAD) {
                float4* gp = reinterpret_cast<float4*>(grad + grow * L.d + UE_DQ * c.q);
                gp[0] = make_float4(gs[0], gs[1], gs[2], gs[3]);
                gp[1] = make_float4(gs[4], gs[5], gs[6], gs[7]);
            }
        }
    } else if (e.warp == 8) {
        if (e.rank == 0) ue_mma<0>(L, e.tmem, 1, GRAD); else ue_relay(L, 1, GRAD);
    } else if (e.warp == 9) {
        ue_producer(L, blob, e.rank, 1, GRAD);
    } else if (e.rank == 0) {
        ue_mma<1>(L, e.tmem, 1, GRAD);
    }
    ue_teardown(e);
}

// ---------------------------------------------------------------------------------------------
// fused HMC outer step on the row-tile engine (hmc.py:129-160; same contract as k_hmc_step)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ue_load8(const float* __restrict__ p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ue_store8(float* __restrict__ p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}

// (grid_reduce_last and hmc_finish: tile_kernels.cuh, included first by fab_b200.cu)
template <int NCH>
__global__ void __launch_bounds__(UE_THREADS, 1)
k_hmc_step_u(ULayout L, const uint8_t* __restrict__ blob, fab_target_desc tgt, fab_hmc_state st, fab_hmc_args a,
             fab_point cur, fab_point prop_in, fab_point prop_out, float* __restrict__ log_w,
             const float* __restrict__ mom_noise, const float* __restrict__ exp_noise,
             const int* __restrict__ n_active, float* __restrict__ stats, float* ws, uint32_t* mscratch, long long n) {
    const UEnv e = ue_setup(L);
    __shared__ float s_rowc[UE_ROWS], s_rowd[UE_ROWS], s_blk[8];
    const long long n_act = n_active ? (long long)(*n_active) : n;
    const long long pair0 = (long long)(blockIdx.x >> 1) * 128;
    const long long row0 = pair0 + (long long)e.rank * UE_ROWS;
    const bool pair_live = pair0 < n_act;
    int np = (int)min((long long)UE_ROWS, n_act - row0);
    if (np < 0) np = 0;
    if (threadIdx.x < 8) s_blk[threadIdx.x] = 0.f;
    if (pair_live) {
        if (e.warp < 8) {
            UCw c = ue_cw(e);
            const long long grow = row0 + c.r;
            const bool live = c.r < np;
            const int d = L.d, c0 = UE_DQ * c.q;
            float px[8], pgq[8], pgp[8], mom[8], mass[8];
            float clq = 0.f, clp = 0.f, plq = 0.f, plp = 0.f;
            ue_load8(st.d_mass + c0, mass);
            if (live) {
                const fab_point& src = prop_in.d_x ? prop_in : cur;
                ue_load8(src.d_x + grow * d + c0, px);
                ue_load8(src.d_grad_log_q + grow * d + c0, pgq);
                ue_load8(src.d_grad_log_p + grow * d + c0, pgp);
                plq = src.d_log_q[grow]; plp = src.d_log_p[grow];
                clq = cur.d_log_q[grow]; clp = cur.d_log_p[grow];
                ue_load8(mom_noise + grow * d + c0, mom);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) { px[i] = 0.f; pgq[i] = 0.f; pgp[i] = 0.f; mom[i] = 0.f; }
            }
            const float eps = __fadd_rn(st.d_epsilons[(size_t)(a.i - 1) * st.n_outer + a.outer], st.d_common_epsilon[0]);
            // momentum p0 = randn * mass (hmc.py:134), kinetic energy sum(p^2/m)/2
            float kpart = 0.f, zero = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                mom[i] = __fmul_rn(mom[i], mass[i]);
                kpart += __fdiv_rn(__fmul_rn(mom[i], mom[i]), mass[i]);
            }
            ue_row_sum2(L, c, kpart, zero);
            const float ke0 = 0.5f * kpart;
            // leapfrog (hmc.py:138-147)
            for (int l = 0; l < a.L; ++l) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float gu = grad_u_of(a.g, pgq[i], pgp[i], a.max_grad);
                    mom[i] = __fsub_rn(mom[i], __fmul_rn(__fmul_rn(eps, gu), 0.5f));
                    px[i] = __fadd_rn(px[i], __fmul_rn(__fdiv_rn(eps, mass[i]), mom[i]));
                }
                float zc[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) zc[i] = px[i];
                UE_T0();
                plq = ue_flow_eval<true, NCH>(L, c, reinterpret_cast<const float*>(blob), mscratch,
                                         (long long)gridDim.x * UE_ROWS, (long long)blockIdx.x * UE_ROWS + c.r, zc, pgq);
                UE_ACC(7, blockIdx.x == 0 && threadIdx.x == 0);
                // many-well target: value and gradient of the thread's columns (pairs stay in one thread)
                float epart = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float ej, gj;
                    manywell_elem(tgt, px[i], c0 + i, ej, gj);
                    epart += ej;
                    pgp[i] = gj;
                }
                zero = 0.f;
                ue_row_sum2(L, c, epart, zero);
                plp = -epart - tgt.log_norm;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float gu = grad_u_of(a.g, pgq[i], pgp[i], a.max_grad);
                    mom[i] = __fsub_rn(mom[i], __fmul_rn(__fmul_rn(eps, gu), 0.5f));
                }
            }
            // Metropolis accept (hmc.py:105-124) and distance moved (hmc.py:173-183)
            float cx[8];
            if (live) ue_load8(cur.d_x + grow * d + c0, cx);
            float kp = 0.f, dd = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                kp += __fdiv_rn(__fmul_rn(mom[i], mom[i]), mass[i]);
                const float df = live ? cx[i] - px[i] : 0.f;
                dd += df * df;
            }
            ue_row_sum2(L, c, kp, dd);
            bool acc = false;
            if (live) {
                const float lj_cur = __fsub_rn(gamma_of(a.g, clq, clp), ke0);
                const float lj_prop = __fsub_rn(gamma_of(a.g, plq, plp), 0.5f * kp);
                float log_a = __fsub_rn(lj_prop, lj_cur);
                const bool ok = fab_isfinite(log_a);
                if (!ok) log_a = -CUDART_INF_F;
                acc = ok && (log_a > -__ldg(exp_noise + grow));
                if (c.q == 0) { s_rowc[c.r] = expf(fminf(log_a, 0.f)); s_rowd[c.r] = acc ? 0.f : sqrtf(dd); }
            } else if (c.q == 0) { s_rowc[c.r] = 0.f; s_rowd[c.r] = 0.f; }
            if (live) {
                if (acc) {       // cur[accept] = prop[accept]  (hmc.py:154)
                    ue_store8(cur.d_x + grow * d + c0, px);
                    ue_store8(cur.d_grad_log_q + grow * d + c0, pgq);
                    ue_store8(cur.d_grad_log_p + grow * d + c0, pgp);
                    clq = plq; clp = plp;
                }
                if (c.q == 0) {
                    if (acc) { cur.d_log_q[grow] = clq; cur.d_log_p[grow] = clp; }
                    if (a.update_log_w) {   // ais.py:93-100
                        const float inc = __fsub_rn(gamma_of(a.g_next, clq, clp), gamma_of(a.g_w, clq, clp));
                        log_w[grow] = __fadd_rn(log_w[grow], inc);
                    }
                }
                if (prop_out.d_x) {
                    ue_store8(prop_out.d_x + grow * d + c0, px);
                    ue_store8(prop_out.d_grad_log_q + grow * d + c0, pgq);
                    ue_store8(prop_out.d_grad_log_p + grow * d + c0, pgp);
                    if (c.q == 0) { prop_out.d_log_q[grow] = plq; prop_out.d_log_p[grow] = plp; }
                }
            }
        } else if (e.warp == 8) {
            if (e.rank == 0) ue_mma<0>(L, e.tmem, a.L, true); else ue_relay(L, a.L, true);
        } else if (e.warp == 9) {
            ue_producer(L, blob, e.rank, a.L, true);
        } else if (e.rank == 0) {
            ue_mma<1>(L, e.tmem, a.L, true);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && pair_live) {
        float s = 0.f, dsum = 0.f;
        for (int p = 0; p < np; ++p) { s += s_rowc[p]; dsum += s_rowd[p]; }
        s_blk[0] = s; s_blk[1] = dsum;
    }
    __syncthreads();
    if (grid_reduce_last<2>(s_blk, ws, s_blk + 4)) {
        if (threadIdx.x == 0) {
            stats[0] = s_blk[4]; stats[1] = (float)n_act; stats[2] = s_blk[5]; stats[3] = 0.f;
            if (!a.defer_stats) hmc_finish(st, a, s_blk[4], (float)n_act, s_blk[5]);
        }
    }
    ue_teardown(e);
}
