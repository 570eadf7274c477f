// Global kernels built on the tile device functions: flow sample / log_prob(+grad), chain
// initialisation, the fused HMC outer step and the fused Metropolis transition.
#pragma once
#include "flow_tile.cuh"
#include "target_tile.cuh"
#include "reduce_finish.cuh"

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_rows(float* dst, int ldd, const float* __restrict__ src, int d,
                                          long long row0, int np, int T) {
    // dst[T][ldd] <- src[row0 .. row0+np)[d], zero elsewhere (incl. pad columns)
    for (int i = threadIdx.x; i < T * ldd; i += FAB_NT) {
        const int p = i / ldd, j = i - p * ldd;
        dst[i] = (p < np && j < d) ? __ldg(src + (row0 + p) * d + j) : 0.f;
    }
}
__device__ __forceinline__ void store_rows(float* __restrict__ dst, int d, const float* src, int lds,
                                           long long row0, int np) {
    for (int i = threadIdx.x; i < np * d; i += FAB_NT) {
        const int p = i / d, j = i - p * d;
        dst[(row0 + p) * d + j] = src[p * lds + j];
    }
}
__device__ __forceinline__ void copy_tile(float* dst, const float* src, int n) {
    for (int i = threadIdx.x; i < n; i += FAB_NT) dst[i] = src[i];
}
// operand buffer (mma_gemm.cuh: element (slot p, row j) at j*S + p) <-> global rows / row-major
// shared tiles.  Slots >= np are fed zeros.
template <int TP>
__device__ __forceinline__ void load_rows_act(float* dst, const float* __restrict__ src, int d,
                                              long long row0, int np) {
    constexpr int S = ActL<TP>::S;
    for (int e = threadIdx.x; e < d * TP; e += FAB_NT) {
        const int j = e / TP, p = e - j * TP;
        dst[(size_t)j * S + p] = p < np ? __ldg(src + (row0 + p) * d + j) : 0.f;
    }
}
template <int TP>
__device__ __forceinline__ void store_rows_act(float* __restrict__ dst, int d, const float* src,
                                               long long row0, int np) {
    constexpr int S = ActL<TP>::S;
    for (int i = threadIdx.x; i < np * d; i += FAB_NT) {
        const int p = i / d, j = i - p * d;
        dst[(row0 + p) * d + j] = src[(size_t)j * S + p];
    }
}
template <int TP>
__device__ __forceinline__ void rows_to_act(float* dst, const float* src_rows, int d, int DP, int T) {
    constexpr int S = ActL<TP>::S;
    for (int e = threadIdx.x; e < d * TP; e += FAB_NT) {
        const int j = e / TP, p = e - j * TP;
        dst[(size_t)j * S + p] = p < T ? src_rows[p * DP + j] : 0.f;
    }
}
template <int TP>
__device__ __forceinline__ void act_to_rows(float* dst_rows, const float* src, int d, int DP, int T) {
    constexpr int S = ActL<TP>::S;
    for (int i = threadIdx.x; i < T * DP; i += FAB_NT) {
        const int p = i / DP, j = i - p * DP;
        dst_rows[i] = j < d ? src[(size_t)j * S + p] : 0.f;
    }
}

// Evaluate log q (+grad), log p (+grad) at the rows x[T][DP] (shared, row-major, pads zero).
// Outputs in shared: lq[TP], lp[T], gq[T][DP], gp[T][DP] (the last two only when GRAD).
template <int TP, bool GRAD>
__device__ void eval_point(const TileLayout& L, const fab_flow_desc& f,
                           const float* __restrict__ blob, const fab_target_desc& tgt,
                           const float* x, float* lq, float* lp, float* gq, float* gp) {
    const TileBufs b = tile_bufs(L);
    prof_mark(1);
    rows_to_act<TP>(zsel(L, 0), x, L.d, L.DP, L.T);
    __syncthreads();
    prof_mark(2);
    flow_inverse<TP, GRAD>(L, f, blob, 0, lq);
    if (GRAD) {
        flow_backward<TP>(L, f, blob);
        act_to_rows<TP>(gq, b.gs, L.d, L.DP, L.T);
    }
    target_tile(tgt, x, L.DP, L.d, L.T, lp, GRAD ? gp : nullptr);
    __syncthreads();
    prof_mark(17);
}

// ---------------------------------------------------------------------------------------------
// K1/K2  flow sample      K3/K4  flow log_prob (+ input gradient)
// ---------------------------------------------------------------------------------------------
template <int TP>
__global__ void __launch_bounds__(FAB_NT, FAB_MIN_CTAS)
k_flow_sample(TileLayout L, fab_flow_desc f, const float* __restrict__ blob,
              const float* __restrict__ eps, float* __restrict__ x, float* __restrict__ log_q,
              long long n) {
    float* const smem = fab_smem;
    const TileBufs b = tile_bufs(L);
    float* lq = smem + L.o_state;
    const long long row0 = (long long)blockIdx.x * L.T;
    const int np = (int)min((long long)L.T, n - row0);
    tile_init(L, f, blob);
    load_rows_act<TP>(zsel(L, 0), eps, L.d, row0, np);
    __syncthreads();
    flow_sample<TP>(L, f, blob, 0, lq);
    store_rows_act<TP>(x, L.d, zsel(L, 0), row0, np);
    for (int p = threadIdx.x; p < np; p += FAB_NT) log_q[row0 + p] = lq[p];
}

template <int TP, bool GRAD>
__global__ void __launch_bounds__(FAB_NT, FAB_MIN_CTAS)
k_flow_logprob(TileLayout L, fab_flow_desc f, const float* __restrict__ blob,
               const float* __restrict__ x, float* __restrict__ log_q, float* __restrict__ grad,
               long long n) {
    float* const smem = fab_smem;
    const TileBufs b = tile_bufs(L);
    float* lq = smem + L.o_state;
    const long long row0 = (long long)blockIdx.x * L.T;
    const int np = (int)min((long long)L.T, n - row0);
    tile_init(L, f, blob);
    load_rows_act<TP>(zsel(L, 0), x, L.d, row0, np);
    __syncthreads();
    flow_inverse<TP, GRAD>(L, f, blob, 0, lq);
    if (GRAD) {
        flow_backward<TP>(L, f, blob);
        store_rows_act<TP>(grad, L.d, b.gs, row0, np);
    }
    for (int p = threadIdx.x; p < np; p += FAB_NT) log_q[row0 + p] = lq[p];
}

// ---------------------------------------------------------------------------------------------
// chain initialisation (ais.py:56-65)
// state area: x[T][DP], gq[T][DP], gp[T][DP], lq0[T], lq[T], lp[T]
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline int init_state_floats(int T, int DP) { return 3 * T * DP + 3 * 16; }

template <int TP, bool GRAD>
__global__ void __launch_bounds__(FAB_NT, FAB_MIN_CTAS)
k_ais_init(TileLayout L, fab_flow_desc f, const float* __restrict__ blob, fab_target_desc tgt,
           const float* __restrict__ eps, fab_gamma g1, fab_point out, float* __restrict__ log_w,
           float* __restrict__ log_q0, uint8_t* __restrict__ valid, long long n) {
    float* const smem = fab_smem;
    const TileBufs b = tile_bufs(L);
    const int T = L.T;
    float* sx = smem + L.o_state;
    float* sgq = sx + T * L.DP;
    float* sgp = sgq + T * L.DP;
    float* slq0 = sgp + T * L.DP;
    float* slq = slq0 + 16;
    float* slp = slq + 16;
    const long long row0 = (long long)blockIdx.x * T;
    const int np = (int)min((long long)T, n - row0);
    prof_mark(-1);
    tile_init(L, f, blob);
    load_rows_act<TP>(zsel(L, 0), eps, L.d, row0, np);
    __syncthreads();
    flow_sample<TP>(L, f, blob, 0, slq0);
    act_to_rows<TP>(sx, zsel(L, 0), L.d, L.DP, T);
    __syncthreads();
    if (GRAD) {
        // create_point(with_grad=True) re-evaluates log q by the inverse pass
        eval_point<TP, true>(L, f, blob, tgt, sx, slq, slp, sgq, sgp);
    } else {
        target_tile(tgt, sx, L.DP, L.d, T, slp, nullptr);
        for (int p = threadIdx.x; p < T; p += FAB_NT) slq[p] = slq0[p];
        __syncthreads();
    }
    store_rows(out.d_x, L.d, sx, L.DP, row0, np);
    if (GRAD) {
        store_rows(out.d_grad_log_q, L.d, sgq, L.DP, row0, np);
        store_rows(out.d_grad_log_p, L.d, sgp, L.DP, row0, np);
    }
    for (int p = threadIdx.x; p < np; p += FAB_NT) {
        const float lq = slq[p], lp = slp[p], lq0 = slq0[p];
        out.d_log_q[row0 + p] = lq;
        out.d_log_p[row0 + p] = lp;
        log_w[row0 + p] = __fsub_rn(gamma_of(g1, lq, lp), lq0);
        if (log_q0) log_q0[row0 + p] = lq0;
        valid[row0 + p] = (fab_isfinite(lq) && fab_isfinite(lp)) ? 1 : 0;
    }
}

// Tuner statistics exchanged over NVLink peer memory instead of an NCCL all-reduce (the
// (sum of clamped acceptance, count, distance) triple of hmc.py:122-123,162-170 summed over the ranks,
// then the identical update on every rank).  Every rank owns a symmetric buffer
//     slot[FAB_PEER_RING][world][8 floats]    (3 statistics, pad, sequence flag, pad)
// mapped into all peers.  Exchange number s (a device-resident counter, so the launch can sit in a
// captured CUDA graph): lane r stores this rank's triple into slot[s % RING][my rank] of peer r,
// fences, stores the flag s + 1; then lane r spins on slot[s % RING][r] of the LOCAL buffer until
// rank r's flag arrives and reads its triple.  The triples are added in rank order, so every rank gets
// the same bits (and the same as the all-reduce of the single-process sum order is NOT required: the
// tuner only compares log(mean) with a threshold).  A rank is at most one exchange ahead of any
// peer (it cannot finish exchange s + 1 without that peer's s + 1 flag), so RING >= 2 keeps a fast
// rank's next store away from a slow rank's read.  A flag that does not arrive within ~4 s traps
// instead of hanging the GPU.
#define FAB_PEER_RING 4
__global__ void k_hmc_finish_peer(fab_hmc_state st, fab_hmc_args a, const float* __restrict__ stats,
                                  float* const* __restrict__ peer_bufs, int world, int rank,
                                  unsigned int* __restrict__ seq_counter) {
    __shared__ float s_sum[3][32];
    const int r = threadIdx.x;
    const unsigned int seq = *seq_counter;
    const size_t slot = (size_t)(seq % FAB_PEER_RING) * world;
    if (r < world) {
        float* dst = peer_bufs[r] + (slot + rank) * 8;
        dst[0] = stats[0]; dst[1] = stats[1]; dst[2] = stats[2];
        __threadfence_system();
        *reinterpret_cast<volatile unsigned int*>(dst + 4) = seq + 1u;
        const float* src = peer_bufs[rank] + (slot + r) * 8;
        const volatile unsigned int* flag = reinterpret_cast<const volatile unsigned int*>(src + 4);
        const long long t0 = clock64();
        while (*flag != seq + 1u) {
            if (clock64() - t0 > 8000000000ll) {
                printf("k_hmc_finish_peer: rank %d waited 4 s for rank %d (exchange %u)\n", rank, r, seq);
                __trap();
            }
        }
        __threadfence_system();
        const volatile float* vs = src;
        s_sum[0][r] = vs[0]; s_sum[1][r] = vs[1]; s_sum[2][r] = vs[2];
    }
    __syncthreads();
    if (r == 0) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        for (int q = 0; q < world; ++q) { s0 += s_sum[0][q]; s1 += s_sum[1][q]; s2 += s_sum[2][q]; }
        hmc_finish(st, a, s0, s1, s2);
        *seq_counter = seq + 1u;
    }
}

__global__ void k_hmc_finish(fab_hmc_state st, fab_hmc_args a, const float* __restrict__ stats) {
    if (threadIdx.x == 0 && blockIdx.x == 0) hmc_finish(st, a, stats[0], stats[1], stats[2]);
}

// ---------------------------------------------------------------------------------------------
// fused HMC outer step (hmc.py:129-160)
// state area: cur{x,gq,gp}, prop{x,gq,gp}, mom : 7*[T][DP]; scalars cur_lq,cur_lp,prop_lq,prop_lp,
// ke0, red2[4]
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline int hmc_state_floats(int T, int DP) { return 7 * T * DP + 8 * 16 + 8; }

template <int TP>
__global__ void __launch_bounds__(FAB_NT, FAB_MIN_CTAS)
k_hmc_step(TileLayout L, fab_flow_desc f, const float* __restrict__ blob, fab_target_desc tgt,
           fab_hmc_state st, fab_hmc_args a, fab_point cur, fab_point prop_in, fab_point prop_out,
           float* __restrict__ log_w, const float* __restrict__ mom_noise,
           const float* __restrict__ exp_noise, const int* __restrict__ n_active,
           float* __restrict__ stats, float* ws, long long n) {
    float* const smem = fab_smem;
    const TileBufs b = tile_bufs(L);
    const int T = L.T;
    const int TD = T * L.DP, T4 = 16;
    float* cx = smem + L.o_state;
    float* cgq = cx + TD;  float* cgp = cgq + TD;
    float* px = cgp + TD;  float* pgq = px + TD;  float* pgp = pgq + TD;
    float* mom = pgp + TD;
    float* clq = mom + TD; float* clp = clq + T4;
    float* plq = clp + T4; float* plp = plq + T4;
    float* ke0 = plp + T4; float* acc_flag = ke0 + T4;
    float* contrib = acc_flag + T4; float* moved = contrib + T4;
    float* blk = moved + T4;             // [0]=sum exp(min(log_a,0)), [1]=sum dist, [4..]=totals

    const long long n_act = n_active ? (long long)(*n_active) : n;
    const long long row0 = (long long)blockIdx.x * T;
    int np = (int)min((long long)T, n_act - row0);
    if (np < 0) np = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    prof_mark(-1);
#ifdef FAB_PROF
    const long long cta_t0 = clock64();
#endif
    if (np > 0) {
        tile_init(L, f, blob);
        load_rows(cx, L.DP, cur.d_x, L.d, row0, np, T);
        load_rows(cgq, L.DP, cur.d_grad_log_q, L.d, row0, np, T);
        load_rows(cgp, L.DP, cur.d_grad_log_p, L.d, row0, np, T);
        for (int p = threadIdx.x; p < T; p += FAB_NT) {
            clq[p] = p < np ? cur.d_log_q[row0 + p] : 0.f;
            clp[p] = p < np ? cur.d_log_p[row0 + p] : 0.f;
        }
        if (prop_in.d_x) {
            load_rows(px, L.DP, prop_in.d_x, L.d, row0, np, T);
            load_rows(pgq, L.DP, prop_in.d_grad_log_q, L.d, row0, np, T);
            load_rows(pgp, L.DP, prop_in.d_grad_log_p, L.d, row0, np, T);
            for (int p = threadIdx.x; p < T; p += FAB_NT) {
                plq[p] = p < np ? prop_in.d_log_q[row0 + p] : 0.f;
                plp[p] = p < np ? prop_in.d_log_p[row0 + p] : 0.f;
            }
        }
        __syncthreads();
        if (!prop_in.d_x) {
            copy_tile(px, cx, TD); copy_tile(pgq, cgq, TD); copy_tile(pgp, cgp, TD);
            for (int p = threadIdx.x; p < T; p += FAB_NT) { plq[p] = clq[p]; plp[p] = clp[p]; }
        }
        // step size eps(i, n) = epsilons[i-1, n] + common_epsilon   (hmc.py:90-100)
        const float eps = __fadd_rn(st.d_epsilons[(size_t)(a.i - 1) * st.n_outer + a.outer],
                                    st.d_common_epsilon[0]);
        // momentum p0 = randn * mass  (hmc.py:134) and its kinetic energy sum(p^2/m)/2
        for (int i = threadIdx.x; i < TD; i += FAB_NT) {
            const int p = i / L.DP, j = i - p * L.DP;
            mom[i] = (p < np && j < L.d)
                         ? __fmul_rn(__ldg(mom_noise + (row0 + p) * L.d + j), __ldg(st.d_mass + j))
                         : 0.f;
        }
        __syncthreads();
        for (int p = warp; p < T; p += FAB_NWARPS) {
            float s = 0.f;
            for (int j = lane; j < L.d; j += 32) {
                const float m = mom[p * L.DP + j];
                s += __fdiv_rn(__fmul_rn(m, m), __ldg(st.d_mass + j));
            }
            s = warp_sum(s);
            if (lane == 0) ke0[p] = 0.5f * s;
        }
        __syncthreads();        // the first leapfrog half-step below overwrites mom (racecheck)
        prof_mark(0);
        // leapfrog (hmc.py:138-147)
        for (int l = 0; l < a.L; ++l) {
            for (int i = threadIdx.x; i < TD; i += FAB_NT) {
                const int j = i % L.DP;
                if (j < L.d) {
                    const float gu = grad_u_of(a.g, pgq[i], pgp[i], a.max_grad);
                    const float m = __fsub_rn(mom[i], __fmul_rn(__fmul_rn(eps, gu), 0.5f));
                    mom[i] = m;
                    px[i] = __fadd_rn(px[i], __fmul_rn(__fdiv_rn(eps, __ldg(st.d_mass + j)), m));
                }
            }
            __syncthreads();
            eval_point<TP, true>(L, f, blob, tgt, px, plq, plp, pgq, pgp);
            for (int i = threadIdx.x; i < TD; i += FAB_NT) {
                const int j = i % L.DP;
                if (j < L.d) {
                    const float gu = grad_u_of(a.g, pgq[i], pgp[i], a.max_grad);
                    mom[i] = __fsub_rn(mom[i], __fmul_rn(__fmul_rn(eps, gu), 0.5f));
                }
            }
            __syncthreads();
        }
        // Metropolis accept (hmc.py:105-124) + "distance moved" (hmc.py:173-183, measured against
        // the already-overwritten current point: 0 for accepted particles)
        if (threadIdx.x < 2) blk[threadIdx.x] = 0.f;
        for (int p = warp; p < T; p += FAB_NWARPS) {
            float s = 0.f, dd = 0.f;
            for (int j = lane; j < L.d; j += 32) {
                const float m = mom[p * L.DP + j];
                s += __fdiv_rn(__fmul_rn(m, m), __ldg(st.d_mass + j));
                const float df = cx[p * L.DP + j] - px[p * L.DP + j];
                dd += df * df;
            }
            s = warp_sum(s);
            dd = warp_sum(dd);
            if (lane == 0) {
                float flag = 0.f, c_p = 0.f, d_p = 0.f;
                if (p < np) {
                    const float lj_cur = __fsub_rn(gamma_of(a.g, clq[p], clp[p]), ke0[p]);
                    const float lj_prop = __fsub_rn(gamma_of(a.g, plq[p], plp[p]), 0.5f * s);
                    float log_a = __fsub_rn(lj_prop, lj_cur);
                    const bool ok = fab_isfinite(log_a);
                    if (!ok) log_a = -CUDART_INF_F;
                    const bool acc = ok && (log_a > -__ldg(exp_noise + row0 + p));
                    flag = acc ? 1.f : 0.f;
                    // per-particle contributions; summed over p below in fixed order
                    c_p = expf(fminf(log_a, 0.f));
                    d_p = acc ? 0.f : sqrtf(dd);
                }
                acc_flag[p] = flag; contrib[p] = c_p; moved[p] = d_p;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f, dsum = 0.f;
            for (int p = 0; p < np; ++p) { s += contrib[p]; dsum += moved[p]; }
            blk[0] = s; blk[1] = dsum;
        }
        // overwrite accepted rows: cur[accept] = prop[accept]  (hmc.py:154)
        for (int i = threadIdx.x; i < TD; i += FAB_NT) {
            const int p = i / L.DP;
            if (acc_flag[p] != 0.f) { cx[i] = px[i]; cgq[i] = pgq[i]; cgp[i] = pgp[i]; }
        }
        for (int p = threadIdx.x; p < T; p += FAB_NT)
            if (acc_flag[p] != 0.f) { clq[p] = plq[p]; clp[p] = plp[p]; }
        __syncthreads();
        store_rows(cur.d_x, L.d, cx, L.DP, row0, np);
        store_rows(cur.d_grad_log_q, L.d, cgq, L.DP, row0, np);
        store_rows(cur.d_grad_log_p, L.d, cgp, L.DP, row0, np);
        for (int p = threadIdx.x; p < np; p += FAB_NT) {
            cur.d_log_q[row0 + p] = clq[p];
            cur.d_log_p[row0 + p] = clp[p];
            if (a.update_log_w) {   // ais.py:93-100
                const float inc = __fsub_rn(gamma_of(a.g_next, clq[p], clp[p]),
                                            gamma_of(a.g_w, clq[p], clp[p]));
                log_w[row0 + p] = __fadd_rn(log_w[row0 + p], inc);
            }
        }
        if (prop_out.d_x) {
            store_rows(prop_out.d_x, L.d, px, L.DP, row0, np);
            store_rows(prop_out.d_grad_log_q, L.d, pgq, L.DP, row0, np);
            store_rows(prop_out.d_grad_log_p, L.d, pgp, L.DP, row0, np);
            for (int p = threadIdx.x; p < np; p += FAB_NT) {
                prop_out.d_log_q[row0 + p] = plq[p];
                prop_out.d_log_p[row0 + p] = plp[p];
            }
        }
    } else {
        if (threadIdx.x < 2) blk[threadIdx.x] = 0.f;
    }
    __syncthreads();
    prof_mark(18);
#ifdef FAB_PROF
    if (threadIdx.x == 0 && blockIdx.x < 1024) {
        g_fab_cta_cycles[blockIdx.x] = (unsigned long long)(clock64() - cta_t0);
        unsigned int smid; asm("mov.u32 %0, %%smid;" : "=r"(smid));
        g_fab_cta_smid[blockIdx.x] = smid;
    }
#endif
    if (grid_reduce_last<2>(blk, ws, blk + 4)) {
        if (threadIdx.x == 0) {
            stats[0] = blk[4]; stats[1] = (float)n_act; stats[2] = blk[5]; stats[3] = 0.f;
            if (!a.defer_stats) hmc_finish(st, a, blk[4], (float)n_act, blk[5]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// fused Metropolis transition (metropolis.py:51-74): all n_updates in one launch
// state area: cur x, prop x : 2*[T][DP]; scalars cur_lq,cur_lp,prop_lq,prop_lp,g_prev,flag;
// blk[FAB_MAX_UPDATES*2]
// ---------------------------------------------------------------------------------------------
#define FAB_MAX_UPDATES 16
__host__ __device__ inline int metro_state_floats(int T, int DP) {
    return 2 * T * DP + 7 * 16 + 4 * FAB_MAX_UPDATES;
}

__device__ __forceinline__ void metropolis_finish(const fab_metropolis_args& a, float* scal,
                                                  int nu, float sum, float count) {
    if (!a.tune) return;
    float* s = scal + (size_t)(a.i - 1) * a.n_updates + nu;
    const float p_acc = sum / count;
    if (p_acc > a.target_p_accept) *s = __fmul_rn(*s, 1.05f);
    else *s = __fdiv_rn(*s, 1.05f);
}

__global__ void k_metropolis_finish(fab_metropolis_args a, float* scal,
                                    const float* __restrict__ stats) {
    if (threadIdx.x == 0 && blockIdx.x == 0)
        for (int u = 0; u < a.n_updates; ++u) metropolis_finish(a, scal, u, stats[2 * u], stats[2 * u + 1]);
}

template <int TP>
__global__ void __launch_bounds__(FAB_NT, FAB_MIN_CTAS)
k_metropolis(TileLayout L, fab_flow_desc f, const float* __restrict__ blob, fab_target_desc tgt,
             fab_metropolis_args a, float* scal, fab_point cur, float* __restrict__ log_w,
             const float* __restrict__ prop_noise, const float* __restrict__ unif,
             const int* __restrict__ n_active, float* __restrict__ stats, float* ws, long long n) {
    float* const smem = fab_smem;
    const TileBufs b = tile_bufs(L);
    const int T = L.T;
    const int TD = T * L.DP, T4 = 16;
    float* cx = smem + L.o_state;
    float* px = cx + TD;
    float* clq = px + TD;  float* clp = clq + T4;
    float* plq = clp + T4; float* plp = plq + T4;
    float* gprev = plp + T4; float* flag = gprev + T4; float* contrib = flag + T4;
    float* blk = contrib + T4;                    // [u] = sum_p min(a,1) ; totals at [2*MAXU + u]

    const long long n_act = n_active ? (long long)(*n_active) : n;
    const long long row0 = (long long)blockIdx.x * T;
    int np = (int)min((long long)T, n_act - row0);
    if (np < 0) np = 0;
    for (int u = threadIdx.x; u < 4 * FAB_MAX_UPDATES; u += FAB_NT) blk[u] = 0.f;
    if (np > 0) {
        tile_init(L, f, blob);
        load_rows(cx, L.DP, cur.d_x, L.d, row0, np, T);
        for (int p = threadIdx.x; p < T; p += FAB_NT) {
            const float lq = p < np ? cur.d_log_q[row0 + p] : 0.f;
            const float lp = p < np ? cur.d_log_p[row0 + p] : 0.f;
            clq[p] = lq; clp[p] = lp;
            gprev[p] = gamma_of(a.g, lq, lp);     // never refreshed
        }
        __syncthreads();
        for (int u = 0; u < a.n_updates; ++u) {
            const float sigma = scal[(size_t)(a.i - 1) * a.n_updates + u];
            const float* nz = prop_noise + (size_t)u * n * L.d;
            for (int i = threadIdx.x; i < TD; i += FAB_NT) {
                const int p = i / L.DP, j = i - p * L.DP;
                px[i] = (p < np && j < L.d)
                            ? __fadd_rn(cx[i], __fmul_rn(__ldg(nz + (row0 + p) * L.d + j), sigma))
                            : 0.f;
            }
            __syncthreads();
            eval_point<TP, false>(L, f, blob, tgt, px, plq, plp, nullptr, nullptr);
            if (threadIdx.x < T) {
                const int p = threadIdx.x;
                float fl = 0.f, c_p = 0.f;
                if (p < np) {
                    float acc = expf(__fsub_rn(gamma_of(a.g, plq[p], plp[p]), gprev[p]));
                    if (!fab_isfinite(acc)) acc = 0.f;          // nan_to_num(nan=0, +-inf=0)
                    fl = (acc > __ldg(unif + (size_t)u * n + row0 + p)) ? 1.f : 0.f;
                    c_p = fminf(acc, 1.f);
                }
                flag[p] = fl; contrib[p] = c_p;
                if (fl != 0.f) { clq[p] = plq[p]; clp[p] = plp[p]; }
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                float s = 0.f;
                for (int p = 0; p < np; ++p) s += contrib[p];
                blk[u] = s;
            }
            for (int i = threadIdx.x; i < TD; i += FAB_NT) {
                const int p = i / L.DP;
                if (flag[p] != 0.f) cx[i] = px[i];
            }
            __syncthreads();
        }
        store_rows(cur.d_x, L.d, cx, L.DP, row0, np);
        for (int p = threadIdx.x; p < np; p += FAB_NT) {
            cur.d_log_q[row0 + p] = clq[p];
            cur.d_log_p[row0 + p] = clp[p];
            if (a.update_log_w) {
                const float inc = __fsub_rn(gamma_of(a.g_next, clq[p], clp[p]),
                                            gamma_of(a.g_w, clq[p], clp[p]));
                log_w[row0 + p] = __fadd_rn(log_w[row0 + p], inc);
            }
        }
    }
    __syncthreads();
    if (grid_reduce_last<FAB_MAX_UPDATES>(blk, ws, blk + 2 * FAB_MAX_UPDATES)) {
        if (threadIdx.x == 0) {
            for (int u = 0; u < a.n_updates; ++u) {
                stats[2 * u] = blk[2 * FAB_MAX_UPDATES + u];
                stats[2 * u + 1] = (float)n_act;
                if (!a.defer_stats)
                    metropolis_finish(a, scal, u, blk[2 * FAB_MAX_UPDATES + u], (float)n_act);
            }
        }
    }
}
