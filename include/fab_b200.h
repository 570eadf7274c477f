/*
 * fab_b200.h -- C ABI of the B200-native AIS hot path (libfab_b200.so).
 *
 * The reference (lollcat/fab-torch) is pure Python and has no FFI of its own; its plugin
 * boundary is the duck-typed Python surface
 *     Distribution / TrainableDistribution   fab/types_.py:8-27, fab/trainable_distributions/base.py:4
 *     TransitionOperator                     fab/sampling_methods/transition_operators/base.py:12-85
 *     AnnealedImportanceSampler              fab/sampling_methods/ais.py:20-213
 *     Point                                  fab/sampling_methods/base.py:7-47
 * The Python classes in fab_torch_b200/ mirror that surface and bind the entry points below
 * with ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer; fp32, row-major, contiguous.
 *   - desc structs are HOST pointers, read synchronously at call time.
 *   - every call only enqueues work on `stream` (a cudaStream_t passed as void*): it never
 *     synchronises, never allocates device memory, never throws.  Return value 0 = ok,
 *     negative = FAB_E_* below; fab_last_error() gives a message for the calling thread.
 *   - workspaces are caller-owned; sizes come from the *_workspace_bytes queries.
 */
#ifndef FAB_B200_H
#define FAB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FAB_OK              0
#define FAB_E_INVALID      -1   /* bad argument / unsupported shape            */
#define FAB_E_CUDA         -2   /* a CUDA runtime call failed (launch config…) */
#define FAB_E_UNSUPPORTED  -3   /* feature not compiled in                     */

#define FAB_TARGET_MANYWELL 0
#define FAB_TARGET_GMM      1
#define FAB_TARGET_ALDP_SURROGATE 2   /* BASELINE config 5: closed-form 60-dof stand-in for the
                                         OpenMM alanine-dipeptide density (aldp.py:17-159)        */

/* ---------------------------------------------------------------------------------------
 * Packed RealNVP parameters ("flow blob").
 *
 * Architecture = the one the reference builds in experiments/make_flow/make_normflow_model.py
 * :11-30,82-96 with act_norm=False:  DiagGaussian base, then n_layers x { AffineCouplingBlock(
 * MLP[d1 -> W -> W -> 2*d2], exp scale map), InvertibleAffine(d) }.
 *
 * Every dense matrix M of out[p][n] = sum_k act[p][k] * M[k][n] is stored in MMA FRAGMENT ORDER
 * for the warp-level tensor instruction mma.sync.m16n8k8 (TF32):
 *      float4 Wf[KT2][NT][32];     KT2 = K16/16 k-tile pairs, NT = N8/8 n-tiles, 32 lanes
 *      Wf[kp][nt][lane] = { M[16kp+t][8nt+g], M[16kp+t+4][8nt+g],
 *                           M[16kp+8+t][8nt+g], M[16kp+12+t][8nt+g] }   g = lane>>2, t = lane&3
 * (0 beyond K or N), i.e. the B fragments of two consecutive k-tiles: a warp reads one coalesced
 * 512-byte line per (k-tile pair, n-tile).  K is padded to K16 = round_up(K,16), N to
 * N8 = round_up(N,8).  Values are plain fp32; the kernels split them into tf32 hi/lo parts on
 * the fly (3xTF32, fp32-grade accuracy).
 *
 * Biases are separate vectors (they initialise the accumulators).  The d x d mixing matrix of
 * each InvertibleAffine is merged into the neighbouring MLP GEMM (one GEMM + one barrier pair less
 * per layer pass in both directions); the merged products are formed in float64 on the host.
 *
 * Blob layout (float offsets):  [base block][layer 0 block][layer 1 block]...[512 floats pad]
 *   base block : loc[DP], log_scale[DP]                         DP = round_up(d,4)
 *   layer block (all offsets relative to the layer block start; layer k of the reference's
 *   flow list, k = 0 is applied first when sampling).  W1 [W,d1], W2 [W,W], W3 [2*d2,W] are the
 *   nn.Linear weights ([out,in]); W3p/b3p = W3/b3 with rows de-interleaved (first the d2 shift
 *   rows 0,2,4.., then the d2 scale rows 1,3,5..).  D8/D16 = round_up(d,8/16), W8/W16 likewise
 *   for W, P8/P16 for 2*d2, D1K = round_up(d1,16).
 *   inverse direction (log_prob):
 *     o_mw1   K16=D16     N8=D8+W8  in = z       out = [v = z@Wmix | h1pre]
 *                              M[k<d][n<d] = Wmix[k][n];  M[k<d][D8+j] = (Wmix[:, :d1] @ W1^T)[k][j]
 *     o_w2    K16=W16     N8=W8     in = h1      M[k][n] = W2[n][k]
 *     o_w3    K16=W16     N8=P8     in = h2      M[k][n] = W3p[n][k]
 *   input-gradient sweep:
 *     o_w3t   K16=P16     N8=W8     in = gparam  M[k][n] = W3p[k][n]
 *     o_w2t   K16=W16     N8=W8     in = gh2     M[k][n] = W2[k][n]
 *     o_w1mt  K16=W16+D16 N8=D8     in = [gh1 | gv]   M[k<W][n] = (W1 @ Wmix[:, :d1]^T)[k][n];
 *                                                     M[W16+i][n] = Wmix[n][i]      (g @ Wmix^T)
 *   sampling direction:
 *     o_w1    K16=D1K     N8=W8     in = z1      M[k][n] = W1[n][k]
 *     o_mix_inv K16=D16   N8=D8     in = [v1,y2] M[k][n] = Wmix^-1[k][n]
 *     (o_w2, o_w3 are shared with the inverse direction)
 *   vectors:
 *     o_b1[D8+W8] = [c (D8) | b1 + c[:d1] @ W1^T (W8)]   bias of the merged inverse-direction GEMM
 *     o_b2[W8] = b2;   o_b3[P8] = b3p;   o_logs[4] : [0] = log|det| of this layer's linear part
 *     o_b1s[W8] = b1 (sampling direction);   o_tmix[D8] = t (added after @ Wmix^-1 when sampling)
 *   ActNorm (make_normflow_model.py:28-29: z*exp(s)+t after the InvertibleAffine) is folded into
 *   the linear part by the host: Wmix := diag(exp(-s)) W, Wmix^-1 := W^-1 diag(exp(s)),
 *   c = -t @ Wmix, logs = sum(log_S) - sum(s).  Without ActNorm c = t = 0 and Wmix = W.
 * ------------------------------------------------------------------------------------- */
typedef struct fab_flow_desc {
    int32_t dim;          /* d                                   */
    int32_t d1;           /* int(d/2 + 0.5): conditioner input   */
    int32_t d2;           /* d - d1: transformed half            */
    int32_t width;        /* W  (as given by the caller)         */
    int32_t width_pad;    /* W8  = round_up(W,8)  (n-tiles)      */
    int32_t width_kpad;   /* W16 = round_up(W,16) (k-tile pairs) */
    int32_t n_layers;     /* coupling blocks; 0 = plain diagonal Gaussian */
    int64_t total_floats; /* blob size                           */
    int64_t off_base_loc, off_base_log_scale;
    int64_t off_layers, layer_stride;
    int64_t o_mw1, o_w2, o_w3;
    int64_t o_w3t, o_w2t, o_w1mt;
    int64_t o_w1, o_mix_inv;
    int64_t o_b1, o_b2, o_b3, o_logs;
    int64_t o_b1s, o_tmix;
} fab_flow_desc;

/* Fills every field of *desc from (dim, width, n_layers); returns total_floats or <0. */
int64_t fab_flow_desc_init(fab_flow_desc* desc, int32_t dim, int32_t width, int32_t n_layers);

/* ---------------------------------------------------------------------------------------
 * Targets (fab/target_distributions/many_well.py:81-90 + double_well.py:44-58; gmm.py:44-66)
 * ------------------------------------------------------------------------------------- */
typedef struct fab_target_desc {
    int32_t kind;             /* FAB_TARGET_*                                            */
    int32_t dim;
    int32_t n_mixes;          /* GMM                                                     */
    int32_t mask_below_1e4;   /* GMM: log_prob < -1e4 -> -inf (gmm.py:63-65)             */
    float   a, b, c;          /* many-well: E = a x1 + b x1^2 + c x1^4 + x2^2/2 per pair */
    float   log_norm;         /* subtracted from log_prob (log Z if `normalised`, else 0)*/
    const float* d_locs;      /* GMM [n_mixes, dim]                                      */
    const float* d_scales;    /* GMM [n_mixes, dim] diagonal of scale_tril               */
    const float* d_log_weights; /* GMM [n_mixes] log mixture weights (normalised)        */
    /* ALDP surrogate: E(x) = sum_j e_j(x_j) + a * (1 - cos(x_ia - x_ib)),  ia = (int)b, ib = (int)c,
     *   harmonic coordinate (d_log_weights[j] == 0): e_j = d_scales[j]/2 * (x_j - d_locs[j])^2
     *   torsion  coordinate (multiplicity n = d_log_weights[j] > 0):
     *                                               e_j = d_scales[j] * (1 - cos(n x_j - d_locs[j]))
     * log_prob = -E(x) - log_norm; all three tables are [dim]. */
} fab_target_desc;

/* Interpolation gamma(x) = cq*log_q + cp*log_p and grad = gq_c*grad_log_q + gp_c*grad_log_p
 * (fab/sampling_methods/base.py:76-118).  The host computes the four coefficients in float64
 * exactly as the reference does (incl. the literal `2*beta`, base.py:116) and rounds to fp32. */
typedef struct fab_gamma {
    float cq, cp;     /* density coefficients   */
    float gq, gp;     /* gradient coefficients  */
} fab_gamma;

int  fab_version(void);
const char* fab_last_error(void);
/* Number of particles one thread block carries for a batch of n (dispatch heuristic). */
int  fab_tile_particles(const fab_flow_desc* flow, int64_t n);

/* K1/K2: eps[n,d] -> x[n,d], log_q[n]     (normflows NormalizingFlow.sample via
 * fab/wrappers/normflows.py:16-18) */
int fab_flow_sample_f32(const fab_flow_desc* flow, const float* d_blob, const float* d_eps,
                        float* d_x, float* d_log_q, int64_t n, void* stream);

/* K3/K4: x[n,d] -> log_q[n] and (if d_grad != NULL) d log_q / d x [n,d]
 * (fab/wrappers/normflows.py:23-24 + torch.autograd.grad in fab/sampling_methods/base.py:50-56) */
int fab_flow_logprob_grad_f32(const fab_flow_desc* flow, const float* d_blob, const float* d_x,
                              float* d_log_q, float* d_grad, int64_t n, void* stream);

/* K5/K15: x[n,d] -> log_p[n] and (if d_grad != NULL) its closed-form gradient */
int fab_target_logprob_grad_f32(const fab_target_desc* target, const float* d_x,
                                float* d_log_p, float* d_grad, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------
 * Chain state ("Point", fab/sampling_methods/base.py:7-47) as struct-of-arrays.
 * grad_* may be NULL for value-only operators (Metropolis).
 * ------------------------------------------------------------------------------------- */
typedef struct fab_point {
    float* d_x;           /* [n,d] */
    float* d_log_q;       /* [n]   */
    float* d_log_p;       /* [n]   */
    float* d_grad_log_q;  /* [n,d] or NULL */
    float* d_grad_log_p;  /* [n,d] or NULL */
} fab_point;

/* Chain initialisation, ais.py:56-65: sample the flow from eps, build the Point (value+grad of
 * log q by the inverse pass when with_grad, else log_q := forward-pass log_q0), target value
 * (+grad), log_w = gamma_1(point) - log_q0, valid[i] = isfinite(log_p) && isfinite(log_q).
 * d_log_q0 (nullable) receives the forward-pass log q. */
int fab_ais_init_f32(const fab_flow_desc* flow, const float* d_blob, const fab_target_desc* target,
                     const float* d_eps, fab_gamma g1, int32_t with_grad,
                     fab_point out, float* d_log_w, float* d_log_q0, uint8_t* d_valid,
                     int64_t n, void* stream);

/* HMC state that lives on the device (hmc.py:36-39 registered buffers + logging scalars). */
typedef struct fab_hmc_state {
    float* d_epsilons;        /* [M, n_outer]   per-distribution step sizes            */
    float* d_common_epsilon;  /* [1]            shared step-size component             */
    const float* d_mass;      /* [d]            mass vector                            */
    float* d_log;             /* [4*n_outer + 2]: first_p_accept[n_outer], last_p_accept[n_outer],
                                 avg_dist_first, avg_dist_last, spare...               */
    int32_t n_dist;           /* M                                                     */
    int32_t n_outer;
} fab_hmc_state;

typedef struct fab_hmc_args {
    int32_t i;               /* intermediate distribution index, 1..M (hmc.py:186)     */
    int32_t outer;           /* which outer step n this launch performs                */
    int32_t L;               /* leapfrog steps                                         */
    int32_t tune;            /* !eval_mode: apply adjust_step_size_p_accept (hmc.py:162-170) */
    float   target_p_accept;
    float   max_grad;        /* clamp for grad U (hmc.py:194-199)                      */
    fab_gamma g;             /* gamma at beta_i                                        */
    int32_t update_log_w;    /* after the accept: log_w += g_next(x) - g_w(x) (ais.py:93-100) */
    fab_gamma g_w;           /* the SAMPLER's gamma at beta_i (its alpha/p_target may differ
                                from the operator's, ais.py:94-98 vs hmc.py:188-191)   */
    fab_gamma g_next;        /* the sampler's gamma at beta_{i+1}                      */
    int32_t defer_stats;     /* 1: only write the local (sum_exp, count, dist) triple to d_stats
                                and leave tuner/logging to fab_hmc_finish_f32 (multi-GPU)     */
} fab_hmc_args;

int64_t fab_hmc_workspace_bytes(const fab_flow_desc* flow, int64_t n);

/* One HMC outer step (momentum draw .. accept/overwrite .. tuner) for all n particles, fused in
 * one launch: hmc.py:129-160.  `cur` is updated in place.  For n_outer>1 the reference carries
 * the *proposal* into the next outer step: pass prop_in = previous prop_out
 * (NULL d_x = start from cur) and a prop_out buffer (NULL d_x = discard).
 * d_mom_noise [n,d] ~ N(0,1), d_exp_noise [n] ~ Exp(1).  d_n_active (nullable) = device count of
 * live particles (<= n) after an on-device NaN filter. */
int fab_hmc_step_f32(const fab_flow_desc* flow, const float* d_blob, const fab_target_desc* target,
                     fab_hmc_state st, fab_hmc_args args,
                     fab_point cur, fab_point prop_in, fab_point prop_out,
                     float* d_log_w, const float* d_mom_noise, const float* d_exp_noise,
                     const int32_t* d_n_active, float* d_stats /* [4] */,
                     void* d_workspace, int64_t n, void* stream);

/* Multi-GPU tail of an outer step: d_stats holds the all-reduced (sum_exp, count, dist_sum). */
int fab_hmc_finish_f32(fab_hmc_state st, fab_hmc_args args, const float* d_stats, void* stream);
/* The same update with the statistics summed over the ranks INSIDE the launch, over peer-mapped
 * memory (NVLink) instead of an NCCL all-reduce between two launches: d_peer_bufs[r] = rank r's
 * symmetric exchange buffer (fab_hmc_peer_buffer_bytes(world) bytes, zero on first use) as mapped in
 * this process, d_seq = a device-resident exchange counter (zero on first use; every rank must issue
 * the same sequence of calls).  Capturable in a CUDA graph.  Replaces hmc.py:122-123,162-170 for a
 * rank-sharded batch. */
int64_t fab_hmc_peer_buffer_bytes(int32_t world);
int fab_hmc_finish_peer_f32(fab_hmc_state st, fab_hmc_args args, const float* d_stats,
                            float* const* d_peer_bufs, int32_t world, int32_t rank, uint32_t* d_seq,
                            void* stream);

/* Metropolis (metropolis.py:51-74): all n_updates of one transition in one launch. */
typedef struct fab_metropolis_args {
    int32_t i;               /* 1..M                                                   */
    int32_t n_updates;
    int32_t tune;            /* adjust_step_size && !eval_mode                         */
    float   target_p_accept;
    fab_gamma g;
    int32_t update_log_w;
    fab_gamma g_w;           /* sampler's gamma at beta_i / beta_{i+1}, as in fab_hmc_args */
    fab_gamma g_next;
    int32_t defer_stats;
} fab_metropolis_args;

int64_t fab_metropolis_workspace_bytes(const fab_flow_desc* flow, int64_t n, int32_t n_updates);

/* d_noise_scalings [M, n_updates] (updated in place when tuning), d_prop_noise [n_updates,n,d],
 * d_unif [n_updates,n], d_stats [2*n_updates] = per update (sum min(a,1), count). */
int fab_metropolis_transition_f32(const fab_flow_desc* flow, const float* d_blob,
                                  const fab_target_desc* target, fab_metropolis_args args,
                                  float* d_noise_scalings, fab_point cur, float* d_log_w,
                                  const float* d_prop_noise, const float* d_unif,
                                  const int32_t* d_n_active, float* d_stats,
                                  void* d_workspace, int64_t n, void* stream);
int fab_metropolis_finish_f32(fab_metropolis_args args, float* d_noise_scalings,
                              const float* d_stats, void* stream);

/* K11 alone (operator-level insertion keeps the reference's own loop): log_w += gn(x) - g(x). */
int fab_logw_update_f32(fab_gamma g, fab_gamma g_next, const float* d_log_q, const float* d_log_p,
                        float* d_log_w, int64_t n, void* stream);

/* K12 (ais.py:190-213): stable compaction of the Point + log_w + any extra [n] / [n,d] rows by
 * valid[i] = isfinite(log_p[i]) && isfinite(log_q[i]); writes the live count to d_n_out.
 * In place; a no-op copy when everything is valid. */
int64_t fab_filter_workspace_bytes(int64_t n, int32_t dim);
int fab_nan_filter_f32(fab_point pt, float* d_log_w, int32_t dim, int64_t n,
                       const int32_t* d_n_in, int32_t* d_n_out, void* d_workspace, void* stream);

/* K13 (numerical.py:18-23, ais.py:80-86): partial = (max, sum e^(lw-max), sum e^(2(lw-max)), count)
 * over this rank's live particles; finalize combines n_parts such quadruples (one per rank after
 * an all-gather) into out[0]=ESS, out[1]=logsumexp(log_w), out[2]=total count. */
int fab_ess_partial_f32(const float* d_log_w, const float* d_sub /* nullable: uses lw - sub */,
                        int64_t n, const int32_t* d_n_active, float* d_partial4, void* stream);
int fab_ess_finalize_f32(const float* d_partials4, int32_t n_parts, float* d_out3, void* stream);

/* R (build extension, oracle/resample.py): systematic resampling with a fixed-point CDF.
 * d_anc[n] = ancestor index for position (k + u0/2^32)/n.  Bit-exact vs the oracle. */
int64_t fab_resample_workspace_bytes(int64_t n);
int fab_resample_systematic_u64(const float* d_log_w, int64_t n, uint32_t u0, int64_t* d_anc,
                                void* d_workspace, void* stream);
/* Gather rows of a Point by ancestor index: dst[k] = src[anc[k]]. */
int fab_gather_rows_f32(const float* d_src, float* d_dst, const int64_t* d_anc, int64_t n,
                        int32_t row_floats, void* stream);

/* ---------------------------------------------------------------------------------------
 * Prioritised replay buffer (fab/utils/prioritised_replay_buffer.py).
 * The buffer lives in caller-owned device arrays x[max_length,dim], log_w[max_length],
 * log_q_old[max_length].
 * ------------------------------------------------------------------------------------- */
/* add (:71-85): ring write of `batch` rows starting at current_index (batch <= max_length). */
int fab_buffer_add_f32(float* d_buf_x, float* d_buf_log_w, float* d_buf_log_q, int64_t max_length,
                       int32_t dim, int64_t current_index, const float* d_x, const float* d_log_w,
                       const float* d_log_q, int64_t batch, void* stream);
/* sample without replacement (:10-17, :88-100): indices of the k largest (gumbel[i] + logits[i]),
 * i < n, written in ascending index order (the reference returns them in an unspecified order and
 * then permutes them); exact radix select, ties towards the lower index, NaN above +inf like
 * torch.topk.  d_gumbel = standard Gumbel noise drawn by the caller. */
int64_t fab_buffer_topk_workspace_bytes(int64_t n);
int fab_buffer_topk_f32(const float* d_logits, const float* d_gumbel, int64_t n, int64_t k,
                        int64_t* d_indices, void* d_workspace, void* stream);
/* adjust (:117-131): finite (adjustment, log_q) -> log_w[idx] += adjustment, log_q_old[idx] = log_q;
 * otherwise log_w[idx] = -inf.  Indices must be unique (sampling without replacement). */
int fab_buffer_adjust_f32(float* d_buf_log_w, float* d_buf_log_q, const int64_t* d_indices,
                          const float* d_log_w_adjustment, const float* d_log_q, int64_t m,
                          void* stream);

/* ---------------------------------------------------------------------------------------
 * The whole chain in one call: AnnealedImportanceSampler.sample_and_log_weights (ais.py:53-87) with a
 * HamiltonianMonteCarlo operator on ONE rank -- chain init (K1-K6, K11, K12), NaN/inf filter
 * (:65, :190-213), ESS over the base weights (:68-71), the M fused transitions with the log-weight
 * update (:74-75, :90-105), filter (:77), ESS and logsumexp of the final weights (:80-86).  Nothing
 * but launches on `stream`: no allocation, no synchronisation (capturable in a CUDA graph).
 * Host arrays: op_gammas / w_gammas [M+2] = gamma at beta_0..beta_{M+1} with the operator's and the
 * sampler's (alpha, p_target) (they may differ, ais.py:94-98 vs hmc.py:188-191); w_update [M+1]:
 * [j] != 0 iff beta_{j+1} != beta_j; d_mom / d_exp [M]: device pointers to the noise of transition
 * j+1 ([n_outer, n, d] ~ N(0,1), [n_outer, n] ~ Exp(1)).
 * Outputs: pt (live particles compacted to the front), d_log_w, d_log_q0 (forward-pass log q),
 * d_counts[2] = live particles after chain init / at the chain end, d_rec[8]: [0..2] ESS,
 * logsumexp, count of the base weights, [3..5] the same for the final weights (with_logging).
 * Multi-rank chains stay with the caller (collectives between the launches).
 * ------------------------------------------------------------------------------------- */
typedef struct fab_chain_hmc_args {
    int32_t n_dist, n_outer, L, tune, with_logging, use_rowtile;
    float   target_p_accept, max_grad;
    const fab_gamma* op_gammas;
    const fab_gamma* w_gammas;
    const uint8_t*   w_update;
    const float* const* d_mom;
    const float* const* d_exp;
} fab_chain_hmc_args;
int64_t fab_ais_chain_workspace_bytes(const fab_flow_desc* flow, int64_t n, int32_t n_outer, int32_t use_rowtile);
int fab_ais_chain_hmc_f32(const fab_flow_desc* flow, const float* d_blob, const void* d_ublob /* row-tile images or NULL */,
                          const fab_target_desc* target, fab_hmc_state st, const fab_chain_hmc_args* args,
                          const float* d_eps, fab_point pt, float* d_log_w, float* d_log_q0, uint8_t* d_valid,
                          int32_t* d_counts, float* d_rec, float* d_stats, void* d_workspace, int64_t n,
                          void* stream);

/* ---------------------------------------------------------------------------------------
 * Parameter gradient of  sum_i g_i log q_theta(x_i)  -- the theta-gradient of the FAB loss
 * (fab/core.py:112-118: loss = -mean(softmax(log_w) * log q(x)); minibatch loop
 * fab/train_with_prioritised_buffer.py:158-186), replacing loss.backward() through
 * fab/wrappers/normflows.py:20-24.
 *   fab_flow_logprob_tape_f32: log q, optionally d log q / dx, and the activation tape (per layer
 *     and particle: z_in | 1 | h1 | 1 | h2 | 1 | gparam | gh2 | gh1 | gv, see csrc/flow_tile.cuh).
 *   fab_flow_param_grad_f32: weight gradients as batch-contraction GEMMs over the tape, written as
 *     dense blocks per layer
 *         Ga [(d+1) x W]   rows 0..d-1: dM1 = d/d(Wmix[:, :d1] W1^T), row d: d/d b1
 *         Gb [(d+1) x d]   rows 0..d-1: direct part of d/d Wmix (full: Gb + [dM1 W1 | 0]); row d: d/d c
 *                          (the bias of v = z @ Wmix + c: the folded ActNorm shift)
 *         Gc [W x (W+1)]   d/d W2 ([out][in]) | d/d b2
 *         Gd [2 d2 x (W+1)] d/d W3 (rows: the d2 shifts, then the d2 scales) | d/d b3
 *     then [d/d loc (d) | d/d log_scale (d) | sum_i g_i (= d/d sum(log_S) of every layer)].
 *     d/d W1 = dM1^T Wmix[:, :d1] and the LU parameters of Wmix follow in parameter space (host).
 * Deterministic (fixed batch slices, added in order).  fab_flow_param_grad_layout reports all sizes.
 * ------------------------------------------------------------------------------------- */
int fab_flow_param_grad_layout(const fab_flow_desc* flow, int64_t n, int64_t* offs /* [20] */);
int fab_flow_logprob_tape_f32(const fab_flow_desc* flow, const float* d_blob, const float* d_x,
                              float* d_log_q, float* d_grad /* may be NULL */, float* d_tape,
                              int64_t n, void* stream);
int fab_flow_param_grad_f32(const fab_flow_desc* flow, const float* d_blob, const float* d_tape,
                            const float* d_g, int64_t n, float* d_out, float* d_workspace,
                            void* stream);

/* ---------------------------------------------------------------------------------------
 * Row-tile engine: the same flow evaluation (fab/wrappers/normflows.py:20-24 log_prob, its input
 * gradient as used by fab/sampling_methods/base.py:50-56) and the same fused HMC outer step
 * (transition_operators/hmc.py:129-160) on the 5th-generation tensor cores: tcgen05.mma with
 * cta_group::2 (a CTA pair carries 128 particles), accumulators in tensor memory, weight stages
 * streamed into shared memory by the TMA engine (cp.async.bulk).  Covers dim = 32 with
 * d1 = d2 = 16, width in {64,128,...,320}, <= 10 coupling layers (BASELINE configs 2 and 3) and the
 * many-well target; everything else stays on the warp-level engine above.
 *
 * Weight images ("ublob", bytes): [scalar block | layer 0 | layer 1 | ...].  Scalar block (fp32):
 * loc[32], log_scale[32], then per layer 16 floats: [0..6] the un-scaling factor 1/s_t of the seven
 * operand types, [7] sum(log_S), [8..11] max_n ||M[:,n]||_2 of the wide types 0 (hidden columns), 1,
 * 3, 4 and [12..13] max_n |bias_n| of types 0, 1 (the a-priori row scales of the hidden operands:
 * |h_n| <= ||h_in||_2 ||M[:,n]||_2 + |b_n|).  A layer block holds, for each operand type and each CTA
 * rank of the pair, the f16 image the tensor core reads, column block by column block, inside a block
 * one slab per k-step in the order the issuer consumes them:
 *     slab[2 k-chunks][rows of the block][8 halves]   (one "core matrix" = 8 rows x 16 bytes, no swizzle)
 * k-chunk c = logical k 8c..8c+7 (k == K: the bias row; beyond: zero).  Consumption order of the
 * k-steps of a type whose A operand is a hidden activation (1, 2, 4, 5): first half of the real
 * k-steps, the bias step (types 1, 2), second half; other types: natural order.  Rows of a block:
 *   wide types  (0: z->[h1pre|v], 1: h1->h2pre, 3: gparam->gh2, 4: gh2->gh1), column block g = 0,1:
 *     N/4 rows "hi" then N/4 rows "lo" of the output columns [q N/4, (q+1) N/4), q = 2g + rank
 *     (type 0: W/4 hidden columns, then the 8 v columns 8q..8q+7);
 *   narrow types (2: h2->[shift|scale], 5: gh1->g, 6: gv->g), one block: N/2 rows "hi" then N/2 rows
 *     "lo" of columns 8 rank + (0..7), then 16 + 8 rank + (0..7).
 * value = hi + lo = M[k][n] * s_t, s_t = the power of two that puts max|M| in [2^13, 2^14).
 * The images are produced on the device by fab_umma_pack_f32 from a plain fp32 buffer:
 * [loc | log_scale | per layer: M_0 [K_0][N_0], bias_0 [N_0], M_1, bias_1, M_2, bias_2, M_3 .. M_6,
 * sum(log_S), 3 pad]; fab_umma_plain_layout reports the offsets.
 * ------------------------------------------------------------------------------------- */
int     fab_umma_supported(const fab_flow_desc* flow);       /* 1 if this flow shape is covered */
int64_t fab_umma_blob_bytes(const fab_flow_desc* flow);
/* offs[32]: [0] total floats of the plain buffer, [1] offset of layer 0, [2] floats per layer,
 * [3] offset of sum(log_S) in a layer, [4+2t],[5+2t] matrix / bias offset of type t (-1: no bias),
 * [18+2t],[19+2t] its (K, N). */
int     fab_umma_plain_layout(const fab_flow_desc* flow, int64_t* offs);
int     fab_umma_pack_f32(const fab_flow_desc* flow, const float* d_plain, void* d_ublob, void* stream);
/* caller-owned scratch: cross-CTA reduction slots + ReLU masks of the input-gradient sweep
 * (960 bytes per particle, L2 resident).  Must be zero on first use. */
int64_t fab_umma_workspace_bytes(const fab_flow_desc* flow, int64_t n);
/* = fab_flow_logprob_grad_f32 (d_grad may be NULL); d_x / d_grad 16-byte aligned. */
int fab_flow_logprob_grad_umma_f32(const fab_flow_desc* flow, const void* d_ublob, const float* d_x,
                                   float* d_log_q, float* d_grad, void* d_workspace, int64_t n,
                                   void* stream);
/* = fab_ais_init_f32 with with_grad = 1 (d_log_q0 required): the flow sample on the tile engine, the
 * inverse-pass log q + input-gradient on the row-tile engine, streaming target kernel, log-weight /
 * validity tail.  d_workspace: fab_umma_workspace_bytes(flow, n). */
int fab_ais_init_umma_f32(const fab_flow_desc* flow, const float* d_blob, const void* d_ublob,
                          const fab_target_desc* target, const float* d_eps, fab_gamma g1, fab_point out,
                          float* d_log_w, float* d_log_q0, uint8_t* d_valid, void* d_workspace, int64_t n,
                          void* stream);
/* = fab_hmc_step_f32 (same arguments and semantics); many-well target only. */
int fab_hmc_step_umma_f32(const fab_flow_desc* flow, const void* d_ublob, const fab_target_desc* target,
                          fab_hmc_state st, fab_hmc_args args, fab_point cur, fab_point prop_in,
                          fab_point prop_out, float* d_log_w, const float* d_mom_noise,
                          const float* d_exp_noise, const int32_t* d_n_active, float* d_stats,
                          void* d_workspace, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FAB_B200_H */
