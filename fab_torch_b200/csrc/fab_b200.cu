// libfab_b200.so -- C ABI (include/fab_b200.h) over the sm_100a kernels.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "tile_kernels.cuh"
#include "param_grad.cuh"
#include "misc_kernels.cuh"
#include "buffer_kernels.cuh"
#include "host_util.h"

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
int cuda_fail(cudaError_t e, const char* what) {
    return fail(FAB_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK_LAUNCH(what)                                             \
    do {                                                            \
        cudaError_t _e = cudaGetLastError();                        \
        if (_e != cudaSuccess) return cuda_fail(_e, what);          \
    } while (0)

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

constexpr int kMaxSmemBytes = 232448;   // 227 KB opt-in dynamic shared memory per CTA on sm_100

int sm_count() {                 // per device: a process may hold several GPUs
    static int cache[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cache[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cache[dev] = n;
    }
    return cache[dev];
}

bool flow_ok(const fab_flow_desc* f) {
    return f && f->dim >= 2 && f->d1 + f->d2 == f->dim && f->d2 >= 1 && f->n_layers >= 0 &&
           (f->n_layers == 0 || f->width >= 1);
}

// Pick the particles per CTA (T) and the slot count of the MMA tile (TP = 8 or 16).  An MMA costs
// the same for 1 or 16 live slots and the chain is sequential per particle, so the fastest split
// spreads the batch over all SMs: T = ceil(n / n_SM) up to the 16 slots, then whole waves of 16.
// Very wide flows whose TP = 16 layout does not fit shared memory fall back to 8 slots.
template <typename StateFn>
int pick_tile(const fab_flow_desc& f, long long n, bool with_grad, StateFn state_floats,
              TileLayout* out) {
    static const int forced = getenv("FAB_FORCE_TILE") ? atoi(getenv("FAB_FORCE_TILE")) : 0;
    long long T = (n + sm_count() - 1) / sm_count();
    if (T > 16) T = 16;
    if (T < 1) T = 1;
    if (forced >= 1 && forced <= 16) T = forced;
    for (int attempt = 0; attempt < 2; ++attempt) {
        const int TP = T <= 8 ? 8 : 16;
        TileLayout L = make_tile_layout(f, (int)T, TP, with_grad, state_floats((int)T, fab_round4(f.dim)));
        if ((long long)L.total_floats * 4 <= kMaxSmemBytes) { *out = L; return (int)T; }
        if (T <= 8) break;
        T = 8;
    }
    return 0;
}

template <typename K>
int set_smem(K kernel, const TileLayout& L) {
    const int bytes = L.total_floats * 4;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    return FAB_OK;
}

inline int sample_state(int, int) { return 16; }

#define DISPATCH_TP(L, ...)                                      \
    switch ((L).TP) {                                            \
        case 16: { constexpr int TT = 16; __VA_ARGS__; } break;  \
        case 8:  { constexpr int TT = 8;  __VA_ARGS__; } break;  \
        default: return fail(FAB_E_INVALID, "no tile variant fits shared memory for this flow"); \
    }

}  // namespace

int fab_fail(int code, const std::string& msg) { return fail(code, msg); }
int fab_cuda_fail(cudaError_t e, const char* what) { return cuda_fail(e, what); }
bool fab_flow_ok(const fab_flow_desc* f) { return flow_ok(f); }

extern "C" {

int fab_version(void) { return 100; }
const char* fab_last_error(void) { return g_err.c_str(); }

int64_t fab_flow_desc_init(fab_flow_desc* d, int32_t dim, int32_t width, int32_t n_layers) {
    if (!d || dim < 2 || n_layers < 0 || (n_layers > 0 && width < 1))
        return fail(FAB_E_INVALID, "fab_flow_desc_init: need dim>=2, n_layers>=0, width>=1");
    std::memset(d, 0, sizeof(*d));
    d->dim = dim;
    d->d1 = (int32_t)((double)dim / 2 + 0.5);      // make_normflow_model.py:21
    d->d2 = dim - d->d1;
    d->width = n_layers > 0 ? width : 0;
    d->width_pad = n_layers > 0 ? fab_round8(width) : 0;
    d->width_kpad = n_layers > 0 ? fab_round16(width) : 0;
    d->n_layers = n_layers;
    const int64_t DP = fab_round4(dim), D8 = fab_round8(dim), D16 = fab_round16(dim),
                  D1K = fab_round16(d->d1), P8 = fab_round8(2 * d->d2), P16 = fab_round16(2 * d->d2),
                  W8 = d->width_pad, W16 = d->width_kpad;
    int64_t o = 0;
    d->off_base_loc = o; o += DP;
    d->off_base_log_scale = o; o += DP;
    d->off_layers = o;
    int64_t l = 0;
    auto frag = [](int64_t K16, int64_t N8) { return (K16 / 16) * (N8 / 8) * 128; };
    d->o_mw1 = l;     l += frag(D16, D8 + W8);
    d->o_w2 = l;      l += frag(W16, W8);
    d->o_w3 = l;      l += frag(W16, P8);
    d->o_w3t = l;     l += frag(P16, W8);
    d->o_w2t = l;     l += frag(W16, W8);
    d->o_w1mt = l;    l += frag(W16 + D16, D8);
    d->o_w1 = l;      l += frag(D1K, W8);
    d->o_mix_inv = l; l += frag(D16, D8);
    d->o_b1 = l;      l += D8 + W8;
    d->o_b2 = l;      l += W8;
    d->o_b3 = l;      l += P8;
    d->o_logs = l;    l += 4;
    d->o_b1s = l;     l += W8;
    d->o_tmix = l;    l += D8;
    d->layer_stride = l;
    d->total_floats = o + (int64_t)n_layers * l + 512;     // tail pad: L1 prefetches may run past the end
    return d->total_floats;
}

int fab_tile_particles(const fab_flow_desc* flow, int64_t n) {
    if (!flow_ok(flow) || n <= 0) return fail(FAB_E_INVALID, "fab_tile_particles: bad arguments");
    TileLayout L;
    const int T = pick_tile(*flow, n, true, hmc_state_floats, &L);
    if (T == 0) return fail(FAB_E_INVALID, "no tile variant fits shared memory for this flow");
    return T;
}

int fab_flow_sample_f32(const fab_flow_desc* flow, const float* d_blob, const float* d_eps,
                        float* d_x, float* d_log_q, int64_t n, void* stream) {
    if (!flow_ok(flow) || !d_blob || !d_eps || !d_x || !d_log_q || n < 0)
        return fail(FAB_E_INVALID, "fab_flow_sample_f32: bad arguments");
    if (n == 0) return FAB_OK;
    TileLayout L;
    const int T = pick_tile(*flow, n, false, sample_state, &L);
    if (T == 0) return fail(FAB_E_INVALID, "no tile variant fits shared memory for this flow");
    const unsigned grid = (unsigned)((n + T - 1) / T);
    DISPATCH_TP(L, {
        if (int e = set_smem(k_flow_sample<TT>, L)) return e;
        k_flow_sample<TT><<<grid, FAB_NT, L.total_floats * 4, (cudaStream_t)stream>>>(
            L, *flow, d_blob, d_eps, d_x, d_log_q, (long long)n);
    });
    CK_LAUNCH("k_flow_sample");
    return FAB_OK;
}

int fab_flow_logprob_grad_f32(const fab_flow_desc* flow, const float* d_blob, const float* d_x,
                              float* d_log_q, float* d_grad, int64_t n, void* stream) {
    if (!flow_ok(flow) || !d_blob || !d_x || !d_log_q || n < 0)
        return fail(FAB_E_INVALID, "fab_flow_logprob_grad_f32: bad arguments");
    if (n == 0) return FAB_OK;
    TileLayout L;
    const bool grad = d_grad != nullptr;
    const int T = pick_tile(*flow, n, grad, sample_state, &L);
    if (T == 0) return fail(FAB_E_INVALID, "no tile variant fits shared memory for this flow");
    const unsigned grid = (unsigned)((n + T - 1) / T);
    const size_t sm = (size_t)L.total_floats * 4;
    cudaStream_t s = (cudaStream_t)stream;
    DISPATCH_TP(L, {
        if (grad) {
            if (int e = set_smem(k_flow_logprob<TT, true>, L)) return e;
            k_flow_logprob<TT, true><<<grid, FAB_NT, sm, s>>>(L, *flow, d_blob, d_x, d_log_q, d_grad,
                                                              (long long)n);
        } else {
            if (int e = set_smem(k_flow_logprob<TT, false>, L)) return e;
            k_flow_logprob<TT, false><<<grid, FAB_NT, sm, s>>>(L, *flow, d_blob, d_x, d_log_q,
                                                               nullptr, (long long)n);
        }
    });
    CK_LAUNCH("k_flow_logprob");
    return FAB_OK;
}

/* ---- the whole single-rank HMC chain as one call (fab_ais_chain_f32) ------------------------- */
namespace {
struct ChainWs { size_t hmc, filter, part, prop, total; };
ChainWs chain_ws(const fab_flow_desc& f, int64_t n, int n_outer, int use_rowtile) {
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    ChainWs w{};
    size_t o = 0;
    w.hmc = o; o += up((size_t)(use_rowtile ? fab_umma_workspace_bytes(&f, n) : fab_hmc_workspace_bytes(&f, n)));
    w.filter = o; o += up((size_t)fab_filter_workspace_bytes(n, f.dim));
    w.part = o; o += 256;
    w.prop = o; if (n_outer > 1) o += up(2 * (size_t)n * (3 * f.dim + 2) * sizeof(float));
    w.total = o;
    return w;
}
}  // namespace

int64_t fab_ais_chain_workspace_bytes(const fab_flow_desc* flow, int64_t n, int32_t n_outer, int32_t use_rowtile) {
    if (!flow_ok(flow) || n < 0 || n_outer < 1) return fail(FAB_E_INVALID, "fab_ais_chain_workspace_bytes: bad arguments");
    if (use_rowtile && !fab_umma_supported(flow)) return fail(FAB_E_UNSUPPORTED, "row-tile engine does not cover this flow");
    return (int64_t)chain_ws(*flow, n, n_outer, use_rowtile).total;
}

int fab_ais_chain_hmc_f32(const fab_flow_desc* flow, const float* d_blob, const void* d_ublob,
                          const fab_target_desc* target, fab_hmc_state st, const fab_chain_hmc_args* a,
                          const float* d_eps, fab_point pt, float* d_log_w, float* d_log_q0, uint8_t* d_valid,
                          int32_t* d_counts, float* d_rec, float* d_stats, void* d_workspace, int64_t n,
                          void* stream) {
    if (!flow_ok(flow) || !a || !a->op_gammas || !a->w_gammas || !a->w_update || !a->d_mom || !a->d_exp ||
        a->n_dist < 1 || a->n_outer < 1 || a->n_outer != st.n_outer || a->n_dist != st.n_dist || !d_counts || !d_rec ||
        !d_workspace || n < 0 || (a->use_rowtile && !d_ublob))
        return fail(FAB_E_INVALID, "fab_ais_chain_hmc_f32: bad arguments");
    if (n == 0) return FAB_OK;
    const int M = a->n_dist, d = flow->dim;
    const ChainWs w = chain_ws(*flow, n, a->n_outer, a->use_rowtile);
    char* ws = (char*)d_workspace;
    float* part = (float*)(ws + w.part);
    int e;
    if (a->use_rowtile)
        e = fab_ais_init_umma_f32(flow, d_blob, d_ublob, target, d_eps, a->w_gammas[1], pt, d_log_w, d_log_q0, d_valid,
                                  ws + w.hmc, n, stream);
    else
        e = fab_ais_init_f32(flow, d_blob, target, d_eps, a->w_gammas[1], 1, pt, d_log_w, d_log_q0, d_valid, n, stream);
    if (e) return e;
    if ((e = fab_nan_filter_f32(pt, d_log_w, d, n, nullptr, d_counts, ws + w.filter, stream))) return e;      // "chain init"
    if (a->with_logging) {
        if ((e = fab_ess_partial_f32(pt.d_log_p, pt.d_log_q, n, d_counts, part, stream))) return e;
        if ((e = fab_ess_finalize_f32(part, 1, d_rec, stream))) return e;
    }
    fab_point prop[2] = {};
    if (a->n_outer > 1) {
        float* p = (float*)(ws + w.prop);
        for (int s = 0; s < 2; ++s) {
            prop[s].d_x = p; p += (size_t)n * d;
            prop[s].d_grad_log_q = p; p += (size_t)n * d;
            prop[s].d_grad_log_p = p; p += (size_t)n * d;
            prop[s].d_log_q = p; p += n;
            prop[s].d_log_p = p; p += n;
        }
    }
    const fab_point none = {};
    for (int j = 1; j <= M; ++j) {
        fab_point prop_in = none;
        for (int no = 0; no < a->n_outer; ++no) {
            const bool last = no == a->n_outer - 1;
            const bool upd = last && a->w_update[j] != 0;
            fab_hmc_args h{};
            h.i = j; h.outer = no; h.L = a->L; h.tune = a->tune;
            h.target_p_accept = a->target_p_accept; h.max_grad = a->max_grad;
            h.g = a->op_gammas[j];
            h.update_log_w = upd ? 1 : 0;
            h.g_w = upd ? a->w_gammas[j] : a->op_gammas[j];
            h.g_next = upd ? a->w_gammas[j + 1] : a->op_gammas[j];
            h.defer_stats = 0;
            const fab_point prop_out = last ? none : prop[no & 1];
            const float* mom = a->d_mom[j - 1] + (size_t)no * n * d;
            const float* ex = a->d_exp[j - 1] + (size_t)no * n;
            if (a->use_rowtile)
                e = fab_hmc_step_umma_f32(flow, d_ublob, target, st, h, pt, prop_in, prop_out, d_log_w, mom, ex, d_counts,
                                          d_stats, ws + w.hmc, n, stream);
            else
                e = fab_hmc_step_f32(flow, d_blob, target, st, h, pt, prop_in, prop_out, d_log_w, mom, ex, d_counts,
                                     d_stats, ws + w.hmc, n, stream);
            if (e) return e;
            prop_in = prop_out;
        }
    }
    if ((e = fab_nan_filter_f32(pt, d_log_w, d, n, d_counts, d_counts + 1, ws + w.filter, stream))) return e;  // "chain end"
    if (a->with_logging) {
        if ((e = fab_ess_partial_f32(d_log_w, nullptr, n, d_counts + 1, part, stream))) return e;
        if ((e = fab_ess_finalize_f32(part, 1, d_rec + 3, stream))) return e;
    }
    return FAB_OK;
}

/* ---- parameter gradient of sum_i g_i log q(x_i) (param_grad.cuh) --------------------------------- */
namespace {
struct PgLayout {            // dense gradient buffer: [K layers][Ga | Gb | Gc | Gd] then [dloc | dlog_scale | sum g]
    int64_t oa, ob, oc, od, layer_floats, tail_off, total, ws_floats;
    int splits, rows_per_split;
};
PgLayout pg_layout(const fab_flow_desc& f, int64_t n) {
    PgLayout P{};
    const int64_t d = f.dim, W = f.width, p2 = 2 * f.d2;
    P.oa = 0; P.ob = P.oa + (d + 1) * W; P.oc = P.ob + (d + 1) * d; P.od = P.oc + W * (W + 1);
    P.layer_floats = P.od + p2 * (W + 1);
    P.tail_off = P.layer_floats * f.n_layers;
    P.total = P.tail_off + 2 * d + 1;
    P.rows_per_split = 256;
    P.splits = (int)std::max<int64_t>(1, (n + P.rows_per_split - 1) / P.rows_per_split);
    const int64_t mx = std::max<int64_t>({(d + 1) * W, (d + 1) * d, W * (W + 1), p2 * (W + 1)});
    P.ws_floats = mx * P.splits * std::max(1, f.n_layers);
    return P;
}
}  // namespace

/* offs[20]: [0] tape floats, [1] tape row stride, [2..8] offsets of z_in, h1, h2, gparam, gh2, gh1, gv
 * in a row, [9] float offset of the final latent block, [10] gradient floats, [11] floats per layer,
 * [12..15] offsets of Ga, Gb, Gc, Gd in a layer, [16] offset of the tail, [17] workspace floats. */
int fab_flow_param_grad_layout(const fab_flow_desc* flow, int64_t n, int64_t* offs) {
    if (!flow_ok(flow) || n < 0 || !offs) return fail(FAB_E_INVALID, "fab_flow_param_grad_layout: bad arguments");
    const FabTape t = make_tape(*flow, nullptr, n);
    const PgLayout P = pg_layout(*flow, n);
    offs[0] = (int64_t)flow->n_layers * n * t.RS + n * fab_round4(flow->dim);
    offs[1] = t.RS; offs[2] = t.o_z; offs[3] = t.o_h1; offs[4] = t.o_h2; offs[5] = t.o_gpar;
    offs[6] = t.o_gh2; offs[7] = t.o_gh1; offs[8] = t.o_gv;
    offs[9] = (int64_t)flow->n_layers * n * t.RS;
    offs[10] = P.total; offs[11] = P.layer_floats; offs[12] = P.oa; offs[13] = P.ob; offs[14] = P.oc; offs[15] = P.od;
    offs[16] = P.tail_off; offs[17] = P.ws_floats;
    return FAB_OK;
}

int fab_flow_logprob_tape_f32(const fab_flow_desc* flow, const float* d_blob, const float* d_x,
                              float* d_log_q, float* d_grad, float* d_tape, int64_t n, void* stream) {
    if (!flow_ok(flow) || !d_blob || !d_x || !d_log_q || !d_tape || n < 0)
        return fail(FAB_E_INVALID, "fab_flow_logprob_tape_f32: bad arguments");
    if (n == 0) return FAB_OK;
    TileLayout L;
    const int T = pick_tile(*flow, n, true, sample_state, &L);
    if (T == 0) return fail(FAB_E_INVALID, "no tile variant fits shared memory for this flow");
    const unsigned grid = (unsigned)((n + T - 1) / T);
    const size_t sm = (size_t)L.total_floats * 4;
    const FabTape tape = make_tape(*flow, d_tape, n);
    DISPATCH_TP(L, {
        if (int e = set_smem(k_flow_tape<TT>, L)) return e;
        k_flow_tape<TT><<<grid, FAB_NT, sm, (cudaStream_t)stream>>>(L, *flow, d_blob, d_x, d_log_q, d_grad, tape,
                                                                    (long long)n);
    });
    CK_LAUNCH("k_flow_tape");
    return FAB_OK;
}

int fab_flow_param_grad_f32(const fab_flow_desc* flow, const float* d_blob, const float* d_tape,
                            const float* d_g, int64_t n, float* d_out, float* d_workspace, void* stream) {
    if (!flow_ok(flow) || !d_blob || !d_tape || !d_g || !d_out || !d_workspace || n < 1)
        return fail(FAB_E_INVALID, "fab_flow_param_grad_f32: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    const FabTape tape = make_tape(*flow, const_cast<float*>(d_tape), n);
    const PgLayout P = pg_layout(*flow, n);
    const int d = flow->dim, W = flow->width, p2 = 2 * flow->d2, K = flow->n_layers;
    auto gemm = [&](int offA, int M, int offB, int N, int64_t out_off) -> int {
        const dim3 grid((unsigned)(((M + WG_BM - 1) / WG_BM) * ((N + WG_BN - 1) / WG_BN)), (unsigned)P.splits, (unsigned)K);
        k_wgrad<<<grid, 256, 0, s>>>(tape, offA, M, offB, N, d_g, P.rows_per_split, d_workspace);
        CK_LAUNCH("k_wgrad");
        const dim3 rg((unsigned)std::min<int64_t>(((int64_t)M * N + 255) / 256, 1024), (unsigned)K);
        k_wgrad_reduce<<<rg, 256, 0, s>>>(d_workspace, P.splits, M * N, d_out, P.layer_floats, out_off);
        CK_LAUNCH("k_wgrad_reduce");
        return FAB_OK;
    };
    if (K > 0) {
        if (int e = gemm(tape.o_z, d + 1, tape.o_gh1, W, P.oa)) return e;
        if (int e = gemm(tape.o_z, d + 1, tape.o_gv, d, P.ob)) return e;
        if (int e = gemm(tape.o_gh2, W, tape.o_h1, W + 1, P.oc)) return e;
        if (int e = gemm(tape.o_gpar, p2, tape.o_h2, W + 1, P.od)) return e;
    }
    k_base_grad<<<d + 1, 256, 0, s>>>(tape.layer(K), fab_round4(d), d, d_blob + flow->off_base_loc,
                                      d_blob + flow->off_base_log_scale, d_g, (long long)n, d_out + P.tail_off);
    CK_LAUNCH("k_base_grad");
    return FAB_OK;
}

static bool target_ok(const fab_target_desc* t) {
    if (!t || t->dim < 1) return false;
    if (t->kind == FAB_TARGET_MANYWELL) return true;
    if (t->kind == FAB_TARGET_GMM)
        return t->n_mixes >= 1 && t->d_locs && t->d_scales && t->d_log_weights;
    if (t->kind == FAB_TARGET_ALDP_SURROGATE)
        return t->d_locs && t->d_scales && t->d_log_weights && (int)t->b >= 0 && (int)t->b < t->dim &&
               (int)t->c >= 0 && (int)t->c < t->dim && (int)t->b != (int)t->c;
    return false;
}

int fab_target_logprob_grad_f32(const fab_target_desc* target, const float* d_x, float* d_log_p,
                                float* d_grad, int64_t n, void* stream) {
    if (!target_ok(target) || !d_x || !d_log_p || n < 0)
        return fail(FAB_E_INVALID, "fab_target_logprob_grad_f32: bad arguments");
    if (n == 0) return FAB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (target->kind == FAB_TARGET_MANYWELL && aligned16(d_x) && (!d_grad || aligned16(d_grad)) &&
        (target->dim == 32 || target->dim == 64 || target->dim == 128)) {
        // streaming form: rows of 8*C float4s, 4 rows per thread in flight
        constexpr int U = 4;
        const int C = target->dim / 32;
        const long long rows_per_cta = (256 / (8 * C)) * U;
        const unsigned grid = (unsigned)((n + rows_per_cta - 1) / rows_per_cta);
        if (C == 1) k_target_manywell_v4<1, U><<<grid, 256, 0, st>>>(*target, (const float4*)d_x, d_log_p, (float4*)d_grad, (long long)n);
        else if (C == 2) k_target_manywell_v4<2, U><<<grid, 256, 0, st>>>(*target, (const float4*)d_x, d_log_p, (float4*)d_grad, (long long)n);
        else k_target_manywell_v4<4, U><<<grid, 256, 0, st>>>(*target, (const float4*)d_x, d_log_p, (float4*)d_grad, (long long)n);
        CK_LAUNCH("k_target_manywell_v4");
        return FAB_OK;
    }
    const int nt = 256;
    const long long threads = n * 32;
    k_target<<<(unsigned)((threads + nt - 1) / nt), nt, 0, (cudaStream_t)stream>>>(
        *target, d_x, d_log_p, d_grad, (long long)n);
    CK_LAUNCH("k_target");
    return FAB_OK;
}

int fab_ais_init_f32(const fab_flow_desc* flow, const float* d_blob, const fab_target_desc* target,
                     const float* d_eps, fab_gamma g1, int32_t with_grad, fab_point out,
                     float* d_log_w, float* d_log_q0, uint8_t* d_valid, int64_t n, void* stream) {
    if (!flow_ok(flow) || !target_ok(target) || target->dim != flow->dim || !d_blob || !d_eps ||
        !out.d_x || !out.d_log_q || !out.d_log_p || !d_log_w || !d_valid || n < 0 ||
        (with_grad && (!out.d_grad_log_q || !out.d_grad_log_p)))
        return fail(FAB_E_INVALID, "fab_ais_init_f32: bad arguments");
    if (n == 0) return FAB_OK;
    TileLayout L;
    const int T = pick_tile(*flow, n, with_grad != 0, init_state_floats, &L);
    if (T == 0) return fail(FAB_E_INVALID, "no tile variant fits shared memory for this flow");
    const unsigned grid = (unsigned)((n + T - 1) / T);
    const size_t sm = (size_t)L.total_floats * 4;
    cudaStream_t s = (cudaStream_t)stream;
    DISPATCH_TP(L, {
        if (with_grad) {
            if (int e = set_smem(k_ais_init<TT, true>, L)) return e;
            k_ais_init<TT, true><<<grid, FAB_NT, sm, s>>>(L, *flow, d_blob, *target, d_eps, g1, out,
                                                          d_log_w, d_log_q0, d_valid, (long long)n);
        } else {
            if (int e = set_smem(k_ais_init<TT, false>, L)) return e;
            k_ais_init<TT, false><<<grid, FAB_NT, sm, s>>>(L, *flow, d_blob, *target, d_eps, g1, out,
                                                           d_log_w, d_log_q0, d_valid, (long long)n);
        }
    });
    CK_LAUNCH("k_ais_init");
    return FAB_OK;
}

int64_t fab_hmc_workspace_bytes(const fab_flow_desc* flow, int64_t n) {
    (void)flow;
    if (n < 0) return FAB_E_INVALID;
    return n * 2 * (int64_t)sizeof(float) + 64;               // smallest tile (1) => most CTAs
}

int fab_hmc_step_f32(const fab_flow_desc* flow, const float* d_blob, const fab_target_desc* target,
                     fab_hmc_state st, fab_hmc_args a, fab_point cur, fab_point prop_in,
                     fab_point prop_out, float* d_log_w, const float* d_mom_noise,
                     const float* d_exp_noise, const int32_t* d_n_active, float* d_stats,
                     void* d_workspace, int64_t n, void* stream) {
    if (!flow_ok(flow) || !target_ok(target) || target->dim != flow->dim || !d_blob ||
        !st.d_epsilons || !st.d_common_epsilon || !st.d_mass || !st.d_log || st.n_outer < 1 ||
        a.i < 1 || a.i > st.n_dist || a.outer < 0 || a.outer >= st.n_outer || a.L < 1 ||
        !cur.d_x || !cur.d_log_q || !cur.d_log_p || !cur.d_grad_log_q || !cur.d_grad_log_p ||
        !d_mom_noise || !d_exp_noise || !d_stats || !d_workspace || n < 0 ||
        (a.update_log_w && !d_log_w))
        return fail(FAB_E_INVALID, "fab_hmc_step_f32: bad arguments");
    if (prop_in.d_x && (!prop_in.d_log_q || !prop_in.d_log_p || !prop_in.d_grad_log_q ||
                        !prop_in.d_grad_log_p))
        return fail(FAB_E_INVALID, "fab_hmc_step_f32: incomplete prop_in");
    if (prop_out.d_x && (!prop_out.d_log_q || !prop_out.d_log_p || !prop_out.d_grad_log_q ||
                         !prop_out.d_grad_log_p))
        return fail(FAB_E_INVALID, "fab_hmc_step_f32: incomplete prop_out");
    if (n == 0) return FAB_OK;
    TileLayout L;
    const int T = pick_tile(*flow, n, true, hmc_state_floats, &L);
    if (T == 0) return fail(FAB_E_INVALID, "no tile variant fits shared memory for this flow");
    const unsigned grid = (unsigned)((n + T - 1) / T);
    const size_t sm = (size_t)L.total_floats * 4;
    DISPATCH_TP(L, {
        if (int e = set_smem(k_hmc_step<TT>, L)) return e;
        k_hmc_step<TT><<<grid, FAB_NT, sm, (cudaStream_t)stream>>>(
            L, *flow, d_blob, *target, st, a, cur, prop_in, prop_out, d_log_w, d_mom_noise,
            d_exp_noise, d_n_active, d_stats, (float*)d_workspace, (long long)n);
    });
    CK_LAUNCH("k_hmc_step");
    return FAB_OK;
}

int fab_hmc_finish_f32(fab_hmc_state st, fab_hmc_args a, const float* d_stats, void* stream) {
    if (!st.d_epsilons || !st.d_common_epsilon || !st.d_log || !d_stats)
        return fail(FAB_E_INVALID, "fab_hmc_finish_f32: bad arguments");
    k_hmc_finish<<<1, 32, 0, (cudaStream_t)stream>>>(st, a, d_stats);
    CK_LAUNCH("k_hmc_finish");
    return FAB_OK;
}

int64_t fab_hmc_peer_buffer_bytes(int32_t world) {
    if (world < 1 || world > 32) return fail(FAB_E_INVALID, "fab_hmc_peer_buffer_bytes: world must be 1..32");
    return (int64_t)FAB_PEER_RING * world * 8 * (int64_t)sizeof(float);
}

int fab_hmc_finish_peer_f32(fab_hmc_state st, fab_hmc_args a, const float* d_stats, float* const* d_peer_bufs,
                            int32_t world, int32_t rank, uint32_t* d_seq, void* stream) {
    if (!st.d_epsilons || !st.d_common_epsilon || !st.d_log || !d_stats || !d_peer_bufs || !d_seq || world < 1 ||
        world > 32 || rank < 0 || rank >= world)
        return fail(FAB_E_INVALID, "fab_hmc_finish_peer_f32: bad arguments (world <= 32)");
    k_hmc_finish_peer<<<1, 32, 0, (cudaStream_t)stream>>>(st, a, d_stats, d_peer_bufs, world, rank, d_seq);
    CK_LAUNCH("k_hmc_finish_peer");
    return FAB_OK;
}

int64_t fab_metropolis_workspace_bytes(const fab_flow_desc* flow, int64_t n, int32_t n_updates) {
    (void)flow; (void)n_updates;
    if (n < 0) return FAB_E_INVALID;
    return n * FAB_MAX_UPDATES * (int64_t)sizeof(float) + 64;
}

int fab_metropolis_transition_f32(const fab_flow_desc* flow, const float* d_blob,
                                  const fab_target_desc* target, fab_metropolis_args a,
                                  float* d_noise_scalings, fab_point cur, float* d_log_w,
                                  const float* d_prop_noise, const float* d_unif,
                                  const int32_t* d_n_active, float* d_stats, void* d_workspace,
                                  int64_t n, void* stream) {
    if (!flow_ok(flow) || !target_ok(target) || target->dim != flow->dim || !d_blob ||
        !d_noise_scalings || a.i < 1 || a.n_updates < 1 || a.n_updates > FAB_MAX_UPDATES ||
        !cur.d_x || !cur.d_log_q || !cur.d_log_p || !d_prop_noise || !d_unif || !d_stats ||
        !d_workspace || n < 0 || (a.update_log_w && !d_log_w))
        return fail(FAB_E_INVALID, "fab_metropolis_transition_f32: bad arguments (n_updates <= 16)");
    if (n == 0) return FAB_OK;
    TileLayout L;
    const int T = pick_tile(*flow, n, false, metro_state_floats, &L);
    if (T == 0) return fail(FAB_E_INVALID, "no tile variant fits shared memory for this flow");
    const unsigned grid = (unsigned)((n + T - 1) / T);
    const size_t sm = (size_t)L.total_floats * 4;
    DISPATCH_TP(L, {
        if (int e = set_smem(k_metropolis<TT>, L)) return e;
        k_metropolis<TT><<<grid, FAB_NT, sm, (cudaStream_t)stream>>>(
            L, *flow, d_blob, *target, a, d_noise_scalings, cur, d_log_w, d_prop_noise, d_unif,
            d_n_active, d_stats, (float*)d_workspace, (long long)n);
    });
    CK_LAUNCH("k_metropolis");
    return FAB_OK;
}

int fab_metropolis_finish_f32(fab_metropolis_args a, float* d_noise_scalings, const float* d_stats,
                              void* stream) {
    if (!d_noise_scalings || !d_stats || a.n_updates < 1 || a.n_updates > FAB_MAX_UPDATES)
        return fail(FAB_E_INVALID, "fab_metropolis_finish_f32: bad arguments");
    k_metropolis_finish<<<1, 32, 0, (cudaStream_t)stream>>>(a, d_noise_scalings, d_stats);
    CK_LAUNCH("k_metropolis_finish");
    return FAB_OK;
}

int fab_logw_update_f32(fab_gamma g, fab_gamma g_next, const float* d_log_q, const float* d_log_p,
                        float* d_log_w, int64_t n, void* stream) {
    if (!d_log_q || !d_log_p || !d_log_w || n < 0)
        return fail(FAB_E_INVALID, "fab_logw_update_f32: bad arguments");
    if (n == 0) return FAB_OK;
    if (aligned16(d_log_q) && aligned16(d_log_p) && aligned16(d_log_w) && n >= 4) {
        const long long threads = (n >> 2) + (n & 3);
        k_logw_update_v4<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
            g, g_next, d_log_q, d_log_p, d_log_w, (long long)n);
        CK_LAUNCH("k_logw_update_v4");
        return FAB_OK;
    }
    k_logw_update<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        g, g_next, d_log_q, d_log_p, d_log_w, (long long)n);
    CK_LAUNCH("k_logw_update");
    return FAB_OK;
}

int64_t fab_filter_workspace_bytes(int64_t n, int32_t dim) {
    if (n < 0 || dim < 1) return FAB_E_INVALID;
    return n * 4 + 64 + n * (3 * (int64_t)dim + 3) * 4;
}

int fab_nan_filter_f32(fab_point pt, float* d_log_w, int32_t dim, int64_t n, const int32_t* d_n_in,
                       int32_t* d_n_out, void* d_workspace, void* stream) {
    if (!pt.d_x || !pt.d_log_q || !pt.d_log_p || !d_log_w || dim < 1 || n < 0 || !d_n_out ||
        !d_workspace || n > 0x7fffffffLL)
        return fail(FAB_E_INVALID, "fab_nan_filter_f32: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    int* pos = (int*)d_workspace;
    int* meta = pos + n;
    float* stage = (float*)(meta + 16);
    k_filter_scan<<<1, FAB_SCAN_NT, 0, s>>>(pt.d_log_q, pt.d_log_p, (long long)n, d_n_in, d_n_out,
                                            pos, meta);
    CK_LAUNCH("k_filter_scan");
    if (n == 0) return FAB_OK;
    const unsigned grid = (unsigned)((n * 32 + 255) / 256);
    k_filter_stage<<<grid, 256, 0, s>>>(pt, d_log_w, dim, pos, meta, stage, (long long)n);
    CK_LAUNCH("k_filter_stage");
    k_filter_unstage<<<grid, 256, 0, s>>>(pt, d_log_w, dim, meta, stage, (long long)n);
    CK_LAUNCH("k_filter_unstage");
    return FAB_OK;
}

int fab_ess_partial_f32(const float* d_log_w, const float* d_sub, int64_t n,
                        const int32_t* d_n_active, float* d_partial4, void* stream) {
    if (!d_log_w || !d_partial4 || n < 0)
        return fail(FAB_E_INVALID, "fab_ess_partial_f32: bad arguments");
    k_ess_partial<<<1, FAB_SCAN_NT, 0, (cudaStream_t)stream>>>(d_log_w, d_sub, (long long)n,
                                                               d_n_active, d_partial4);
    CK_LAUNCH("k_ess_partial");
    return FAB_OK;
}

int fab_ess_finalize_f32(const float* d_partials4, int32_t n_parts, float* d_out3, void* stream) {
    if (!d_partials4 || n_parts < 1 || !d_out3)
        return fail(FAB_E_INVALID, "fab_ess_finalize_f32: bad arguments");
    k_ess_finalize<<<1, 32, 0, (cudaStream_t)stream>>>(d_partials4, n_parts, d_out3);
    CK_LAUNCH("k_ess_finalize");
    return FAB_OK;
}

int64_t fab_resample_workspace_bytes(int64_t n) {
    if (n < 0) return FAB_E_INVALID;
    return (n + 2) * 8;
}

int fab_resample_systematic_u64(const float* d_log_w, int64_t n, uint32_t u0, int64_t* d_anc,
                                void* d_workspace, void* stream) {
    if (!d_log_w || !d_anc || !d_workspace || n < 1 || n > (1LL << 24))
        return fail(FAB_E_INVALID, "fab_resample_systematic_u64: need 1 <= n <= 2^24");
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long* cdf = (unsigned long long*)d_workspace;
    k_resample_cdf<<<1, FAB_SCAN_NT, 0, s>>>(d_log_w, (long long)n, cdf);
    CK_LAUNCH("k_resample_cdf");
    k_resample_search<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(cdf, (long long)n, u0,
                                                                   (long long*)d_anc);
    CK_LAUNCH("k_resample_search");
    return FAB_OK;
}

int fab_gather_rows_f32(const float* d_src, float* d_dst, const int64_t* d_anc, int64_t n,
                        int32_t row_floats, void* stream) {
    if (!d_src || !d_dst || !d_anc || n < 0 || row_floats < 1)
        return fail(FAB_E_INVALID, "fab_gather_rows_f32: bad arguments");
    if (n == 0) return FAB_OK;
    if (row_floats % 4 == 0 && aligned16(d_src) && aligned16(d_dst)) {
        constexpr int U = 4;
        const long long tot4 = n * (row_floats / 4);
        k_gather_rows_v4<U><<<(unsigned)((tot4 + 256 * U - 1) / (256 * U)), 256, 0, (cudaStream_t)stream>>>(
            (const float4*)d_src, (float4*)d_dst, (const long long*)d_anc, (long long)n, row_floats / 4);
        CK_LAUNCH("k_gather_rows_v4");
        return FAB_OK;
    }
    const long long tot = n * row_floats;
    k_gather_rows<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_src, d_dst, (const long long*)d_anc, (long long)n, row_floats);
    CK_LAUNCH("k_gather_rows");
    return FAB_OK;
}

int fab_buffer_add_f32(float* d_buf_x, float* d_buf_log_w, float* d_buf_log_q, int64_t max_length,
                       int32_t dim, int64_t current_index, const float* d_x, const float* d_log_w,
                       const float* d_log_q, int64_t batch, void* stream) {
    if (!d_buf_x || !d_buf_log_w || !d_buf_log_q || max_length < 1 || dim < 1 || current_index < 0 ||
        current_index >= max_length || batch < 0 || batch > max_length || (batch > 0 && (!d_x || !d_log_w || !d_log_q)))
        return fail(FAB_E_INVALID, "fab_buffer_add_f32: bad arguments (batch <= max_length)");
    if (batch == 0) return FAB_OK;
    if (dim % 4 == 0 && aligned16(d_buf_x) && aligned16(d_x)) {
        constexpr int U = 4;
        const long long tot4 = batch * (dim / 4);
        k_buffer_add_v4<U><<<(unsigned)((tot4 + 256 * U - 1) / (256 * U)), 256, 0, (cudaStream_t)stream>>>(
            (float4*)d_buf_x, d_buf_log_w, d_buf_log_q, (long long)max_length, dim / 4,
            (long long)current_index, (const float4*)d_x, d_log_w, d_log_q, (long long)batch);
        CK_LAUNCH("k_buffer_add_v4");
        return FAB_OK;
    }
    const long long tot = batch * dim;
    unsigned grid = (unsigned)((tot + 255) / 256);
    if (grid > 148u * 16u) grid = 148u * 16u;
    k_buffer_add<<<grid, 256, 0, (cudaStream_t)stream>>>(d_buf_x, d_buf_log_w, d_buf_log_q,
                                                         (long long)max_length, dim,
                                                         (long long)current_index, d_x, d_log_w,
                                                         d_log_q, (long long)batch);
    CK_LAUNCH("k_buffer_add");
    return FAB_OK;
}

// workspace: fab_sel_ctl | uint32 keys[n]
int64_t fab_buffer_topk_workspace_bytes(int64_t n) {
    if (n < 0) return FAB_E_INVALID;
    return (int64_t)((sizeof(fab_sel_ctl) + 255) & ~(size_t)255) + n * 4 + 64;
}

int fab_buffer_topk_f32(const float* d_logits, const float* d_gumbel, int64_t n, int64_t k,
                        int64_t* d_indices, void* d_workspace, void* stream) {
    if (!d_logits || !d_gumbel || !d_indices || !d_workspace || n < 1 || k < 1 || k > n)
        return fail(FAB_E_INVALID, "fab_buffer_topk_f32: need 1 <= k <= n");
    cudaStream_t s = (cudaStream_t)stream;
    fab_sel_ctl* ctl = (fab_sel_ctl*)d_workspace;
    unsigned int* keys = (unsigned int*)((char*)d_workspace + ((sizeof(fab_sel_ctl) + 255) & ~(size_t)255));
    // CTA b owns a contiguous run of `chunk` keys (a multiple of the CTA width, >= 4096 keys)
    long long nb = (n + 4095) / 4096;
    if (nb > FAB_SEL_MAXB) nb = FAB_SEL_MAXB;
    long long chunk = (n + nb - 1) / nb;
    chunk = (chunk + FAB_SEL_NT - 1) / FAB_SEL_NT * FAB_SEL_NT;
    const unsigned grid = (unsigned)((n + chunk - 1) / chunk);
    cudaError_t e = cudaMemsetAsync(ctl, 0, offsetof(fab_sel_ctl, cnt), s);
    if (e != cudaSuccess) return cuda_fail(e, "fab_buffer_topk_f32: memset");
    k_buffer_keys<<<grid, FAB_SEL_NT, 0, s>>>(d_logits, d_gumbel, (long long)n, chunk, (long long)k,
                                              keys, ctl);
    CK_LAUNCH("k_buffer_keys");
    for (int pass = 1; pass < 4; ++pass) {
        k_buffer_select_hist<<<grid, FAB_SEL_NT, 0, s>>>(keys, (long long)n, chunk, pass, ctl);
        CK_LAUNCH("k_buffer_select_hist");
    }
    k_buffer_select_count<<<grid, FAB_SEL_NT, 0, s>>>(keys, (long long)n, chunk, ctl);
    CK_LAUNCH("k_buffer_select_count");
    k_buffer_select_scatter<<<grid, FAB_SEL_NT, 0, s>>>(keys, (long long)n, chunk, ctl,
                                                        (long long*)d_indices);
    CK_LAUNCH("k_buffer_select_scatter");
    return FAB_OK;
}

int fab_buffer_adjust_f32(float* d_buf_log_w, float* d_buf_log_q, const int64_t* d_indices,
                          const float* d_log_w_adjustment, const float* d_log_q, int64_t m,
                          void* stream) {
    if (!d_buf_log_w || !d_buf_log_q || m < 0 || (m > 0 && (!d_indices || !d_log_w_adjustment || !d_log_q)))
        return fail(FAB_E_INVALID, "fab_buffer_adjust_f32: bad arguments");
    if (m == 0) return FAB_OK;
    k_buffer_adjust<<<(unsigned)((m + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_buf_log_w, d_buf_log_q, (const long long*)d_indices, d_log_w_adjustment, d_log_q, (long long)m);
    CK_LAUNCH("k_buffer_adjust");
    return FAB_OK;
}

#ifdef FAB_PROF
// experiment builds only: read (and optionally clear) the per-phase cycle counters of CTA 0
int fab_debug_prof(unsigned long long* out32, int reset) {
    if (cudaMemcpyFromSymbol(out32, g_fab_prof, sizeof(unsigned long long) * 32) != cudaSuccess)
        return FAB_E_CUDA;
    if (reset) {
        unsigned long long z[32] = {0};
        if (cudaMemcpyToSymbol(g_fab_prof, z, sizeof(z)) != cudaSuccess) return FAB_E_CUDA;
    }
    return FAB_OK;
}
int fab_debug_cta_cycles(unsigned long long* out1024, unsigned int* smid1024) {
    if (cudaMemcpyFromSymbol(out1024, g_fab_cta_cycles, sizeof(unsigned long long) * 1024) != cudaSuccess ||
        cudaMemcpyFromSymbol(smid1024, g_fab_cta_smid, sizeof(unsigned int) * 1024) != cudaSuccess)
        return FAB_E_CUDA;
    return FAB_OK;
}
#endif

}  // extern "C"

bool fab_target_ok(const fab_target_desc* t) { return target_ok(t); }
