// libfab_b200.so, second translation unit: C ABI of the row-tile engine (umma_engine.cuh).
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "host_util.h"
#include "reduce_finish.cuh"
#include "umma_engine.cuh"

// ---------------------------------------------------------------------------------------------
// row-tile engine (tcgen05 / TMEM / TMA): umma_engine.cuh
// ---------------------------------------------------------------------------------------------
namespace {
// FAB_UE_TRUNC="d0,d1": calibration hook (profiles/rowtile_debug.py) -- per-accumulation
// compensation for the short z-path chain (type 0 / 3 / 6) and for the long chains
ULayout layout_for(const fab_flow_desc* f) {
    ULayout L = make_ulayout(f->dim, f->width, f->n_layers);
    if (const char* tr = getenv("FAB_UE_TRUNC")) {
        float a = UE_TRUNC_PER_ACC, b = UE_TRUNC_PER_ACC;
        if (sscanf(tr, "%f,%f", &a, &b) == 2) {
            L.dl[0] = a * (float)(L.t[0].KS - 1);
            L.dl[1] = b * (float)L.t[0].KS;
            for (int i = 1; i < 7; ++i) L.dl[1 + i] = (L.t[i].KS <= 3 ? a : b) * (float)L.t[i].KS;
        }
    }
    return L;
}
bool umma_flow_ok(const fab_flow_desc* f) {
    return fab_flow_ok(f) && f->d1 == f->d2 && umma_supported(f->dim, f->width, f->n_layers);
}
template <typename K>
int launch_pairs(K kernel, int pairs, int smem_bytes, cudaStream_t s, void** args) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return fab_cuda_fail(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(UE_THREADS); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    e = cudaLaunchKernelExC(&cfg, (const void*)kernel, args);
    if (e != cudaSuccess) return fab_cuda_fail(e, "cudaLaunchKernelExC(cluster 2)");
    return FAB_OK;
}
size_t umma_mask_offset(int pairs) { return ((size_t)(16 + 2 * pairs * 2 * sizeof(float)) + 255) & ~(size_t)255; }
}  // namespace

extern "C" {

int fab_umma_supported(const fab_flow_desc* flow) { return umma_flow_ok(flow) ? 1 : 0; }

int64_t fab_umma_blob_bytes(const fab_flow_desc* flow) {
    if (!umma_flow_ok(flow)) return fab_fail(FAB_E_UNSUPPORTED, "row-tile engine: needs dim 32, width in {64,...,320} step 64, 1..10 layers");
    return make_ulayout(flow->dim, flow->width, flow->n_layers).blob_bytes;
}

/* offs[0] = total floats, [1] = float offset of layer 0, [2] = floats per layer, [3] = offset of
 * sum(log_S) inside a layer block, then for the 7 operand types (matrix offset, bias offset or -1),
 * (K, N) of each type in offs[18..31]. */
int fab_umma_plain_layout(const fab_flow_desc* flow, int64_t* offs) {
    if (!umma_flow_ok(flow) || !offs) return fab_fail(FAB_E_UNSUPPORTED, "fab_umma_plain_layout: unsupported flow shape");
    const ULayout L = make_ulayout(flow->dim, flow->width, flow->n_layers);
    offs[0] = L.plain_floats; offs[1] = L.plain_off_layers; offs[2] = L.plain_layer_floats; offs[3] = L.plain_logs_off;
    for (int i = 0; i < 7; ++i) {
        offs[4 + 2 * i] = L.t[i].plain_off;
        offs[5 + 2 * i] = L.t[i].bias ? L.t[i].plain_bias_off : -1;
        offs[18 + 2 * i] = L.t[i].Kreal;
        offs[19 + 2 * i] = L.t[i].N;
    }
    return FAB_OK;
}

int fab_umma_pack_f32(const fab_flow_desc* flow, const float* d_plain, void* d_ublob, void* stream) {
    if (!umma_flow_ok(flow) || !d_plain || !d_ublob)
        return fab_fail(FAB_E_INVALID, "fab_umma_pack_f32: bad arguments");
    const ULayout L = make_ulayout(flow->dim, flow->width, flow->n_layers);
    cudaStream_t s = (cudaStream_t)stream;
    k_umma_scales<<<7 * L.K, 256, 0, s>>>(L, d_plain, (float*)d_ublob);
    FAB_CK_LAUNCH("k_umma_scales");
    long long mx = 0;
    for (int i = 0; i < 7; ++i) mx = std::max(mx, 4LL * L.t[i].KS * L.t[i].R);
    dim3 grid((unsigned)((mx + 255) / 256), (unsigned)L.K, 7);
    k_umma_pack<<<grid, 256, 0, s>>>(L, d_plain, (uint8_t*)d_ublob);
    FAB_CK_LAUNCH("k_umma_pack");
    return FAB_OK;
}

int64_t fab_umma_workspace_bytes(const fab_flow_desc* flow, int64_t n) {
    if (!umma_flow_ok(flow) || n < 0) return FAB_E_INVALID;
    const int64_t pairs = (n + 127) / 128;
    return (int64_t)umma_mask_offset((int)pairs) + (int64_t)flow->n_layers * 24 * 4 * (pairs * 128);
}

int fab_flow_logprob_grad_umma_f32(const fab_flow_desc* flow, const void* d_ublob, const float* d_x,
                                   float* d_log_q, float* d_grad, void* d_workspace, int64_t n, void* stream) {
    if (!umma_flow_ok(flow) || !d_ublob || !d_x || !d_log_q || !d_workspace || n < 0 || !fab_aligned16(d_x) ||
        (d_grad && !fab_aligned16(d_grad)))
        return fab_fail(FAB_E_INVALID, "fab_flow_logprob_grad_umma_f32: bad arguments");
    if (n == 0) return FAB_OK;
    ULayout L = layout_for(flow);
    const int pairs = (int)((n + 127) / 128);
    const uint8_t* blob = (const uint8_t*)d_ublob;
    uint32_t* ms = (uint32_t*)((char*)d_workspace + umma_mask_offset(pairs));
    long long nn = n;
    void* args[] = {&L, &blob, &d_x, &d_log_q, &d_grad, &ms, &nn};
#define UE_CASE(N)                                                                                               \
    case N: return d_grad ? launch_pairs(k_flow_logprob_u<true, N>, pairs, L.smem_bytes, (cudaStream_t)stream, args) \
                          : launch_pairs(k_flow_logprob_u<false, N>, pairs, L.smem_bytes, (cudaStream_t)stream, args);
    switch (L.WQ / 16) { UE_CASE(1) UE_CASE(2) UE_CASE(3) UE_CASE(4) UE_CASE(5) }
#undef UE_CASE
    return fab_fail(FAB_E_UNSUPPORTED, "row-tile engine: unsupported width");
}

/* Chain initialisation (ais.py:56-65) around the row-tile engine: the flow sample on the tile engine
 * (eps -> x, forward-pass log q), log q and its input-gradient re-evaluated by the inverse pass on the
 * row-tile engine (create_point(with_grad=True)), the streaming target kernel, and
 * a one-line tail for the first log-weight and the validity flags.  Same contract as fab_ais_init_f32
 * with with_grad = 1; d_log_q0 is required. */
__global__ void k_ais_init_tail(fab_gamma g1, const float* __restrict__ lq, const float* __restrict__ lp,
                                const float* __restrict__ lq0, float* __restrict__ log_w, uint8_t* __restrict__ valid,
                                long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float q = lq[i], p = lp[i];
    log_w[i] = __fsub_rn(gamma_of(g1, q, p), lq0[i]);
    valid[i] = (fab_isfinite(q) && fab_isfinite(p)) ? 1 : 0;
}

int fab_ais_init_umma_f32(const fab_flow_desc* flow, const float* d_blob, const void* d_ublob,
                          const fab_target_desc* target, const float* d_eps, fab_gamma g1, fab_point out,
                          float* d_log_w, float* d_log_q0, uint8_t* d_valid, void* d_workspace, int64_t n,
                          void* stream) {
    if (!umma_flow_ok(flow) || !fab_target_ok(target) || target->dim != flow->dim || !d_blob || !d_ublob || !d_eps ||
        !out.d_x || !out.d_log_q || !out.d_log_p || !out.d_grad_log_q || !out.d_grad_log_p || !d_log_w || !d_log_q0 ||
        !d_valid || !d_workspace || n < 0)
        return fab_fail(FAB_E_INVALID, "fab_ais_init_umma_f32: bad arguments");
    if (n == 0) return FAB_OK;
    int e;
    if ((e = fab_flow_sample_f32(flow, d_blob, d_eps, out.d_x, d_log_q0, n, stream))) return e;
    if ((e = fab_flow_logprob_grad_umma_f32(flow, d_ublob, out.d_x, out.d_log_q, out.d_grad_log_q, d_workspace, n, stream)))
        return e;
    if ((e = fab_target_logprob_grad_f32(target, out.d_x, out.d_log_p, out.d_grad_log_p, n, stream))) return e;
    k_ais_init_tail<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g1, out.d_log_q, out.d_log_p, d_log_q0,
                                                                                d_log_w, d_valid, (long long)n);
    FAB_CK_LAUNCH("k_ais_init_tail");
    return FAB_OK;
}

int fab_hmc_step_umma_f32(const fab_flow_desc* flow, const void* d_ublob, const fab_target_desc* target,
                          fab_hmc_state st, fab_hmc_args a, fab_point cur, fab_point prop_in,
                          fab_point prop_out, float* d_log_w, const float* d_mom_noise,
                          const float* d_exp_noise, const int32_t* d_n_active, float* d_stats,
                          void* d_workspace, int64_t n, void* stream) {
    if (!umma_flow_ok(flow) || !fab_target_ok(target) || target->kind != FAB_TARGET_MANYWELL ||
        target->dim != flow->dim || !d_ublob ||
        !st.d_epsilons || !st.d_common_epsilon || !st.d_mass || !st.d_log || st.n_outer < 1 ||
        a.i < 1 || a.i > st.n_dist || a.outer < 0 || a.outer >= st.n_outer || a.L < 1 ||
        !cur.d_x || !cur.d_log_q || !cur.d_log_p || !cur.d_grad_log_q || !cur.d_grad_log_p ||
        !d_mom_noise || !d_exp_noise || !d_stats || !d_workspace || n < 0 ||
        (a.update_log_w && !d_log_w))
        return fab_fail(FAB_E_INVALID, "fab_hmc_step_umma_f32: bad arguments (many-well target, dim 32 only)");
    if (prop_in.d_x && (!prop_in.d_log_q || !prop_in.d_log_p || !prop_in.d_grad_log_q || !prop_in.d_grad_log_p))
        return fab_fail(FAB_E_INVALID, "fab_hmc_step_umma_f32: incomplete prop_in");
    if (prop_out.d_x && (!prop_out.d_log_q || !prop_out.d_log_p || !prop_out.d_grad_log_q || !prop_out.d_grad_log_p))
        return fab_fail(FAB_E_INVALID, "fab_hmc_step_umma_f32: incomplete prop_out");
    const void* al[] = {cur.d_x, cur.d_grad_log_q, cur.d_grad_log_p, prop_in.d_x, prop_in.d_grad_log_q,
                        prop_in.d_grad_log_p, prop_out.d_x, prop_out.d_grad_log_q, prop_out.d_grad_log_p,
                        d_mom_noise, st.d_mass};
    for (const void* p : al) if (p && !fab_aligned16(p)) return fab_fail(FAB_E_INVALID, "fab_hmc_step_umma_f32: pointers must be 16-byte aligned");
    if (n == 0) return FAB_OK;
    ULayout L = layout_for(flow);
    const int pairs = (int)((n + 127) / 128);
    const uint8_t* blob = (const uint8_t*)d_ublob;
    float* ws = (float*)d_workspace;
    uint32_t* ms = (uint32_t*)((char*)d_workspace + umma_mask_offset(pairs));
    fab_target_desc tg = *target;
    long long nn = n;
    void* args[] = {&L, &blob, &tg, &st, &a, &cur, &prop_in, &prop_out, &d_log_w, &d_mom_noise, &d_exp_noise,
                    &d_n_active, &d_stats, &ws, &ms, &nn};
    switch (L.WQ / 16) {
        case 1: return launch_pairs(k_hmc_step_u<1>, pairs, L.smem_bytes, (cudaStream_t)stream, args);
        case 2: return launch_pairs(k_hmc_step_u<2>, pairs, L.smem_bytes, (cudaStream_t)stream, args);
        case 3: return launch_pairs(k_hmc_step_u<3>, pairs, L.smem_bytes, (cudaStream_t)stream, args);
        case 4: return launch_pairs(k_hmc_step_u<4>, pairs, L.smem_bytes, (cudaStream_t)stream, args);
        case 5: return launch_pairs(k_hmc_step_u<5>, pairs, L.smem_bytes, (cudaStream_t)stream, args);
    }
    return fab_fail(FAB_E_UNSUPPORTED, "row-tile engine: unsupported width");
}

#ifdef UE_PROF
/* experiment builds only (-DUE_PROF): read and clear the 16 cycle counters of umma_engine.cuh */
int fab_umma_prof_read(unsigned long long* out16) {
    if (cudaDeviceSynchronize() != cudaSuccess) return FAB_E_CUDA;
    if (cudaMemcpyFromSymbol(out16, g_ue_prof, sizeof(unsigned long long) * 16) != cudaSuccess) return FAB_E_CUDA;
    unsigned long long z[16] = {};
    if (cudaMemcpyToSymbol(g_ue_prof, z, sizeof(z)) != cudaSuccess) return FAB_E_CUDA;
    return FAB_OK;
}
#endif

}  // extern "C"
