// Cross-CTA reduction and the HMC tuner / logging update shared by the warp-level tile kernels
// (tile_kernels.cuh) and the row-tile engine (umma_engine.cuh).
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// deterministic cross-CTA reduction: every CTA deposits `NV` floats, the last one to arrive sums
// them in block order.  Workspace: one uint32 arrival counter in the FIRST 16 bytes (a fixed place:
// it must not move when the grid size changes between launches that share a workspace), then
// float partial[grid][NV].  The counter must be 0 on entry and is reset to 0 on exit.
// Returns true (block-uniform) in the last CTA, with totals[NV] (shared) filled.
// ---------------------------------------------------------------------------------------------
template <int NV>
__device__ bool grid_reduce_last(const float* mine /*shared[NV]*/, float* ws, float* totals) {
    __shared__ int s_last;
    unsigned int* counter = reinterpret_cast<unsigned int*>(ws);
    float* partial = ws + 4;
    if (threadIdx.x == 0) {
        for (int v = 0; v < NV; ++v) partial[(size_t)blockIdx.x * NV + v] = mine[v];
        __threadfence();
        const unsigned int t = atomicAdd(counter, 1u);
        s_last = (t == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    if (threadIdx.x < 32) {
        for (int v = 0; v < NV; ++v) {
            float s = 0.f;
            for (unsigned int blk = threadIdx.x; blk < gridDim.x; blk += 32)
                s += __ldcg(partial + (size_t)blk * NV + v);
            s = warp_sum(s);
            if (threadIdx.x == 0) totals[v] = s;
        }
        if (threadIdx.x == 0) *counter = 0u;
    }
    __syncthreads();
    return true;
}

// HMC tuner + logging scalars (hmc.py:157-183), executed by one thread.
__device__ __forceinline__ void hmc_finish(const fab_hmc_state& st, const fab_hmc_args& a,
                                           float sum_exp, float count, float dist_sum) {
    const float log_mean = logf(sum_exp) - logf(count);
    const float p_mean = expf(log_mean);
    if (a.i == 1) {
        st.d_log[a.outer] = p_mean;
        st.d_log[2 * st.n_outer] = dist_sum / count;
    } else if (a.i == st.n_dist) {
        st.d_log[st.n_outer + a.outer] = p_mean;
        st.d_log[2 * st.n_outer + 1] = dist_sum / count;
    }
    if (a.tune) {
        float* e = st.d_epsilons + (size_t)(a.i - 1) * st.n_outer + a.outer;
        if (log_mean > logf(a.target_p_accept)) {
            *e = __fmul_rn(*e, 1.05f);
            *st.d_common_epsilon = __fmul_rn(*st.d_common_epsilon, 1.02f);
        } else {
            *e = __fdiv_rn(*e, 1.05f);
            *st.d_common_epsilon = __fdiv_rn(*st.d_common_epsilon, 1.02f);
        }
    }
}

