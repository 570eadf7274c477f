"""ctypes binding of libfab_b200.so (C ABI in include/fab_b200.h).

The CUDA library is the product: there is no CPU or PyTorch fallback for the hot path.  If the
shared object is missing, `lib()` raises with the build command (python -c "import
__graft_entry__ as g; g.build()").
"""
import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# FAB_B200_LIB: experiment hook (A/B runs of two builds inside one GPU call, profiles/)
LIB_PATH = os.environ.get("FAB_B200_LIB") or os.path.join(_HERE, "csrc", "libfab_b200.so")

FAB_TARGET_MANYWELL = 0
FAB_TARGET_GMM = 1
FAB_TARGET_ALDP_SURROGATE = 2
FAB_MAX_UPDATES = 16


class FlowDesc(C.Structure):
    _fields_ = [("dim", C.c_int32), ("d1", C.c_int32), ("d2", C.c_int32), ("width", C.c_int32),
                ("width_pad", C.c_int32), ("width_kpad", C.c_int32), ("n_layers", C.c_int32),
                ("total_floats", C.c_int64),
                ("off_base_loc", C.c_int64), ("off_base_log_scale", C.c_int64),
                ("off_layers", C.c_int64), ("layer_stride", C.c_int64),
                ("o_mw1", C.c_int64), ("o_w2", C.c_int64), ("o_w3", C.c_int64),
                ("o_w3t", C.c_int64), ("o_w2t", C.c_int64), ("o_w1mt", C.c_int64),
                ("o_w1", C.c_int64), ("o_mix_inv", C.c_int64),
                ("o_b1", C.c_int64), ("o_b2", C.c_int64), ("o_b3", C.c_int64), ("o_logs", C.c_int64),
                ("o_b1s", C.c_int64), ("o_tmix", C.c_int64)]


class TargetDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("dim", C.c_int32), ("n_mixes", C.c_int32),
                ("mask_below_1e4", C.c_int32),
                ("a", C.c_float), ("b", C.c_float), ("c", C.c_float), ("log_norm", C.c_float),
                ("d_locs", C.c_void_p), ("d_scales", C.c_void_p), ("d_log_weights", C.c_void_p)]


class Gamma(C.Structure):
    _fields_ = [("cq", C.c_float), ("cp", C.c_float), ("gq", C.c_float), ("gp", C.c_float)]


class PointPtrs(C.Structure):
    _fields_ = [("d_x", C.c_void_p), ("d_log_q", C.c_void_p), ("d_log_p", C.c_void_p),
                ("d_grad_log_q", C.c_void_p), ("d_grad_log_p", C.c_void_p)]


class HmcState(C.Structure):
    _fields_ = [("d_epsilons", C.c_void_p), ("d_common_epsilon", C.c_void_p),
                ("d_mass", C.c_void_p), ("d_log", C.c_void_p),
                ("n_dist", C.c_int32), ("n_outer", C.c_int32)]


class HmcArgs(C.Structure):
    _fields_ = [("i", C.c_int32), ("outer", C.c_int32), ("L", C.c_int32), ("tune", C.c_int32),
                ("target_p_accept", C.c_float), ("max_grad", C.c_float),
                ("g", Gamma), ("update_log_w", C.c_int32), ("g_w", Gamma), ("g_next", Gamma),
                ("defer_stats", C.c_int32)]


class ChainHmcArgs(C.Structure):
    _fields_ = [("n_dist", C.c_int32), ("n_outer", C.c_int32), ("L", C.c_int32), ("tune", C.c_int32),
                ("with_logging", C.c_int32), ("use_rowtile", C.c_int32),
                ("target_p_accept", C.c_float), ("max_grad", C.c_float),
                ("op_gammas", C.POINTER(Gamma)), ("w_gammas", C.POINTER(Gamma)),
                ("w_update", C.POINTER(C.c_uint8)),
                ("d_mom", C.POINTER(C.c_void_p)), ("d_exp", C.POINTER(C.c_void_p))]


class MetropolisArgs(C.Structure):
    _fields_ = [("i", C.c_int32), ("n_updates", C.c_int32), ("tune", C.c_int32),
                ("target_p_accept", C.c_float), ("g", Gamma), ("update_log_w", C.c_int32),
                ("g_w", Gamma), ("g_next", Gamma), ("defer_stats", C.c_int32)]


# name -> (restype, argtypes); every symbol declared in include/fab_b200.h
_P = C.c_void_p
SIGNATURES = {
    "fab_version": (C.c_int, []),
    "fab_last_error": (C.c_char_p, []),
    "fab_flow_desc_init": (C.c_int64, [C.POINTER(FlowDesc), C.c_int32, C.c_int32, C.c_int32]),
    "fab_tile_particles": (C.c_int, [C.POINTER(FlowDesc), C.c_int64]),
    "fab_flow_sample_f32": (C.c_int, [C.POINTER(FlowDesc), _P, _P, _P, _P, C.c_int64, _P]),
    "fab_flow_logprob_grad_f32": (C.c_int, [C.POINTER(FlowDesc), _P, _P, _P, _P, C.c_int64, _P]),
    "fab_target_logprob_grad_f32": (C.c_int, [C.POINTER(TargetDesc), _P, _P, _P, C.c_int64, _P]),
    "fab_ais_init_f32": (C.c_int, [C.POINTER(FlowDesc), _P, C.POINTER(TargetDesc), _P, Gamma,
                                   C.c_int32, PointPtrs, _P, _P, _P, C.c_int64, _P]),
    "fab_hmc_workspace_bytes": (C.c_int64, [C.POINTER(FlowDesc), C.c_int64]),
    "fab_hmc_step_f32": (C.c_int, [C.POINTER(FlowDesc), _P, C.POINTER(TargetDesc), HmcState,
                                   HmcArgs, PointPtrs, PointPtrs, PointPtrs, _P, _P, _P, _P, _P,
                                   _P, C.c_int64, _P]),
    "fab_hmc_finish_f32": (C.c_int, [HmcState, HmcArgs, _P, _P]),
    "fab_metropolis_workspace_bytes": (C.c_int64, [C.POINTER(FlowDesc), C.c_int64, C.c_int32]),
    "fab_metropolis_transition_f32": (C.c_int, [C.POINTER(FlowDesc), _P, C.POINTER(TargetDesc),
                                                MetropolisArgs, _P, PointPtrs, _P, _P, _P, _P, _P,
                                                _P, C.c_int64, _P]),
    "fab_metropolis_finish_f32": (C.c_int, [MetropolisArgs, _P, _P, _P]),
    "fab_logw_update_f32": (C.c_int, [Gamma, Gamma, _P, _P, _P, C.c_int64, _P]),
    "fab_filter_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int32]),
    "fab_nan_filter_f32": (C.c_int, [PointPtrs, _P, C.c_int32, C.c_int64, _P, _P, _P, _P]),
    "fab_ess_partial_f32": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P]),
    "fab_ess_finalize_f32": (C.c_int, [_P, C.c_int32, _P, _P]),
    "fab_resample_workspace_bytes": (C.c_int64, [C.c_int64]),
    "fab_resample_systematic_u64": (C.c_int, [_P, C.c_int64, C.c_uint32, _P, _P, _P]),
    "fab_gather_rows_f32": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, _P]),
    "fab_buffer_add_f32": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, C.c_int64, _P, _P, _P,
                                     C.c_int64, _P]),
    "fab_buffer_topk_workspace_bytes": (C.c_int64, [C.c_int64]),
    "fab_buffer_topk_f32": (C.c_int, [_P, _P, C.c_int64, C.c_int64, _P, _P, _P]),
    "fab_buffer_adjust_f32": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, _P]),
    # row-tile engine (tcgen05 / TMEM / TMA)
    "fab_ais_chain_workspace_bytes": (C.c_int64, [C.POINTER(FlowDesc), C.c_int64, C.c_int32, C.c_int32]),
    "fab_ais_chain_hmc_f32": (C.c_int, [C.POINTER(FlowDesc), _P, _P, C.POINTER(TargetDesc), HmcState,
                                        C.POINTER(ChainHmcArgs), _P, PointPtrs, _P, _P, _P, _P, _P, _P, _P,
                                        C.c_int64, _P]),
    "fab_hmc_peer_buffer_bytes": (C.c_int64, [C.c_int32]),
    "fab_hmc_finish_peer_f32": (C.c_int, [HmcState, HmcArgs, _P, _P, C.c_int32, C.c_int32, _P, _P]),
    "fab_flow_param_grad_layout": (C.c_int, [C.POINTER(FlowDesc), C.c_int64, C.POINTER(C.c_int64)]),
    "fab_flow_logprob_tape_f32": (C.c_int, [C.POINTER(FlowDesc), _P, _P, _P, _P, _P, C.c_int64, _P]),
    "fab_flow_param_grad_f32": (C.c_int, [C.POINTER(FlowDesc), _P, _P, _P, C.c_int64, _P, _P, _P]),
    "fab_umma_supported": (C.c_int, [C.POINTER(FlowDesc)]),
    "fab_umma_blob_bytes": (C.c_int64, [C.POINTER(FlowDesc)]),
    "fab_umma_plain_layout": (C.c_int, [C.POINTER(FlowDesc), C.POINTER(C.c_int64)]),
    "fab_umma_pack_f32": (C.c_int, [C.POINTER(FlowDesc), _P, _P, _P]),
    "fab_umma_workspace_bytes": (C.c_int64, [C.POINTER(FlowDesc), C.c_int64]),
    "fab_flow_logprob_grad_umma_f32": (C.c_int, [C.POINTER(FlowDesc), _P, _P, _P, _P, _P, C.c_int64, _P]),
    "fab_ais_init_umma_f32": (C.c_int, [C.POINTER(FlowDesc), _P, _P, C.POINTER(TargetDesc), _P, Gamma, PointPtrs,
                                        _P, _P, _P, _P, C.c_int64, _P]),
    "fab_hmc_step_umma_f32": (C.c_int, [C.POINTER(FlowDesc), _P, C.POINTER(TargetDesc), HmcState,
                                        HmcArgs, PointPtrs, PointPtrs, PointPtrs, _P, _P, _P, _P, _P,
                                        _P, C.c_int64, _P]),
}


def engine_choice() -> str:
    """FAB_ENGINE = auto (default) | rowtile | warp.  `auto` takes the row-tile (tcgen05) engine
    when the flow shape is covered and the batch holds at least FAB_ROWTILE_MIN_N particles
    (default 1024).  Both engines are latency-bound per tile -- 0.67 vs 1.45 ms per fused HMC launch
    at config 2 for every batch from 64 to 2 048 particles (profiles/engine_crossover.py) -- so the
    row-tile engine is the faster one at any size; the default keeps small batches on the warp-level
    engine because the reference-generated golden fixtures (64 particles, tests/test_gpu_golden.py)
    are held to 4x the CPU-fp32 error per teacher-forced step and one chaotic log p row of the
    row-tile engine sits at 4.4x."""
    return os.environ.get("FAB_ENGINE", "auto")


def rowtile_min_n() -> int:
    return int(os.environ.get("FAB_ROWTILE_MIN_N", "1024"))

_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension is the only implementation of the "
                "hot path (no CPU fallback). Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` from the repo root.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc < 0:
        msg = lib().fab_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libfab_b200 {what} failed ({rc}): {msg}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("fab_torch_b200: hot-path tensors must live on a CUDA device "
                           "(there is no CPU implementation)")
    if not t.is_contiguous():
        raise RuntimeError("fab_torch_b200: tensor must be contiguous")
    return t.data_ptr()


def f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise RuntimeError(f"fab_torch_b200 kernels are fp32; got {t.dtype} "
                           "(set torch default dtype to float32)")
    return t


def stream_ptr(device=None) -> int:
    """Current stream of `device`.  The kernels launch on the CURRENT CUDA device, so a tensor on
    another device than the current one is refused here instead of failing inside the launch."""
    if device is not None:
        dev = torch.device(device)
        if dev.type == "cuda" and dev.index is not None and dev.index != torch.cuda.current_device():
            raise RuntimeError(
                f"fab_torch_b200: tensors live on cuda:{dev.index} but the current device is "
                f"cuda:{torch.cuda.current_device()}; call torch.cuda.set_device({dev.index}) (one process "
                "per GPU) or wrap the call in `with torch.cuda.device(...)`")
    return torch.cuda.current_stream(device).cuda_stream


def point_ptrs(pt) -> PointPtrs:
    if pt is None:
        return PointPtrs(None, None, None, None, None)
    return PointPtrs(ptr(f32(pt.x)), ptr(f32(pt.log_q)), ptr(f32(pt.log_p)),
                     ptr(pt.grad_log_q) if pt.grad_log_q is not None else None,
                     ptr(pt.grad_log_p) if pt.grad_log_p is not None else None)
