"""`effective_sample_size` (fab/utils/numerical.py:18-23) as a device reduction."""
import torch

from fab_torch_b200 import _lib


def effective_sample_size(log_w: torch.Tensor, normalised: bool = False) -> torch.Tensor:
    """1 / (N * sum softmax(log_w)^2); returns a 0-dim CUDA tensor (no host sync)."""
    assert len(log_w.shape) == 1
    if normalised:
        return 1 / torch.sum(log_w ** 2) / log_w.shape[0]
    lw = _lib.f32(log_w.detach()).contiguous()
    L = _lib.lib()
    part = torch.empty(4, dtype=torch.float32, device=lw.device)
    out = torch.empty(3, dtype=torch.float32, device=lw.device)
    s = _lib.stream_ptr(lw.device)
    _lib.check(L.fab_ess_partial_f32(_lib.ptr(lw), None, lw.shape[0], None, _lib.ptr(part), s),
               "fab_ess_partial_f32")
    _lib.check(L.fab_ess_finalize_f32(_lib.ptr(part), 1, _lib.ptr(out), s), "fab_ess_finalize_f32")
    return out[0]
