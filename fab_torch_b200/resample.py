"""Systematic resampling of a Point by its AIS log-weights (build-side extension; the reference
only has multinomial `resample`, fab/sampling_methods/base.py:121-124).  The ancestor indices come
from an integer kernel with a fixed-point CDF and are bit-exact against oracle/resample.py."""
from typing import Optional

import torch

from fab_torch_b200 import _lib
from fab_torch_b200.point import Point


def systematic_ancestors(log_w: torch.Tensor, u0: int) -> torch.Tensor:
    """int64[N] ancestors for offset u0 in [0, 2^32) (position k uses (k + u0/2^32)/N)."""
    lw = _lib.f32(log_w.detach()).contiguous()
    n = lw.shape[0]
    L = _lib.lib()
    ws = torch.empty(int(L.fab_resample_workspace_bytes(n)), dtype=torch.uint8, device=lw.device)
    anc = torch.empty(n, dtype=torch.int64, device=lw.device)
    rc = L.fab_resample_systematic_u64(_lib.ptr(lw), n, int(u0) & 0xffffffff, _lib.ptr(anc),
                                       _lib.ptr(ws), _lib.stream_ptr(lw.device))
    _lib.check(rc, "fab_resample_systematic_u64")
    return anc


def _gather(t: Optional[torch.Tensor], anc: torch.Tensor) -> Optional[torch.Tensor]:
    if t is None:
        return None
    src = _lib.f32(t).contiguous()
    rowf = 1 if src.dim() == 1 else src.shape[1]
    dst = torch.empty_like(src)
    rc = _lib.lib().fab_gather_rows_f32(_lib.ptr(src), _lib.ptr(dst), _lib.ptr(anc), src.shape[0],
                                        rowf, _lib.stream_ptr(src.device))
    _lib.check(rc, "fab_gather_rows_f32")
    return dst


def systematic_resample(point: Point, log_w: torch.Tensor, u0: Optional[int] = None):
    """Returns (resampled Point, ancestors).  u0 defaults to a draw from torch's CPU generator."""
    if u0 is None:
        u0 = int(torch.randint(0, 2 ** 32, (1,), dtype=torch.int64).item())
    anc = systematic_ancestors(log_w, u0)
    return Point(_gather(point.x, anc), _gather(point.log_q, anc), _gather(point.log_p, anc),
                 _gather(point.grad_log_q, anc), _gather(point.grad_log_p, anc)), anc
