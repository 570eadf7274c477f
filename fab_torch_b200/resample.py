"""Systematic resampling of a Point by its AIS log-weights (build-side extension; the reference
only has multinomial `resample`, fab/sampling_methods/base.py:121-124).  The ancestor indices come
from an integer kernel with a fixed-point CDF and are bit-exact against oracle/resample.py."""
from typing import Callable, Optional, Tuple

import torch

from fab_torch_b200 import _lib
from fab_torch_b200 import dist as fdist
from fab_torch_b200.point import Point


def systematic_ancestors(log_w: torch.Tensor, u0: int) -> torch.Tensor:
    """int64[N] ancestors for offset u0 in [0, 2^32) (position k uses (k + u0/2^32)/N)."""
    lw = _lib.f32(log_w.detach()).contiguous()
    n = lw.shape[0]
    if not bool(torch.isfinite(lw).any()):
        # the AIS NaN filter tests log_q / log_p, not log_w (ais.py:190-213): a batch whose
        # log-weights are all non-finite has no resampling distribution
        raise ValueError("systematic resampling needs at least one finite log-weight")
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


# ------------------------------------------------------------------------------------ multi-rank
def _all_gather_rows(t: torch.Tensor, counts, n_max: int, group) -> torch.Tensor:
    """Concatenate the first counts[r] rows of every rank's `t` (rank-major).  Shards are padded
    to n_max rows for the equal-size all-gather and compacted with the known counts."""
    import torch.distributed as dist
    w = len(counts)
    row = t.shape[1:] if t.dim() > 1 else ()
    pad = torch.zeros((n_max,) + tuple(row), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    out = torch.empty((w * n_max,) + tuple(row), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    if all(c == n_max for c in counts):
        return out
    return torch.cat([out[r * n_max: r * n_max + counts[r]] for r in range(w)], dim=0)


def global_systematic_resample(point: Point, log_w: torch.Tensor, u0: int, group=None,
                               ancestors_fn: Optional[Callable] = None
                               ) -> Tuple[Point, torch.Tensor, torch.Tensor]:
    """Systematic resampling of a particle set that is SHARDED over the ranks of `group`
    (BASELINE config 3).  Every rank passes its live particles; the result is exactly the
    single-device answer on the rank-major concatenation: rank r receives the particles of the
    global positions [offset_r, offset_r + n_r).

    Exchange: one all-gather of the shard sizes (host), one of the log-weights (4 B/particle) and
    one of the Point rows ((3d+2)*4 B/particle -- 6.4 MB at 16384 x 32, microseconds over NVSwitch).
    Every rank then runs the same integer kernel on the same global weight vector, so the
    ancestors are bit-identical on all ranks and to oracle/resample.py by construction.

    `u0` must be the same on every rank.  `ancestors_fn(log_w_global, u0) -> int64[N]` defaults to
    the CUDA kernel (`systematic_ancestors`); the CPU/gloo wiring test injects the oracle.
    Returns (resampled local Point, local ancestors as GLOBAL indices, uniform local log_w =
    logsumexp(global log_w) - log N, the mean weight every survivor carries)."""
    w, rank = fdist.world(group)
    if ancestors_fn is None:
        ancestors_fn = systematic_ancestors
    n_local = int(log_w.shape[0])
    if w == 1:
        counts, n_max = [n_local], n_local
        lw_all = log_w.detach()
        gather = lambda t: t
    else:
        import torch.distributed as dist
        cnt = [None] * w
        dist.all_gather_object(cnt, n_local, group=group)
        counts, n_max = [int(c) for c in cnt], max(int(c) for c in cnt)
        gather = lambda t: _all_gather_rows(t.detach().contiguous(), counts, n_max, group)
        lw_all = gather(log_w)
    anc_all = ancestors_fn(lw_all, int(u0) & 0xffffffff)
    off = sum(counts[:rank])
    anc = anc_all[off: off + n_local].contiguous()
    pick = lambda t: None if t is None else gather(t).index_select(0, anc.to(t.device))
    out = Point(pick(point.x), pick(point.log_q), pick(point.log_p), pick(point.grad_log_q),
                pick(point.grad_log_p))
    n_tot = float(sum(counts))
    lw_new = (torch.logsumexp(lw_all.double(), 0) - torch.log(torch.tensor(n_tot, dtype=torch.float64))
              ).to(log_w.dtype).expand(n_local).contiguous().to(log_w.device)
    return out, anc, lw_new


def resample_if_ess_below(point: Point, log_w: torch.Tensor, ess: float, threshold: float,
                          u0: Optional[int] = None, group=None, ancestors_fn=None):
    """The global ESS / resample trigger of BASELINE's north-star: `ess` is the GLOBAL effective
    sample size fraction (e.g. `ais.get_logging_info()["ess_ais"]`, already all-reduced over the
    ranks); when it is below `threshold` the particle set is resampled systematically across all
    ranks and the weights are reset to the mean weight.  Returns (point, log_w, resampled?).
    Every rank takes the same branch because `ess` and `u0` are rank-invariant."""
    if not (ess < threshold):
        return point, log_w, False
    if u0 is None:
        g = torch.tensor([int(torch.randint(0, 2 ** 32, (1,), dtype=torch.int64).item())], dtype=torch.int64)
        w, _ = fdist.world(group)
        if w > 1:
            import torch.distributed as dist
            g = g.to(log_w.device)
            dist.broadcast(g, src=dist.get_global_rank(group, 0) if hasattr(dist, "get_global_rank") else 0,
                           group=group)
        u0 = int(g.item())
    pt, _, lw = global_systematic_resample(point, log_w, u0, group, ancestors_fn)
    return pt, lw, True
