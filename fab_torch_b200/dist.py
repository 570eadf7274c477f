"""Particle parallelism over GPUs: one process per GPU, particles sharded over ranks, weights and
tuner state replicated.  Within a transition every op is row-wise, so the only exchanges are the
batch-global scalars of SURVEY §8e:

  * ESS / log Z (ais.py:68-71,80-86): each rank reduces its live particles to the quadruple
    (max, sum e^(lw-max), sum e^(2(lw-max)), count); the quadruples are all-gathered (16 bytes per
    rank) and merged by `fab_ess_finalize_f32` -- one collective per ESS.
  * step-size tuners (hmc.py:122-123,162-170; metropolis.py:69-73): (sum of clamped acceptance,
    count, distance) triples are all-reduced (SUM), then every rank applies the same update to
    its replica of the tuner state -- one collective per HMC outer step / Metropolis transition,
    none when the tuner is off (`set_eval_mode`).

The same helpers run on CPU tensors with the gloo backend (tests/test_dist_gloo.py) and on CUDA
tensors with NCCL over NVLink.
"""
from typing import Optional, Tuple

import torch


def world(group) -> Tuple[int, int]:
    """(world_size, rank) of `group`; (1, 0) when running single-process."""
    if group is None:
        return 1, 0
    import torch.distributed as dist
    return dist.get_world_size(group), dist.get_rank(group)


def shard_size(batch_size: int, group) -> int:
    """Particles carried by this rank for a global batch (ranks carry equal, contiguous shards)."""
    w, _ = world(group)
    if batch_size % w != 0:
        raise ValueError(f"batch_size {batch_size} must be divisible by the world size {w}")
    return batch_size // w


def gather_partials(part: torch.Tensor, group) -> torch.Tensor:
    """All-gather one [k] tensor per rank into [world*k] (rank-major)."""
    w, _ = world(group)
    if w == 1:
        return part
    import torch.distributed as dist
    out = torch.empty(w * part.numel(), dtype=part.dtype, device=part.device)
    dist.all_gather_into_tensor(out, part.contiguous(), group=group)
    return out


def reduce_stats(stats: torch.Tensor, group) -> torch.Tensor:
    """In-place SUM all-reduce of the tuner statistics."""
    w, _ = world(group)
    if w > 1:
        import torch.distributed as dist
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def merge_ess_partials(parts: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Reference merge of [world*4] quadruples -> (ESS, logsumexp, count); the device path uses
    the kernel `fab_ess_finalize_f32`, this torch restatement documents the formula and is what
    the CPU/gloo tests check the wiring with."""
    q = parts.reshape(-1, 4).double()
    m, s1, s2, cnt = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    M = m.max()
    scale = torch.where(torch.isfinite(m), torch.exp(m - M), torch.zeros_like(m))
    S1 = (s1 * scale).sum()
    S2 = (s2 * scale * scale).sum()
    N = cnt.sum()
    return S1 * S1 / (S2 * N), M + torch.log(S1), N


class PeerExchange:
    """Symmetric exchange buffers for the tuner statistics (`fab_hmc_finish_peer_f32`): every rank's
    buffer is mapped into all peers (torch symmetric memory over NVLink), the kernel stores its triple
    into the peers and sums theirs -- no NCCL launch between two transitions.  `create` returns None
    when the group is not an NCCL group of CUDA devices on one node, when symmetric memory is not
    available, or with FAB_PEER_EXCHANGE=0; the decision is all-reduced so that every rank takes the
    same path."""

    def __init__(self, buf, handle, ptrs, seq, world_size, rank):
        self.buf, self.handle, self.ptrs, self.seq = buf, handle, ptrs, seq
        self.world, self.rank = world_size, rank

    @staticmethod
    def create(group, device) -> Optional["PeerExchange"]:
        import os
        import torch.distributed as dist
        w, r = world(group)
        if w == 1 or device.type != "cuda" or dist.get_backend(group) != "nccl":
            return None
        ok, made = 1, None
        if os.environ.get("FAB_PEER_EXCHANGE", "1") == "0" or w > 32:
            ok = 0
        else:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                from fab_torch_b200 import _lib
                nbytes = int(_lib.lib().fab_hmc_peer_buffer_bytes(w))
                _lib.check(nbytes, "fab_hmc_peer_buffer_bytes")
                buf = symm_mem.empty(nbytes // 4, dtype=torch.float32, device=device)
                buf.zero_()
                handle = symm_mem.rendezvous(buf, group.group_name)
                ptrs = torch.tensor([int(p) for p in handle.buffer_ptrs], dtype=torch.int64, device=device)
                seq = torch.zeros(1, dtype=torch.int32, device=device)
                made = PeerExchange(buf, handle, ptrs, seq, w, r)
            except Exception as e:           # noqa: BLE001 -- any failure means "use the all-reduce"
                ok = 0
                if r == 0:
                    print(f"[fab_torch_b200] peer exchange unavailable ({type(e).__name__}: {e}); using NCCL all-reduce")
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)             # every buffer is zeroed before anybody stores into it
        return made if int(flag.item()) == 1 else None
