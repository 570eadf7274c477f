"""CPU, world_size 2, gloo: the particle-parallel host logic (fab_torch_b200/dist.py) -- sharding,
the all-gather of ESS partial quadruples and the all-reduce of tuner statistics -- gives every
rank the single-device answer."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fab_torch_b200 import dist as fdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _local_partial(lw: torch.Tensor) -> torch.Tensor:
    """What fab_ess_partial_f32 produces for one rank's shard (torch restatement for the CPU test)."""
    m = lw.max()
    e = torch.exp(lw - m)
    return torch.stack([m, e.sum(), (e * e).sum(), torch.tensor(float(lw.numel()))]).float()


def _worker(rank, world_size, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        group = dist.group.WORLD
        assert fdist.world(group) == (world_size, rank)
        B = 1000
        n_local = fdist.shard_size(B, group)
        assert n_local == B // world_size
        with pytest.raises(ValueError):
            fdist.shard_size(B + 1, group)
        g = torch.Generator().manual_seed(7)
        log_w = torch.randn(B, generator=g) * 4 + 100.0          # identical on every rank
        shard = log_w[rank * n_local:(rank + 1) * n_local]
        parts = fdist.gather_partials(_local_partial(shard), group)
        assert parts.shape == (4 * world_size,)
        ess, lse, cnt = fdist.merge_ess_partials(parts)
        w = torch.softmax(log_w.double(), 0)
        want_ess = 1 / (w ** 2).sum() / B
        want_lse = torch.logsumexp(log_w.double(), 0)
        assert abs(float(ess) - float(want_ess)) < 1e-6 * float(want_ess)
        assert abs(float(lse) - float(want_lse)) < 1e-5
        assert int(cnt) == B
        # tuner statistics: (sum of clamped acceptance, count, distance) -> identical decision
        stats = torch.tensor([0.3 * n_local + rank, float(n_local), 2.0 * (rank + 1), 0.0])
        fdist.reduce_stats(stats, group)
        assert float(stats[1]) == B
        assert abs(float(stats[0]) - (0.3 * B + sum(range(world_size)))) < 1e-4
        out[rank] = (float(ess), float(lse), float(stats[0]))
        # the NVLink peer-memory exchange of the tuner statistics is for NCCL groups of CUDA devices:
        # on a gloo / CPU group the operators must fall back to the all-reduce above
        assert fdist.PeerExchange.create(group, torch.device("cpu")) is None
        # global systematic resample over ragged shards (rank 0 lost 3 particles to the NaN
        # filter): every rank must receive exactly its slice of the single-device answer.  The
        # ancestor routine is injected: on a GPU it is the integer kernel, here the oracle.
        import numpy as np
        from oracle.resample import systematic_ancestors as oracle_anc
        from fab_torch_b200.point import Point
        from fab_torch_b200.resample import global_systematic_resample, resample_if_ess_below
        counts = [n_local - 3, n_local] if world_size == 2 else [n_local] * world_size
        offs = [sum(counts[:r]) for r in range(world_size + 1)]
        N = offs[-1]
        gg = torch.Generator().manual_seed(11)
        X = torch.randn(N, 5, generator=gg)
        LW = torch.randn(N, generator=gg) * 3
        LW[7] = float("-inf")                                    # zero-weight particle
        mine = slice(offs[rank], offs[rank + 1])
        pt = Point(X[mine].clone(), LW[mine].clone() * 2, LW[mine].clone() * 3, X[mine].clone() + 1, None)
        anc_fn = lambda lw, u0: torch.from_numpy(oracle_anc(lw.numpy(), u0))
        u0 = 123456789
        new_pt, anc, lw_new = global_systematic_resample(pt, LW[mine].clone(), u0, group, anc_fn)
        want = torch.from_numpy(oracle_anc(LW.numpy(), u0))[mine]
        assert torch.equal(anc, want)
        assert torch.equal(new_pt.x, X[want]) and torch.equal(new_pt.log_q, (LW * 2)[want])
        assert torch.equal(new_pt.grad_log_q, (X + 1)[want]) and new_pt.grad_log_p is None
        assert 7 not in set(want.tolist())
        mean_w = torch.logsumexp(LW.double(), 0) - np.log(N)
        assert lw_new.shape == (counts[rank],) and abs(float(lw_new[0]) - float(mean_w)) < 1e-5
        # trigger: above the threshold nothing happens, below every rank resamples
        same = resample_if_ess_below(pt, LW[mine], ess=0.9, threshold=0.5, u0=u0, group=group,
                                     ancestors_fn=anc_fn)
        assert same[2] is False and same[0] is pt
        trig = resample_if_ess_below(pt, LW[mine], ess=0.1, threshold=0.5, u0=u0, group=group,
                                     ancestors_fn=anc_fn)
        assert trig[2] is True and torch.equal(trig[0].x, X[want])
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world_size = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world_size, port, out), nprocs=world_size, join=True)
    assert len(out) == world_size
    assert out[0] == out[1]                                      # every rank holds the same scalars


def test_single_process_is_identity():
    assert fdist.world(None) == (1, 0)
    assert fdist.shard_size(2048, None) == 2048
    p = torch.tensor([1.0, 2.0, 3.0, 4.0])
    assert fdist.gather_partials(p, None) is p
    lw = torch.randn(257) * 3
    ess, lse, cnt = fdist.merge_ess_partials(_local_partial(lw))
    w = torch.softmax(lw.double(), 0)
    assert abs(float(ess) - float(1 / (w ** 2).sum() / 257)) < 1e-6
    assert abs(float(lse) - float(torch.logsumexp(lw.double(), 0))) < 1e-5
