"""Device prioritised replay buffer (through the C ABI) vs the reference fixture and the oracle:
ring writes and adjust are bit-exact, the Gumbel-top-k index SET is exact (integer work), also at
the sizes the training configs use."""
import pytest
import torch

import fab_torch_b200 as fb
from fab_torch_b200 import _lib
from oracle.buffer import OracleBuffer, gumbel_like, topk_set
from test_oracle_buffer import replay

pytestmark = pytest.mark.gpu


def _never():
    raise AssertionError("initial_sampler must not be called")


def test_cuda_buffer_matches_reference_fixture():
    def make(c):
        return fb.PrioritisedReplayBuffer(c["dim"], c["max_length"], c["min_sample_length"], _never,
                                          device="cuda", fill_buffer_during_init=False)
    replay(make,
           lambda b, x, lw, lq: b.add(x, lw, lq),
           lambda b, k, z: b._topk(b.buffer.log_w[:(b.max_length if b.is_full else b.current_index)].contiguous(),
                                   z.cuda().contiguous(), k),
           lambda b, adj, lq, idx: b.adjust(adj, lq, idx),
           lambda b: dict(x=b.buffer.x, log_w=b.buffer.log_w, log_q_old=b.buffer.log_q_old,
                          current_index=b.current_index, is_full=b.is_full, can_sample=b.can_sample))


@pytest.mark.parametrize("n,k", [(1, 1), (37, 37), (1000, 1), (5000, 999), (200000, 4096), (1 << 20, 10240)])
def test_topk_index_set_is_exact(n, k):
    g = torch.Generator().manual_seed(n + k)
    logits = torch.randn(n, generator=g) * 5 + 20
    if n > 100:
        logits[7] = -float("inf")                 # killed sample
        logits[11:15] = logits[10]                # exact ties in the logits
    z = gumbel_like(logits) if n < (1 << 20) else -torch.log(-torch.log(torch.rand(n, generator=g).clamp_min(1e-30)))
    want = topk_set(logits, z, k)
    L = _lib.lib()
    ws = torch.empty(int(L.fab_buffer_topk_workspace_bytes(n)), dtype=torch.uint8, device="cuda")
    idx = torch.empty(k, dtype=torch.int64, device="cuda")
    lg, zg = logits.cuda(), z.cuda()
    _lib.check(L.fab_buffer_topk_f32(_lib.ptr(lg), _lib.ptr(zg), n, k, _lib.ptr(idx), _lib.ptr(ws),
                                     _lib.stream_ptr()))
    got = idx.cpu()
    assert bool((got[1:] > got[:-1]).all()) if k > 1 else True          # ascending, unique
    if not torch.equal(got, want):
        # only exact ties of the perturbed value at the threshold may differ
        v = (z + logits)
        thr = torch.sort(v, descending=True).values[k - 1]
        diff = set(got.tolist()) ^ set(want.tolist())
        assert all(v[i] == thr for i in diff), f"{len(diff)} indices differ beyond threshold ties"


def test_sample_interface_and_properties():
    dim, N = 8, 4096
    g = torch.Generator().manual_seed(3)
    data = (torch.randn(N, dim, generator=g), torch.randn(N, generator=g) * 2, torch.randn(N, generator=g))
    buf = fb.PrioritisedReplayBuffer(dim, N, 1024, lambda: tuple(t[:2048] for t in data), device="cuda")
    assert buf.can_sample and not buf.is_full and buf.current_index == 2048
    buf.add(*(t[2048:] for t in data))
    assert buf.is_full and buf.current_index == 0
    torch.manual_seed(5)
    batches = buf.sample_n_batches(128, 4)
    assert len(batches) == 4 and batches[0][0].shape == (128, dim)
    idx = torch.cat([b[3] for b in batches])
    assert idx.unique().numel() == 512                                   # without replacement
    x = torch.cat([b[0] for b in batches])
    assert torch.equal(x, buf.buffer.x[idx]) and torch.equal(batches[1][1], buf.buffer.log_w[batches[1][3]])
    # the same seed gives the reference's index set (oracle restatement on the same RNG calls)
    torch.manual_seed(5)
    z = gumbel_like(data[1])
    assert torch.equal(torch.sort(idx.cpu()).values, topk_set(data[1], z, 512))
    # high-weight samples are preferred
    assert buf.buffer.log_w[idx].mean() > buf.buffer.log_w.mean() + 1.0
    # save / load round trip
    import tempfile, os
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "buf.pt")
        buf.save(p)
        other = fb.PrioritisedReplayBuffer(dim, N, 1024, _never, device="cuda", fill_buffer_during_init=False)
        other.load(p)
        assert torch.equal(other.buffer.x, buf.buffer.x) and other.is_full and other.can_sample
    with pytest.raises(RuntimeError):
        fb.PrioritisedReplayBuffer(dim, N, 10, _never, device="cpu", fill_buffer_during_init=False)


def test_config5_chain_feeds_the_buffer():
    """BASELINE config 5 path: ALDP-surrogate AIS (20 distributions, HMC) -> buffer.add ->
    sample -> adjust with the weight correction of train_with_prioritised_buffer.py:158-186."""
    dim, M, B = 60, 20, 256
    flow = fb.B200RealNVP(dim, 4, 5).cuda()
    target = fb.AldpSurrogateEnergy(dim)
    op = fb.HamiltonianMonteCarlo(M, dim, flow.log_prob, target.log_prob, alpha=2.0, p_target=False,
                                  epsilon=0.05, L=4).cuda()
    ais = fb.AnnealedImportanceSampler(flow, target.log_prob, op, p_target=False, alpha=2.0,
                                       n_intermediate_distributions=M)

    def sampler():
        pt, lw = ais.sample_and_log_weights(B, logging=False)
        return pt.x, lw, pt.log_q
    buf = fb.PrioritisedReplayBuffer(dim, 2048, 512, sampler, device="cuda")
    assert buf.can_sample and buf.current_index == 512
    x, lw, lq_old, idx = buf.sample(128)
    lq_new = flow.log_prob(x).detach()
    adj = (1 - 2.0) * (lq_new - lq_old)                     # (1 - alpha)(log q_new - log q_old)
    before = buf.buffer.log_w[idx].clone()
    buf.adjust(adj, lq_new, idx)
    assert torch.allclose(buf.buffer.log_w[idx], before + adj) and torch.equal(buf.buffer.log_q_old[idx], lq_new)
    assert torch.isfinite(buf.buffer.log_w[:512]).all()


# ---- fab/utils/replay_buffer.py (un-prioritised buffer) ------------------------------------------
def test_cuda_replay_buffer_matches_reference_fixture():
    """Device ReplayBuffer vs the fixture written by the unmodified reference class: ring state bit-exact,
    sampled rows identical and in the reference's order when the exponential variates are the recorded ones."""
    from test_oracle_buffer import replay_uniform, _init_add

    def make(c, T):
        return fb.ReplayBuffer(c["dim"], c["max_length"], c["min_sample_length"], _never, device="cuda",
                               temperature=T, fill_buffer_during_init=False)

    def sample(b, k, q):
        max_index = b.max_length if b.is_full else b.current_index
        rank = b.current_add_count - b.buffer.add_count[:max_index]
        idx = b._race(torch.pow(1 / rank, b.temperature), q.cuda(), k)
        return b.buffer.x[idx], b.buffer.log_w[idx], idx
    replay_uniform(make, _init_add, sample,
                   lambda b: dict(x=b.buffer.x, log_w=b.buffer.log_w, add_count=b.buffer.add_count,
                                  current_index=b.current_index, current_add_count=b.current_add_count,
                                  is_full=b.is_full, can_sample=b.can_sample))


def test_replay_buffer_interface_and_seeded_sampling():
    """Constructor fill loop, sample_n_batches, and: the same seed gives the rows the reference algorithm
    (oracle restatement on the same RNG call) samples; newer batches are preferred at temperature 1."""
    from oracle.buffer import OracleReplayBuffer
    dim, N, B = 6, 3000, 500
    g = torch.Generator().manual_seed(9)
    batches = [(torch.randn(B, dim, generator=g), torch.randn(B, generator=g)) for _ in range(9)]
    it, it2 = iter(batches), iter(batches)
    buf = fb.ReplayBuffer(dim, N, 900, lambda: next(it), device="cuda")
    orc = OracleReplayBuffer(dim, N, 900)
    orc.fill(lambda: next(it2))
    assert buf.can_sample and buf.current_index == 1000 and buf.current_add_count == 1
    for b in it:
        buf.add(*b)
    for b in it2:
        orc.add(*b)
    assert buf.is_full and buf.current_index == orc.current_index == 1500
    assert torch.equal(buf.buffer.add_count.cpu(), orc.add_count)
    torch.manual_seed(4)
    data = buf.sample_n_batches(100, 4)
    torch.manual_seed(4)
    x_o, lw_o, idx_o = orc.sample(400)
    assert len(data) == 4 and data[0][0].shape == (100, dim)
    assert torch.equal(torch.cat([d[0] for d in data]).cpu(), x_o) and torch.equal(torch.cat([d[1] for d in data]).cpu(), lw_o)
    assert idx_o.unique().numel() == 400
    # rank weighting: rows of the latest add (rank 1) are drawn far more often than the oldest (rank 6)
    newest = ((idx_o >= 1000) & (idx_o < 1500)).sum().item()
    oldest = ((idx_o >= 1500) & (idx_o < 2000)).sum().item()
    assert newest > 2 * oldest
    with pytest.raises(Exception):
        fb.ReplayBuffer(dim, N, 900, _never, device="cuda", fill_buffer_during_init=False).sample(10)
    with pytest.raises(RuntimeError):
        fb.ReplayBuffer(dim, N, 10, _never, device="cpu", fill_buffer_during_init=False)
