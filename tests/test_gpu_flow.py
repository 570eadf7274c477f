"""CUDA flow kernels (through the C ABI) vs. the fp64 oracle.  Bar: 1e-5 relative, or no worse
than 4x the reference's own fp32 CPU arithmetic where rounding is amplified (helpers.assert_parity)."""
import pytest
import torch

from helpers import make_flows, rel_err, assert_parity

pytestmark = pytest.mark.gpu

CASES = [(32, 10, 10), (2, 4, 40), (5, 3, 3), (6, 2, 10), (128, 2, 10), (8, 0, 1)]


@pytest.mark.parametrize("dim,K,npd", CASES)
@pytest.mark.parametrize("n", [1, 37, 512, 2048])
def test_log_prob_and_grad(dim, K, npd, n):
    fo64, fo, fp = make_flows(dim, K, npd, last_std=0.05 if dim < 100 else 0.01)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(n, dim, generator=g) * 1.5
    x64 = x.double().requires_grad_(True)
    lq_ref = fo64.log_prob(x64)
    g_ref = torch.autograd.grad(lq_ref.sum(), x64)[0]
    x32 = x.clone().requires_grad_(True)
    lq_32 = fo.log_prob(x32)
    g_32 = torch.autograd.grad(lq_32.sum(), x32)[0]
    lq, grad = fp.cuda_log_prob(x.cuda(), with_grad=True)
    lq_only, none = fp.cuda_log_prob(x.cuda(), with_grad=False)
    assert none is None
    assert_parity(lq, lq_ref, lq_32, "log_q")
    assert_parity(lq_only, lq_ref, lq_32, "log_q (value only)")
    assert_parity(grad, g_ref, g_32, "grad_log_q", floor=5e-5, outlier_frac=0.005)   # ReLU kinks


@pytest.mark.parametrize("dim,K,npd", CASES)
def test_sample(dim, K, npd):
    fo64, fo, fp = make_flows(dim, K, npd, last_std=0.05 if dim < 100 else 0.01)
    g = torch.Generator().manual_seed(6)
    eps = torch.randn(300, dim, generator=g)
    x_ref, lq_ref = fo64._nf_model.sample(300, eps=eps.double())
    x_32, lq_32 = fo._nf_model.sample(300, eps=eps)
    x, lq = fp.cuda_sample(eps.cuda())
    assert_parity(x, x_ref, x_32, "x")
    assert_parity(lq, lq_ref, lq_32, "log_q")
    # self-consistency: log_prob(sample) == returned log_q (oracle header: flow parity is unpinned)
    lq2, _ = fp.cuda_log_prob(x, with_grad=False)
    assert rel_err(lq2, lq) < 5e-5


def test_autograd_surface():
    """`grad_and_value` of the reference (base.py:50-56) must work on B200RealNVP.log_prob, and the
    parameter gradient used by the FAB loss (core.py:112-118) must match the oracle."""
    fo64, fo, fp = make_flows(6, 3, 8)
    x = torch.randn(50, 6)
    xg = x.cuda().requires_grad_(True)
    y = fp.log_prob(xg)
    gx = torch.autograd.grad(y, xg, grad_outputs=torch.ones_like(y), retain_graph=True)[0]
    x64 = x.double().requires_grad_(True)
    y64 = fo64.log_prob(x64)
    gx64 = torch.autograd.grad(y64.sum(), x64, retain_graph=True)[0]
    assert rel_err(gx, gx64) < 5e-5
    w = torch.softmax(torch.randn(50), 0)
    loss = -(w.cuda() * fp.log_prob(x.cuda())).mean()
    loss.backward()
    loss64 = -(w.double() * y64).mean()
    loss64.backward()
    for (n1, p1), (n2, p2) in zip(fp.named_parameters(), fo64.named_parameters()):
        assert n1 == n2
        assert p1.grad is not None, n1
        assert rel_err(p1.grad, p2.grad) < 1e-4, n1
    # reparameterised sampling path (flow_reverse_kl, core.py:130-133)
    fp.zero_grad()
    xs, lqs = fp.sample_and_log_prob((64,))
    (lqs.mean() + xs.pow(2).mean()).backward()
    assert all(p.grad is not None for p in fp.parameters())


def test_state_dict_roundtrip_and_repack():
    fo64, fo, fp = make_flows(4, 2, 5)
    x = torch.randn(9, 4).cuda()
    a, _ = fp.cuda_log_prob(x, False)
    with torch.no_grad():
        for p in fp.parameters():
            p.mul_(1.01)
    b, _ = fp.cuda_log_prob(x, False)
    assert (a - b).abs().max() > 0          # blob was repacked after the in-place update
    fp.load_state_dict(fo.state_dict())
    c, _ = fp.cuda_log_prob(x, False)
    assert torch.equal(a, c)


def test_results_do_not_depend_on_the_batch_they_are_evaluated_in(monkeypatch):
    """(Warp-level engine; the row-tile engine's twin is tests/test_gpu_rowtile.py.  FAB_ENGINE=auto
    switches engines with the batch size, and the two engines round differently, so the property
    holds per engine.)
    A particle's log q / gradient / sample is bit-identical whether it is evaluated alone, in a
    batch that maps to the 8-slot tile layout or in one that maps to the 16-slot layout (different
    particles per CTA): MMA rows are independent and every per-particle reduction runs in a
    canonical order.  This is what makes rank-sharded runs equal single-device runs bit for bit."""
    def same(a, b):                      # bitwise equality that treats NaN == NaN
        return torch.equal(torch.nan_to_num(a, nan=12345.0), torch.nan_to_num(b, nan=12345.0))

    monkeypatch.setenv("FAB_ENGINE", "warp")
    _, _, fp = make_flows(32, 10, 10, last_std=0.05)
    g = torch.Generator().manual_seed(11)
    x = (torch.randn(2048, 32, generator=g) * 1.5).cuda()
    lq_big, g_big = fp.cuda_log_prob(x, with_grad=True)            # 14 particles per CTA, 16 slots
    for n in (1, 5, 300, 1024):                                     # 1..7 per CTA, 8 slots
        lq, gr = fp.cuda_log_prob(x[:n].contiguous(), with_grad=True)
        assert same(lq, lq_big[:n]) and same(gr, g_big[:n]), n
    eps = torch.randn(2048, 32, generator=g).cuda()
    xs_big, lqs_big = fp.cuda_sample(eps)
    xs, lqs = fp.cuda_sample(eps[:640].contiguous())
    assert same(xs, xs_big[:640]) and same(lqs, lqs_big[:640])
