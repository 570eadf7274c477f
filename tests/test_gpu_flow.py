"""CUDA flow kernels (through the C ABI) vs. the fp64 oracle.  Tolerance: 1e-5 relative
(BASELINE.json north_star: "within 1e-5 relative fp32")."""
import pytest
import torch

from helpers import make_flows, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5

CASES = [(32, 10, 10), (2, 4, 40), (5, 3, 3), (6, 2, 10), (128, 2, 10), (8, 0, 1)]


@pytest.mark.parametrize("dim,K,npd", CASES)
@pytest.mark.parametrize("n", [1, 37, 512])
def test_log_prob_and_grad(dim, K, npd, n):
    fo64, _, fp = make_flows(dim, K, npd)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(n, dim, generator=g) * 1.5
    x64 = x.double().requires_grad_(True)
    lq_ref = fo64.log_prob(x64)
    g_ref = torch.autograd.grad(lq_ref.sum(), x64)[0]
    lq, grad = fp.cuda_log_prob(x.cuda(), with_grad=True)
    lq_only, none = fp.cuda_log_prob(x.cuda(), with_grad=False)
    assert none is None
    assert rel_err(lq, lq_ref) < TOL
    assert rel_err(lq_only, lq_ref) < TOL
    assert rel_err(grad, g_ref) < 5 * TOL


@pytest.mark.parametrize("dim,K,npd", CASES)
def test_sample(dim, K, npd):
    fo64, _, fp = make_flows(dim, K, npd)
    g = torch.Generator().manual_seed(6)
    eps = torch.randn(300, dim, generator=g)
    x_ref, lq_ref = fo64._nf_model.sample(300, eps=eps.double())
    x, lq = fp.cuda_sample(eps.cuda())
    assert rel_err(x, x_ref) < TOL
    assert rel_err(lq, lq_ref) < TOL
    # self-consistency: log_prob(sample) == returned log_q (oracle header: flow parity is unpinned)
    lq2, _ = fp.cuda_log_prob(x, with_grad=False)
    assert rel_err(lq2, lq) < 2 * TOL


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
    assert rel_err(gx, gx64) < 5 * TOL
    w = torch.softmax(torch.randn(50), 0)
    loss = -(w.cuda() * fp.log_prob(x.cuda())).mean()
    loss.backward()
    loss64 = -(w.double() * y64).mean()
    loss64.backward()
    for (n1, p1), (n2, p2) in zip(fp.named_parameters(), fo64.named_parameters()):
        assert n1 == n2
        assert p1.grad is not None, n1
        assert rel_err(p1.grad, p2.grad) < 1e-4, n1


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
