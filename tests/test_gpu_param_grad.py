"""Parameter gradient of the FAB loss through the CUDA kernels: forward tape
(fab_flow_logprob_tape_f32) + batch-contraction weight-gradient GEMMs (fab_flow_param_grad_f32) vs
float64 autograd of the oracle flow, at the BASELINE config-2 and config-5 architectures.

loss = -mean(softmax(log_w) * log q(x))   (fab/core.py:112-118).  Bar: every parameter's gradient
within 1e-4 of the fp64 gradient relative to that tensor's largest entry, or no worse than 4x the
error of fp32 autograd (the reference's own arithmetic) on the same inputs."""
import pytest
import torch

from helpers import make_flows

pytestmark = pytest.mark.gpu


def _tensor_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


CASES = [
    ("config 2 (32, 10 x 320)", 32, 10, 10, 2048, 0.02),
    ("config 5 (60, 10 x 300)", 60, 10, 5, 1024, 0.01),
    ("odd sizes (5, 3 x 15)", 5, 3, 3, 77, 0.05),
    ("no coupling layers", 8, 0, 1, 33, 0.05),
]


@pytest.mark.parametrize("name,dim,K,npd,n,last_std", CASES)
def test_param_grad_matches_fp64_autograd(name, dim, K, npd, n, last_std, monkeypatch):
    monkeypatch.setenv("FAB_ENGINE", "warp")
    fo64, fo, fp = make_flows(dim, K, npd, last_std=last_std)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(n, dim, generator=g) * 1.2
    log_w = torch.randn(n, generator=g) * 2.0
    w = torch.softmax(log_w, 0)

    def loss_of(flow, xx, ww):
        return -(ww * flow.log_prob(xx)).mean()

    loss64 = loss_of(fo64, x.double(), w.double())
    loss64.backward()
    loss32 = loss_of(fo, x, w)
    loss32.backward()
    fp.zero_grad()
    loss = loss_of(fp, x.cuda(), w.cuda())
    loss.backward()
    assert abs(loss.item() - loss64.item()) < 1e-5 * max(1.0, abs(loss64.item()))
    worst, worst32, who = 0.0, 0.0, ""
    for (n1, p1), (n2, p2), (n3, p3) in zip(fp.named_parameters(), fo64.named_parameters(), fo.named_parameters()):
        assert n1 == n2 == n3
        assert p1.grad is not None and p1.grad.shape == p2.grad.shape, n1
        e, e32 = _tensor_err(p1.grad, p2.grad), _tensor_err(p3.grad, p2.grad)
        if e > worst:
            worst, who = e, n1
        worst32 = max(worst32, e32)
        assert e <= max(1e-4, 4 * e32), f"{name}: {n1}: cuda {e:.3e}, fp32 autograd {e32:.3e}"
    print(f"\n{name} n={n}: worst parameter-gradient error vs fp64: cuda {worst:.3e} ({who}); fp32 autograd {worst32:.3e}")


def test_param_grad_with_input_grad_and_repeat():
    """x.requires_grad and parameters together; a second backward through a fresh forward gives the
    same bits (fixed batch slices, no atomics)."""
    fo64, fo, fp = make_flows(6, 3, 8)
    x = torch.randn(300, 6)
    outs = []
    for _ in range(2):
        fp.zero_grad()
        xg = x.cuda().requires_grad_(True)
        y = fp.log_prob(xg)
        (y * torch.linspace(0.5, 1.5, 300).cuda()).sum().backward()
        outs.append([p.grad.clone() for p in fp.parameters()] + [xg.grad.clone()])
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    x64 = x.double().requires_grad_(True)
    (fo64.log_prob(x64) * torch.linspace(0.5, 1.5, 300).double()).sum().backward()
    assert _tensor_err(outs[0][-1], x64.grad) < 5e-5
    for p1, p2 in zip(fp.parameters(), fo64.parameters()):
        assert _tensor_err(p1.grad, p2.grad) < 1e-4


def test_param_grad_follows_parameter_updates_through_the_captured_chain_rule():
    """The parameter-space chain rule runs eagerly on the first call, is captured into a CUDA graph on
    the second and replayed afterwards (flow.py: cuda_param_grad / _repack).  With an optimiser step
    between the calls every one of them must give the gradient at the CURRENT parameters, and the
    gradients handed out earlier must not change when the persistent buffer is overwritten."""
    fo64, fo, fp = make_flows(6, 3, 8)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(200, 6, generator=g)
    w = torch.softmax(torch.randn(200, generator=g), 0)
    kept = []
    for it in range(5):
        fp.zero_grad()
        fo64.zero_grad()
        (-(w.cuda() * fp.log_prob(x.cuda())).mean()).backward()
        (-(w.double() * fo64.log_prob(x.double())).mean()).backward()
        for (n1, p1), (n2, p2) in zip(fp.named_parameters(), fo64.named_parameters()):
            assert _tensor_err(p1.grad, p2.grad) < 1e-4, f"call {it}: {n1}"
        kept.append(([p.grad for p in fp.parameters()], [p.grad.clone() for p in fp.parameters()]))
        with torch.no_grad():
            for p1, p2 in zip(fp.parameters(), fo64.parameters()):
                step = 0.05 * torch.randn(p2.shape, generator=g, dtype=torch.float64)
                p2.add_(step)
                p1.add_(step.float().cuda())
    assert isinstance(fp._pack_graphs.get(f"pg{len(x)}", {}).get("graph"), torch.cuda.CUDAGraph)
    for held, copies in kept:
        for a, b in zip(held, copies):
            assert torch.equal(a, b)
