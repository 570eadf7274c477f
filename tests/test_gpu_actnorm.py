"""ActNorm (normflows `ActNorm(dim)`, make_normflow_model.py:28-29; `act_norm=True` is the default of the
reference factory `make_wrapped_normflow_realnvp`, :82-96) through the CUDA kernels: the host folds the
layer into the packed linear part of each block (include/fab_b200.h), so every kernel entry point is
checked against the fp64 oracle, which applies the layer as its own module.  Bars as in
test_gpu_flow.py / test_gpu_param_grad.py."""
import pytest
import torch

import fab_torch_b200 as fb
from helpers import make_flows, rel_err, assert_parity

pytestmark = pytest.mark.gpu

CASES = [(32, 10, 10), (2, 4, 40), (5, 3, 3), (60, 4, 5)]


@pytest.mark.parametrize("dim,K,npd,engine", [c + ("warp",) for c in CASES] + [(32, 10, 10, "rowtile"), (32, 3, 4, "rowtile")])
@pytest.mark.parametrize("n", [1, 300, 2048])
def test_log_prob_grad_and_sample_with_act_norm(dim, K, npd, engine, n, monkeypatch):
    monkeypatch.setenv("FAB_ENGINE", engine)
    fo64, fo, fp = make_flows(dim, K, npd, act_norm=True, last_std=0.02 if K >= 10 else 0.05)
    assert fp.act_norm and fp.use_rowtile(n) == (engine == "rowtile")
    g = torch.Generator().manual_seed(5)
    x = torch.randn(n, dim, generator=g) * 1.5
    x64 = x.double().requires_grad_(True)
    lq_ref = fo64.log_prob(x64)
    g_ref = torch.autograd.grad(lq_ref.sum(), x64)[0]
    x32 = x.clone().requires_grad_(True)
    lq_32 = fo.log_prob(x32)
    g_32 = torch.autograd.grad(lq_32.sum(), x32)[0]
    lq, grad = fp.cuda_log_prob(x.cuda(), with_grad=True)
    assert_parity(lq, lq_ref, lq_32, "log_q")
    # ReLU kinks: a hidden unit within rounding of zero takes the other branch (helpers.assert_parity)
    assert_parity(grad, g_ref, g_32, "grad_log_q", floor=5e-5, outlier_frac=0.005, outlier_cap=None)
    eps = torch.randn(n, dim, generator=g)
    x_ref, lqs_ref = fo64._nf_model.sample(n, eps=eps.double())
    x_32, lqs_32 = fo._nf_model.sample(n, eps=eps)
    xs, lqs = fp.cuda_sample(eps.cuda())
    assert_parity(xs, x_ref, x_32, "x")
    assert_parity(lqs, lqs_ref, lqs_32, "log_q of the sample")
    lq2, _ = fp.cuda_log_prob(xs, with_grad=False)
    assert rel_err(lq2, lqs) < 5e-5


@pytest.mark.parametrize("dim,K,npd,n", [(32, 10, 10, 2048), (6, 3, 8, 77)])
def test_param_grad_with_act_norm(dim, K, npd, n):
    """FAB loss gradient (core.py:112-118) for every parameter incl. the ActNorm s, t, from the tape
    kernels and the parameter-space chain rule, vs fp64 autograd of the oracle.  The small case is
    repeated with parameter updates in between (eager -> captured -> replayed chain rule); the
    config-2 size is evaluated once: with 13 M hidden units per pass, every new parameter set puts a few
    pre-activations within rounding of zero, and one flipped ReLU of one particle moves a weight
    gradient by ~1/n (3e-4 here; fp32 autograd does the same on other units)."""
    fo64, fo, fp = make_flows(dim, K, npd, act_norm=True, last_std=0.02)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(n, dim, generator=g) * 1.2
    w = torch.softmax(torch.randn(n, generator=g) * 2.0, 0)
    for it in range(3 if dim < 10 else 1):
        for f in (fp, fo64, fo):
            f.zero_grad()
        (-(w.cuda() * fp.log_prob(x.cuda())).mean()).backward()
        (-(w.double() * fo64.log_prob(x.double())).mean()).backward()
        (-(w * fo.log_prob(x)).mean()).backward()
        names = [nm for nm, _ in fp.named_parameters()]
        assert any(nm.endswith(".s") for nm in names) and any(nm.endswith(".t") for nm in names)
        for (n1, p1), (n2, p2), (n3, p3) in zip(fp.named_parameters(), fo64.named_parameters(), fo.named_parameters()):
            assert n1 == n2 == n3 and p1.grad is not None and p1.grad.shape == p2.grad.shape, n1
            e, e32 = rel_tensor(p1.grad, p2.grad), rel_tensor(p3.grad, p2.grad)
            assert e <= max(1e-4, 4 * e32), f"call {it}: {n1}: cuda {e:.3e}, fp32 autograd {e32:.3e}"
        with torch.no_grad():
            for p1, p2, p3 in zip(fp.parameters(), fo64.parameters(), fo.parameters()):
                step = 0.003 * torch.randn(p2.shape, generator=g, dtype=torch.float64)
                p2.add_(step); p3.add_(step.float()); p1.add_(step.float().cuda())


def rel_tensor(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def test_act_norm_init_on_device_and_sampling_surface():
    """Constructed on the CPU (init from 500 samples like the reference factory), moved to the GPU:
    sample_and_log_prob / log_prob agree with each other and the reparameterised sampling path
    differentiates w.r.t. s and t."""
    torch.manual_seed(2)
    fp = fb.make_wrapped_b200_realnvp(6, 3, 8, act_norm=True).cuda()
    xs, lq = fp.sample_and_log_prob((400,))
    assert rel_err(fp.log_prob(xs), lq) < 5e-5
    (lq.mean() + xs.pow(2).mean()).backward()
    acts = fp._acts()
    assert len(acts) == 3 and all(a.s.grad is not None and a.t.grad is not None for a in acts)
    # the data-dependent init standardised each block's output for its init batch: s, t moved off zero
    assert all(float(a.data_dep_init_done) == 1.0 and a.s.abs().max() > 1e-3 for a in acts)
