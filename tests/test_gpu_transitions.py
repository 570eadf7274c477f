"""Fused HMC / Metropolis transitions vs. the fp64 oracle, teacher-forced: both start from the same
Point and consume the same noise, so differences are rounding only (SURVEY §7 'chain-level parity
is fragile').  A particle whose accept test lands within rounding of the threshold may flip; at
most one such flip per case is tolerated and excluded from the value comparison."""
import copy

import pytest
import torch

import fab_torch_b200 as fb
from helpers import make_flows, make_manywell, make_gmm, rel_err
from oracle.noise import Float32RecordingNoise
from oracle.sampler import OracleHMC, OracleMetropolis, Point as OPoint, make_point, beta_schedule

pytestmark = pytest.mark.gpu


def _to_cuda_point(pt: OPoint) -> fb.Point:
    c = lambda t: None if t is None else t.detach().float().cuda().contiguous()
    return fb.Point(c(pt.x), c(pt.log_q), c(pt.log_p), c(pt.grad_log_q), c(pt.grad_log_p))


def _quantize(pt: OPoint) -> OPoint:
    """Round the oracle's starting Point to fp32 so both sides start from identical numbers."""
    q = lambda t: None if t is None else t.detach().float().double()
    return OPoint(q(pt.x), q(pt.log_q), q(pt.log_p), q(pt.grad_log_q), q(pt.grad_log_p))


HMC_CASES = [
    # dim, K, npd, target, M, i, L, n_outer, eps, p_target, alpha, B
    (32, 10, 10, "mw", 4, 1, 5, 1, 0.08, False, 2.0, 200),
    (32, 10, 10, "mw", 4, 4, 5, 1, 0.08, True, None, 33),
    (6, 3, 8, "mw", 3, 2, 3, 3, 0.15, False, 0.5, 500),
    (2, 0, 1, "gmm", 10, 5, 5, 2, 1.0, False, 2.0, 256),
    (2, 4, 40, "gmm", 8, 8, 2, 1, 0.5, True, None, 64),
]


@pytest.mark.parametrize("dim,K,npd,tk,M,i,L,n_outer,eps,p_target,alpha,B", HMC_CASES)
@pytest.mark.parametrize("tune", [True, False])
def test_hmc_transition(dim, K, npd, tk, M, i, L, n_outer, eps, p_target, alpha, B, tune):
    fo64, fo, fp = make_flows(dim, K, npd, last_std=0.05)
    if tk == "mw":
        to, tp = make_manywell(dim)
    else:
        to, _, tp = make_gmm(dim, 4, 8.0)
    beta = beta_schedule("linear", M)[i]
    torch.manual_seed(11)
    x = fo.sample((B,)).detach()
    op_o = OracleHMC(M, dim, fo64.log_prob, to.log_prob, alpha=alpha, p_target=p_target,
                     epsilon=eps, n_outer=n_outer, L=L, eval_mode=not tune).double()
    op_o.noise = Float32RecordingNoise()
    pt_o = _quantize(make_point(x.double(), fo64.log_prob, to.log_prob, with_grad=True))
    pt_p = _to_cuda_point(pt_o)
    x_before = pt_o.x.clone()
    out_o = op_o.transition(pt_o, i, beta)

    op_p = fb.HamiltonianMonteCarlo(M, dim, fp.log_prob, tp.log_prob, alpha=alpha,
                                    p_target=p_target, epsilon=eps, n_outer=n_outer, L=L,
                                    eval_mode=not tune).cuda()
    op_p.noise = fb.InjectedNoise(op_o.noise.record)
    out_p = op_p.transition(pt_p, i, beta)
    assert out_p is pt_p
    torch.cuda.synchronize()

    moved_o = (out_o.x != x_before).any(dim=1)
    agree = torch.ones(B, dtype=torch.bool)
    # a flipped accept shows up as an O(1) difference in x
    dx = (out_p.x.cpu().double() - out_o.x).abs().max(dim=1).values
    flipped = dx > 1e-2 * (1 + out_o.x.abs().max(dim=1).values)
    assert flipped.sum() <= 1, f"{int(flipped.sum())} accept flips"
    ok = ~flipped
    assert moved_o.any() and (~moved_o).any() or B < 50, "case should exercise accept and reject"
    for name, tol in (("x", 2e-5), ("log_q", 2e-5), ("log_p", 2e-5), ("grad_log_q", 2e-4),
                      ("grad_log_p", 2e-4)):
        err = rel_err(getattr(out_p, name).cpu()[ok], getattr(out_o, name)[ok])
        assert err < tol, f"{name}: rel err {err:.3e}"
    # tuner state and logging scalars
    assert rel_err(op_p.epsilons, op_o.epsilons) < 1e-6
    assert rel_err(op_p.common_epsilon, op_o.common_epsilon) < 1e-6
    if i in (1, M):
        info_p, info_o = op_p.get_logging_info(), op_o.get_logging_info()
        assert set(info_p) == set(info_o)
        for k in info_o:
            assert abs(info_p[k] - info_o[k]) <= 2e-4 * max(1.0, abs(info_o[k])), (k, info_p[k], info_o[k])


MET_CASES = [
    # dim, K, npd, target, M, i, n_updates, step, B
    (2, 4, 40, "gmm", 8, 3, 1, 5.0, 512),
    (2, 0, 1, "gmm", 10, 10, 5, 1.0, 300),
    (32, 2, 4, "mw", 4, 2, 3, 0.05, 100),
]


@pytest.mark.parametrize("dim,K,npd,tk,M,i,n_updates,step,B", MET_CASES)
@pytest.mark.parametrize("tune", [True, False])
def test_metropolis_transition(dim, K, npd, tk, M, i, n_updates, step, B, tune):
    fo64, fo, fp = make_flows(dim, K, npd)
    if tk == "mw":
        to, tp = make_manywell(dim)
    else:
        to, _, tp = make_gmm(dim, 4, 8.0)
    beta = beta_schedule("linear", M)[i]
    torch.manual_seed(12)
    x = fo.sample((B,)).detach()
    kw = dict(n_updates=n_updates, alpha=2.0, p_target=False, max_step_size=step,
              min_step_size=step * 0.2, adjust_step_size=tune)
    # fp32 oracle on purpose: `exp(gamma' - gamma)` overflows to inf (-> 0 -> reject,
    # metropolis.py:63-64) at 88.7 in fp32 but not in fp64, so fp64 is a different algorithm here.
    op_o = OracleMetropolis(M, dim, fo.log_prob, to.log_prob, **kw)
    op_o.noise = Float32RecordingNoise()
    pt_o = make_point(x, fo.log_prob, to.log_prob, with_grad=False)
    pt_p = _to_cuda_point(pt_o)
    out_o = op_o.transition(pt_o, i, beta)
    op_p = fb.Metropolis(M, dim, fp.log_prob, tp.log_prob, **kw).cuda()
    op_p.noise = fb.InjectedNoise(op_o.noise.record)
    out_p = op_p.transition(pt_p, i, beta)
    torch.cuda.synchronize()
    dx = (out_p.x.cpu().double() - out_o.x).abs().max(dim=1).values
    flipped = dx > 1e-3 * (1 + out_o.x.abs().max(dim=1).values)
    assert flipped.sum() <= 1
    ok = ~flipped
    for name in ("x", "log_q", "log_p"):
        err = rel_err(getattr(out_p, name).cpu()[ok], getattr(out_o, name)[ok])
        assert err < 2e-5, f"{name}: rel err {err:.3e}"
    assert rel_err(op_p.noise_scalings, op_o.noise_scalings) < 1e-6
    assert op_p.get_logging_info().keys() == op_o.get_logging_info().keys()
