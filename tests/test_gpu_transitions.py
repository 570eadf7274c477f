"""Fused HMC / Metropolis transitions vs. the oracle, teacher-forced: both sides start from the
same fp32 Point and consume the same noise, so differences are rounding only (SURVEY §7
'chain-level parity is fragile').  HMC is compared with the fp64 oracle under the assert_parity
bar (1e-5 relative, or no worse than 4x the reference's own fp32 CPU run); Metropolis is compared
with the fp32 oracle because its `exp(gamma' - gamma)` overflow -> reject rule (metropolis.py:63-64)
triggers at 88.7 in fp32 but not in fp64.  A particle whose accept test lands within rounding of
the threshold may flip; at most 1 % such flips are tolerated and excluded from value comparison.
Likewise at most 1 % of the particles may sit on a ReLU kink of the flow during the leapfrog (the
gradient is discontinuous there; helpers.assert_parity `outlier_frac`) and must stay within 100x
the bar."""
import copy

import pytest
import torch

import fab_torch_b200 as fb
from helpers import make_flows, make_manywell, make_gmm, rel_err, assert_parity
from oracle.noise import Float32RecordingNoise, ReplayNoise
from oracle.sampler import OracleHMC, OracleMetropolis, Point as OPoint, make_point, beta_schedule

pytestmark = pytest.mark.gpu


def _to_cuda_point(pt: OPoint) -> fb.Point:
    c = lambda t: None if t is None else t.detach().float().cuda().contiguous()
    return fb.Point(c(pt.x), c(pt.log_q), c(pt.log_p), c(pt.grad_log_q), c(pt.grad_log_p))


def _cast(pt: OPoint, dtype) -> OPoint:
    q = lambda t: None if t is None else t.detach().to(dtype).clone()
    return OPoint(q(pt.x), q(pt.log_q), q(pt.log_p), q(pt.grad_log_q), q(pt.grad_log_p))


def _flipped(x_a, x_b):
    dx = (x_a.cpu().double() - x_b.double()).abs().max(dim=1).values
    return dx > 1e-2 * (1 + x_b.double().abs().max(dim=1).values)


HMC_CASES = [
    # dim, K, npd, target, M, i, L, n_outer, eps, p_target, alpha, B
    (32, 10, 10, "mw", 4, 1, 5, 1, 0.25, False, 2.0, 200),
    (32, 10, 10, "mw", 4, 4, 5, 1, 0.25, True, None, 33),
    (6, 3, 8, "mw", 3, 2, 3, 3, 0.3, False, 0.5, 500),
    (2, 0, 1, "gmm", 10, 5, 5, 2, 1.0, False, 2.0, 256),
    (2, 4, 40, "gmm", 8, 8, 2, 1, 1.5, True, None, 64),
    (128, 2, 10, "mw", 2, 1, 2, 1, 0.1, False, 2.0, 40),
]


@pytest.mark.parametrize("dim,K,npd,tk,M,i,L,n_outer,eps,p_target,alpha,B", HMC_CASES)
@pytest.mark.parametrize("tune", [True, False])
def test_hmc_transition(dim, K, npd, tk, M, i, L, n_outer, eps, p_target, alpha, B, tune):
    fo64, fo, fp = make_flows(dim, K, npd, last_std=0.05 if dim < 100 else 0.01)
    if tk == "mw":
        to, tp = make_manywell(dim)
        to32 = to
    else:
        to, to32, tp = make_gmm(dim, 4, 8.0)
    beta = beta_schedule("linear", M)[i]
    torch.manual_seed(11)
    x = fo.sample((B,)).detach()
    kw = dict(alpha=alpha, p_target=p_target, epsilon=eps, n_outer=n_outer, L=L, eval_mode=not tune)
    # start point: computed once in fp32 (like the reference chain would hold it)
    pt0 = make_point(x, fo.log_prob, to32.log_prob, with_grad=True)
    # fp64 ground truth
    op_o = OracleHMC(M, dim, fo64.log_prob, to.log_prob, **kw).double()
    op_o.noise = Float32RecordingNoise()
    out_o = op_o.transition(_cast(pt0, torch.float64), i, beta)
    # the reference algorithm in fp32 on the CPU (yardstick)
    op_32 = OracleHMC(M, dim, fo.log_prob, to32.log_prob, **kw)
    op_32.noise = ReplayNoise(copy.deepcopy(op_o.noise.record))
    out_32 = op_32.transition(_cast(pt0, torch.float32), i, beta)
    # CUDA
    op_p = fb.HamiltonianMonteCarlo(M, dim, fp.log_prob, tp.log_prob, **kw).cuda()
    op_p.noise = fb.InjectedNoise(op_o.noise.record)
    pt_p = _to_cuda_point(pt0)
    out_p = op_p.transition(pt_p, i, beta)
    assert out_p is pt_p
    torch.cuda.synchronize()

    fl = _flipped(out_p.x, out_o.x)
    fl32 = _flipped(out_32.x, out_o.x)
    assert fl.sum() <= max(1, 0.01 * B), f"{int(fl.sum())} accept flips (cpu fp32: {int(fl32.sum())})"
    ok = ~(fl | fl32)
    report = {}
    for name, floor in (("x", 1e-5), ("log_q", 1e-5), ("log_p", 1e-5), ("grad_log_q", 1e-4),
                        ("grad_log_p", 1e-4)):
        report[name] = assert_parity(getattr(out_p, name), getattr(out_o, name),
                                     getattr(out_32, name), name, floor=floor, mask=ok,
                                     outlier_frac=0.01)
    print(f"\nHMC d={dim} i={i}: (cuda err, cpu-fp32 err) {report}; flips {int(fl.sum())}")
    # tuner state and logging scalars
    assert rel_err(op_p.epsilons, op_o.epsilons) < 1e-6
    assert rel_err(op_p.common_epsilon, op_o.common_epsilon) < 1e-6
    if i == 1:
        for n in range(n_outer):
            assert abs(op_p.first_dist_p_accepts[n].item() - op_o.first_dist_p_accepts[n].item()) < 2e-4
        a, b = op_p.average_distance_first_dist.item(), op_o.average_distance_first_dist.item()
        assert abs(a - b) <= 2e-4 * max(1.0, abs(b))
    elif i == M:
        for n in range(n_outer):
            assert abs(op_p.last_dist_p_accepts[n].item() - op_o.last_dist_p_accepts[n].item()) < 2e-4
        a, b = op_p.average_distance_last_dist.item(), op_o.average_distance_last_dist.item()
        assert abs(a - b) <= 2e-4 * max(1.0, abs(b))


def test_hmc_logging_keys():
    _, fo, fp = make_flows(4, 1, 4)
    to, tp = make_manywell(4)
    M = 3
    op_p = fb.HamiltonianMonteCarlo(M, 4, fp.log_prob, tp.log_prob, alpha=2.0, epsilon=0.1, L=2).cuda()
    op_o = OracleHMC(M, 4, fo.log_prob, to.log_prob, alpha=2.0, epsilon=0.1, L=2)
    x = torch.randn(16, 4)
    pt_o = make_point(x, fo.log_prob, to.log_prob, with_grad=True)
    pt_p = op_p.create_new_point(x.cuda())
    assert rel_err(pt_p.log_q, pt_o.log_q) < 1e-5 and rel_err(pt_p.grad_log_p, pt_o.grad_log_p) < 1e-5
    for i in range(1, M + 1):
        op_o.transition(pt_o, i, 0.3)
        op_p.transition(pt_p, i, 0.3)
    assert set(op_p.get_logging_info()) == set(op_o.get_logging_info())


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
        _, to, tp = make_gmm(dim, 4, 8.0)
    beta = beta_schedule("linear", M)[i]
    torch.manual_seed(12)
    x = fo.sample((B,)).detach()
    kw = dict(n_updates=n_updates, alpha=2.0, p_target=False, max_step_size=step,
              min_step_size=step * 0.2, adjust_step_size=tune)
    op_o = OracleMetropolis(M, dim, fo.log_prob, to.log_prob, **kw)
    op_o.noise = Float32RecordingNoise()
    pt_o = make_point(x, fo.log_prob, to.log_prob, with_grad=False)
    pt_p = _to_cuda_point(pt_o)
    out_o = op_o.transition(pt_o, i, beta)
    op_p = fb.Metropolis(M, dim, fp.log_prob, tp.log_prob, **kw).cuda()
    op_p.noise = fb.InjectedNoise(op_o.noise.record)
    out_p = op_p.transition(pt_p, i, beta)
    torch.cuda.synchronize()
    fl = _flipped(out_p.x, out_o.x)
    assert fl.sum() <= max(1, 0.01 * B)
    ok = ~fl
    for name in ("x", "log_q", "log_p"):
        err = rel_err(getattr(out_p, name).cpu()[ok], getattr(out_o, name)[ok])
        assert err < 2e-5, f"{name}: rel err {err:.3e}"
    assert rel_err(op_p.noise_scalings, op_o.noise_scalings) < 1e-6
    assert op_p.get_logging_info().keys() == op_o.get_logging_info().keys()
