"""Row-tile (tcgen05 / TMEM / TMA) engine vs. the oracle: RealNVP log-density + input-gradient and
the fused HMC step at the BASELINE config-2 architecture (dim 32, 10 layers x 320), through the C
ABI (fab_flow_logprob_grad_umma_f32 / fab_hmc_step_umma_f32).  Same bars as the warp-level engine's
tests (helpers.assert_parity): 1e-5 relative to the fp64 truth, or no worse than 4x the error of the
reference's own fp32 CPU arithmetic on the same inputs."""
import copy

import pytest
import torch

import fab_torch_b200 as fb
from fab_torch_b200 import _lib
from helpers import make_flows, make_manywell, rel_err, assert_parity
from oracle.noise import Float32RecordingNoise, ReplayNoise
from oracle.sampler import OracleHMC, make_point, beta_schedule
from test_gpu_transitions import _to_cuda_point, _cast, _flipped

pytestmark = pytest.mark.gpu


@pytest.fixture
def rowtile(monkeypatch):
    monkeypatch.setenv("FAB_ENGINE", "rowtile")


@pytest.fixture
def warp(monkeypatch):
    monkeypatch.setenv("FAB_ENGINE", "warp")


FLOW_CASES = [
    # K, nodes_per_dim (W = 32 * npd), n
    (1, 2, 64),
    (1, 10, 1),
    (2, 10, 200),        # ragged: one full pair + 72 rows
    (3, 6, 129),
    (10, 10, 2048),      # BASELINE config 2
    (10, 4, 300),
]


@pytest.mark.parametrize("K,npd,n", FLOW_CASES)
def test_rowtile_flow_logprob_and_grad(rowtile, K, npd, n):
    fo64, fo, fp = make_flows(32, K, npd, last_std=0.02)
    assert fp.rowtile_supported()
    torch.manual_seed(5)
    x = torch.randn(n, 32)
    xd = x.double().requires_grad_(True)
    lq64 = fo64.log_prob(xd)
    g64 = torch.autograd.grad(lq64.sum(), xd)[0]
    xf = x.clone().requires_grad_(True)
    lq32 = fo.log_prob(xf)
    g32 = torch.autograd.grad(lq32.sum(), xf)[0]
    lq, g = fp.cuda_log_prob(x.cuda(), with_grad=True)
    lq0, g0 = fp.cuda_log_prob(x.cuda(), with_grad=False)
    assert g0 is None and torch.equal(lq, lq0)
    e1 = assert_parity(lq, lq64.detach(), lq32.detach(), "log_q", floor=1e-5)
    e2 = assert_parity(g, g64, g32, "grad_log_q", floor=1e-4, outlier_frac=0.005, outlier_cap=None)
    print(f"\nrow-tile flow K={K} W={32 * npd} n={n}: (cuda err, cpu-fp32 err) log_q {e1} grad {e2}")


def test_rowtile_matches_warp_engine_closely(monkeypatch):
    """The two engines evaluate the same flow: results agree to fp32 rounding level."""
    _, _, fp = make_flows(32, 10, 10, last_std=0.02)
    torch.manual_seed(6)
    x = torch.randn(777, 32, device="cuda")
    monkeypatch.setenv("FAB_ENGINE", "warp")
    lq_w, g_w = fp.cuda_log_prob(x, with_grad=True)
    monkeypatch.setenv("FAB_ENGINE", "rowtile")
    lq_r, g_r = fp.cuda_log_prob(x, with_grad=True)
    assert rel_err(lq_r, lq_w) < 3e-6
    rows = ((g_r - g_w).abs() / g_w.abs().clamp_min(1.0)).max(dim=1).values
    assert (rows > 1e-4).sum() <= 8, "more than 1 % of the rows differ (ReLU kinks excepted)"


def test_rowtile_results_do_not_depend_on_the_batch(rowtile):
    """A particle's result is independent of the tile / CTA pair it is evaluated in (needed for
    sharded == single-device bit-equality, tests/multi_gpu_check.py)."""
    _, _, fp = make_flows(32, 4, 10, last_std=0.02)
    torch.manual_seed(7)
    x = torch.randn(500, 32, device="cuda")
    lq_all, g_all = fp.cuda_log_prob(x, with_grad=True)
    lq_sub, g_sub = fp.cuda_log_prob(x[137:333].contiguous(), with_grad=True)
    assert torch.equal(lq_all[137:333], lq_sub) and torch.equal(g_all[137:333], g_sub)


def test_rowtile_more_pairs_than_sm_pairs(rowtile):
    """16 384 particles (config 3's batch on one GPU) = 128 CTA pairs on 74 SM pairs: the second wave of
    clusters must produce the same bits as the same rows evaluated alone -- flow value + gradient and a
    fused HMC transition."""
    _, _, fp = make_flows(32, 3, 10, last_std=0.02)
    g = torch.Generator().manual_seed(21)
    n = 16384
    x = (torch.randn(n, 32, generator=g) * 1.3).cuda()
    lq_all, g_all = fp.cuda_log_prob(x, with_grad=True)
    for lo, hi in ((0, 256), (9472, 9472 + 384), (n - 200, n)):          # first wave, across the wave boundary, tail
        lq_sub, g_sub = fp.cuda_log_prob(x[lo:hi].contiguous(), with_grad=True)
        assert torch.equal(lq_all[lo:hi], lq_sub) and torch.equal(g_all[lo:hi], g_sub), (lo, hi)
    assert torch.isfinite(lq_all).all()
    _, tp = make_manywell(32)
    M = 3

    def run(xs, mom, ex):
        op = fb.HamiltonianMonteCarlo(M, 32, fp.log_prob, tp.log_prob, alpha=2.0, epsilon=0.1, L=2).cuda()
        op.set_eval_mode(True)
        pt = op.create_new_point(xs.clone())        # (the transition mutates the point in place)
        pt = fb.Point(*(t.detach().contiguous() for t in (pt.x, pt.log_q, pt.log_p, pt.grad_log_q, pt.grad_log_p)))
        op.run(pt, 2, 0.5, noise=(mom, ex))
        return pt
    mom = torch.randn(1, n, 32, generator=g).cuda()
    ex = torch.empty(1, n).exponential_(1.0, generator=g).cuda()
    big = run(x, mom, ex)
    lo, hi = 9472 - 128, 9472 + 256
    sub = run(x[lo:hi].contiguous(), mom[:, lo:hi].contiguous(), ex[:, lo:hi].contiguous())
    assert torch.equal(big.x[lo:hi], sub.x) and torch.equal(big.log_q[lo:hi], sub.log_q)
    assert torch.equal(big.grad_log_q[lo:hi], sub.grad_log_q) and torch.equal(big.log_p[lo:hi], sub.log_p)


def test_rowtile_repack_after_parameter_update(rowtile):
    fo64, fo, fp = make_flows(32, 2, 10, last_std=0.02)
    x = torch.randn(128, 32)
    lq_a, _ = fp.cuda_log_prob(x.cuda(), with_grad=False)
    with torch.no_grad():
        for p in fp.parameters():
            p.mul_(1.01)
        for p in fo64.parameters():
            p.mul_(1.01)
    lq_b, _ = fp.cuda_log_prob(x.cuda(), with_grad=False)
    assert not torch.equal(lq_a, lq_b)
    assert rel_err(lq_b, fo64.log_prob(x.double()).detach()) < 1e-5


def test_rowtile_unsupported_shapes_fall_back_or_raise(monkeypatch):
    _, _, fp = make_flows(8, 2, 4)
    assert not fp.rowtile_supported()
    monkeypatch.setenv("FAB_ENGINE", "auto")
    assert not fp.use_rowtile(4096)
    monkeypatch.setenv("FAB_ENGINE", "rowtile")
    with pytest.raises(RuntimeError):
        fp.use_rowtile(4096)


HMC_CASES = [
    # K, npd, M, i, L, n_outer, eps, p_target, alpha, B
    (10, 10, 16, 1, 5, 1, 0.1, False, 2.0, 2048),      # BASELINE config 2, first distribution
    (10, 10, 16, 16, 5, 1, 0.1, False, 2.0, 300),
    (3, 4, 4, 2, 3, 3, 0.2, True, None, 200),          # n_outer > 1: proposal carried between outer steps
    (2, 10, 4, 4, 2, 1, 0.3, False, 0.5, 64),
]


@pytest.mark.parametrize("K,npd,M,i,L,n_outer,eps,p_target,alpha,B", HMC_CASES)
@pytest.mark.parametrize("tune", [True, False])
def test_rowtile_hmc_transition(rowtile, K, npd, M, i, L, n_outer, eps, p_target, alpha, B, tune):
    dim = 32
    fo64, fo, fp = make_flows(dim, K, npd, last_std=0.02)
    to, tp = make_manywell(dim)
    beta = beta_schedule("linear", M)[i]
    torch.manual_seed(11)
    x = fo.sample((B,)).detach()
    kw = dict(alpha=alpha, p_target=p_target, epsilon=eps, n_outer=n_outer, L=L, eval_mode=not tune)
    pt0 = make_point(x, fo.log_prob, to.log_prob, with_grad=True)
    op_o = OracleHMC(M, dim, fo64.log_prob, to.log_prob, **kw).double()
    op_o.noise = Float32RecordingNoise()
    out_o = op_o.transition(_cast(pt0, torch.float64), i, beta)
    op_32 = OracleHMC(M, dim, fo.log_prob, to.log_prob, **kw)
    op_32.noise = ReplayNoise(copy.deepcopy(op_o.noise.record))
    out_32 = op_32.transition(_cast(pt0, torch.float32), i, beta)
    op_p = fb.HamiltonianMonteCarlo(M, dim, fp.log_prob, tp.log_prob, **kw).cuda()
    op_p.noise = fb.InjectedNoise(op_o.noise.record)
    pt_p = _to_cuda_point(pt0)
    out_p = op_p.transition(pt_p, i, beta)
    torch.cuda.synchronize()
    fl = _flipped(out_p.x, out_o.x)
    fl32 = _flipped(out_32.x, out_o.x)
    assert fl.sum() <= max(1, 0.01 * B), f"{int(fl.sum())} accept flips (cpu fp32: {int(fl32.sum())})"
    ok = ~(fl | fl32)
    report = {}
    for name, floor in (("x", 1e-5), ("log_q", 1e-5), ("log_p", 1e-5), ("grad_log_q", 1e-4),
                        ("grad_log_p", 1e-4)):
        report[name] = assert_parity(getattr(out_p, name), getattr(out_o, name), getattr(out_32, name),
                                     name, floor=floor, mask=ok, outlier_frac=0.01)
    print(f"\nrow-tile HMC K={K} i={i} B={B}: (cuda err, cpu-fp32 err) {report}; flips {int(fl.sum())}")
    assert rel_err(op_p.epsilons, op_o.epsilons) < 1e-6
    assert rel_err(op_p.common_epsilon, op_o.common_epsilon) < 1e-6
    if i == 1:
        for n in range(n_outer):
            assert abs(op_p.first_dist_p_accepts[n].item() - op_o.first_dist_p_accepts[n].item()) < 2e-4
        a, b = op_p.average_distance_first_dist.item(), op_o.average_distance_first_dist.item()
        assert abs(a - b) <= 2e-4 * max(1.0, abs(b))


def test_rowtile_hmc_with_device_count_and_ragged_tail(rowtile):
    """n_active on the device (after a NaN filter) below the launch size: rows beyond it untouched."""
    dim, M = 32, 4
    _, _, fp = make_flows(dim, 2, 10, last_std=0.02)
    _, tp = make_manywell(dim)
    op = fb.HamiltonianMonteCarlo(M, dim, fp.log_prob, tp.log_prob, alpha=2.0, epsilon=0.2, L=2).cuda()
    torch.manual_seed(3)
    pt = op.create_new_point(torch.randn(300, dim, device="cuda"))
    before = copy.deepcopy(pt)
    n_active = torch.tensor([170], dtype=torch.int32, device="cuda")
    op.run(pt, 1, 0.25, n_active=n_active)
    torch.cuda.synchronize()
    assert torch.equal(pt.x[170:], before.x[170:]) and torch.equal(pt.log_q[170:], before.log_q[170:])
    assert not torch.equal(pt.x[:170], before.x[:170])
    assert float(op._stats[1]) == 170.0
