"""GPU: the Many-Well evaluation path through the product classes --
`ManyWellEnergy.performance_metrics` with `B200RealNVP.log_prob` as the model density, against
oracle/eval_manywell.py (pinned to the reference) on the same test points, and
`generate_eval_data` -> `performance_metrics` the way `FABModel.get_eval_info`
(fab/core.py:191-220) chains them."""
import pytest
import torch

import fab_torch_b200 as fb
from golden_util import load_fixture
from helpers import make_flows
from oracle import eval_manywell as ev
from oracle.targets import OracleManyWell

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,batch_size", [(4, 10), (8, 64)])
def test_performance_metrics_match_oracle(dim, batch_size):
    e = load_fixture("eval_manywell")[dim]
    fo64, _, fp = make_flows(dim, 3, 5, seed=dim)
    tgt = fb.ManyWellEnergy(dim)
    orc = OracleManyWell(dim)
    fixed = e["samples"]                                  # the reference's own exact samples
    calls = {"n": 0}

    def fixed_sampler(shape):
        i = calls["n"] * shape[0]
        calls["n"] += 1
        return fixed[i:i + shape[0]]

    want = ev.performance_metrics(orc, e["log_w"].double(),
                                  lambda x: fo64.log_prob(x.double()).detach(), batch_size,
                                  sampler=fixed_sampler)
    calls["n"] = 0
    tgt.sample = lambda shape: fixed_sampler(shape).cuda()
    got = tgt.performance_metrics(None, e["log_w"], lambda x: fp.log_prob(x).detach(), batch_size)
    assert set(got) == set(want) and got["eval_batch_size"] == want["eval_batch_size"]
    # fp32 kernels vs the fp64 oracle: 1e-5 relative to the size of the log-densities averaged
    scale = 1.0 + abs(want["test_set_exact_mean_log_prob"]) + abs(want["test_set_modes_mean_log_prob"])
    for k in want:
        assert abs(got[k] - want[k]) <= 1e-5 * (scale + abs(want[k])), (k, got[k], want[k])
    # without a model density: pure log-weight arithmetic, the reference's numbers exactly
    assert tgt.performance_metrics(None, e["log_w"]) == e["metrics_no_q"]


def test_exact_sampler_on_device():
    tgt = fb.ManyWellEnergy(6)
    torch.manual_seed(3)
    x = tgt.sample((20000,))
    assert x.is_cuda and x.shape == (20000, 6) and bool(torch.isfinite(x).all())
    right = (x[:, 0::2] > 0).float().mean().item()
    assert abs(right - 0.8443) < 0.01                     # mass of the deep well (quadrature)
    assert abs(x[:, 1::2].mean().item()) < 0.02 and abs(x[:, 1::2].var().item() - 1.0) < 0.03
    # exact samples are typical under the target: E_p[log p] by quadrature = 3 wells x 8.3418
    lp = tgt.log_prob(x)
    assert bool(torch.isfinite(lp).all())
    modes = next(tgt.get_modes_test_set_iterator(8))
    assert modes.is_cuda and modes.shape[1] == 6


def test_eval_info_chain():
    """generate_eval_data -> performance_metrics as FABModel.get_eval_info does (core.py:191-220)."""
    dim, M = 8, 4
    _, _, flow = make_flows(dim, 3, 5, seed=1)
    tgt = fb.ManyWellEnergy(dim)
    op = fb.HamiltonianMonteCarlo(M, dim, flow.log_prob, tgt.log_prob, alpha=2.0, p_target=True,
                                  epsilon=0.1, L=3).cuda()
    ais = fb.AnnealedImportanceSampler(flow, tgt.log_prob, op, p_target=True, alpha=2.0,
                                       n_intermediate_distributions=M)
    base_x, base_lw, ais_x, ais_lw = ais.generate_eval_data(400, 200)
    assert base_x.device.type == "cpu" and base_x.shape[1] == dim and ais_lw.shape[0] <= 400
    info = {"eval_ess_flow": fb.effective_sample_size(base_lw.cuda()).item(),
            "eval_ess_ais": fb.effective_sample_size(ais_lw.cuda()).item()}
    info.update({"flow_" + k: v for k, v in
                 tgt.performance_metrics(base_x, base_lw, flow.log_prob, batch_size=200).items()})
    info.update({"ais_" + k: v for k, v in tgt.performance_metrics(ais_x, ais_lw).items()})
    for k in ("flow_relative_MSE_Z_estimate", "flow_test_set_modes_mean_log_prob", "flow_forward_kl",
              "flow_test_set_exact_mean_log_prob", "ais_abs_MSE_log_Z_estimate", "eval_ess_ais"):
        assert k in info and info[k] == info[k], k
    assert info["flow_eval_batch_size"] == 200 and 0.0 < info["eval_ess_ais"] <= 1.0


def test_gmm_model_density_branch_on_device():
    """gmm.py:71-99 with `log_q_fn`: the metrics that need the model density (test-set mean log q,
    forward KL, ESS over p) through the CUDA flow / target kernels vs float64 torch on the SAME test
    set (the reference draws a fresh test set on every access, so the set is pinned here)."""
    from oracle.targets import OracleGMM, to_double
    import copy
    dim, n_mixes = 2, 6
    torch.manual_seed(3)
    tgt = fb.GMM(dim, n_mixes, 6.0, 0.5, true_expectation_estimation_n_samples=int(1e4))
    torch.manual_seed(3)
    orc = to_double(copy.deepcopy(OracleGMM(dim, n_mixes, 6.0, 0.5)))
    assert torch.equal(orc.locs.float(), tgt.locs.cpu())
    fo64, _, fp = make_flows(dim, 3, 10, seed=5)
    g = torch.Generator().manual_seed(8)
    test_set = torch.randn(1000, dim, generator=g) * 3.0
    original = fb.GMM.__dict__["test_set"]
    fb.GMM.test_set = property(lambda self: test_set.cuda())              # pinned for this test
    try:
        x = torch.randn(400, dim, generator=g).cuda()
        log_w = torch.randn(400, generator=g).cuda()
        tgt._true_expectation = torch.tensor(1.0, device="cuda")
        got = tgt.performance_metrics(x, log_w, lambda t: fp.log_prob(t).detach())
    finally:
        fb.GMM.test_set = original
    lq = fo64.log_prob(test_set.double()).detach()
    lp = orc.log_prob(test_set.double())
    ratio = lp - lq
    want = dict(test_set_mean_log_prob=lq.mean().item(), kl_forward=ratio.mean().item(),
                ess_over_p=(1 / torch.exp(ratio).mean()).item())
    for k, v in want.items():
        assert abs(got[k] - v) <= 2e-5 * (1.0 + abs(v)), (k, got[k], v)
    assert set(got) == {"test_set_mean_log_prob", "bias_normed", "bias_no_correction", "ess_over_p", "kl_forward"}
