"""CPU: host-side logic of the plugin surface (packing, descriptors, coefficient maths, API
contracts, loud failure without a GPU)."""
import numpy as np
import pytest
import torch

import fab_torch_b200 as fb
from blob_interp import interp_log_prob_and_grad, interp_sample
from helpers import make_flows
from oracle.sampler import OracleHMC, OracleMetropolis, Point as OPoint, beta_schedule, gamma, grad_gamma


@pytest.mark.parametrize("dim,K,npd,act_norm", [(32, 3, 10, False), (2, 4, 40, False), (5, 2, 3, False), (6, 2, 7, False),
                                                 (9, 1, 5, False), (6, 3, 7, True), (5, 2, 3, True)])
def test_packed_blob_matches_oracle(dim, K, npd, act_norm):
    fo64, fo, fp = make_flows(dim, K, npd, device=None, act_norm=act_norm)
    x = torch.randn(11, dim, dtype=torch.float64).requires_grad_(True)
    lq = fo64.log_prob(x)
    g = torch.autograd.grad(lq.sum(), x)[0]
    blob, d = fp.blob().numpy(), fp.desc()
    lq2, g2 = interp_log_prob_and_grad(blob, d, x.detach().numpy())
    assert np.abs(lq.detach().numpy() - lq2).max() < 2e-5
    assert np.abs(g.numpy() - g2).max() < 2e-5
    eps = torch.randn(11, dim, dtype=torch.float64)
    xs, lqs = fo64._nf_model.sample(11, eps=eps)
    xs2, lqs2 = interp_sample(blob, d, eps.numpy())
    assert np.abs(xs.detach().numpy() - xs2).max() < 2e-5
    assert np.abs(lqs.detach().numpy() - lqs2).max() < 2e-5


def test_same_seed_same_weights_and_state_dict_keys():
    from oracle.realnvp import OracleRealNVP
    torch.manual_seed(3)
    a = OracleRealNVP(6, 2, 5)
    torch.manual_seed(3)
    b = fb.B200RealNVP(6, 2, 5)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa) == list(sb)
    assert all(torch.equal(sa[k], sb[k]) for k in sa)
    assert "_nf_model.flows.0.flows.1.param_map.net.4.weight" in sb and "_nf_model.flows.1.log_S" in sb
    assert b.event_shape == (6,)


def test_act_norm_same_seed_same_init_and_state_dict_keys():
    """make_normflow_model.py:28-29,94-95: ActNorm after every InvertibleAffine, initialised from the
    statistics of 500 samples at construction.  Same seed -> the same s, t as the oracle restatement
    (same RNG consumption, same arithmetic), normflows-style keys, and the init leaves every layer's
    output standardised for that batch."""
    from oracle.realnvp import OracleRealNVP
    torch.manual_seed(5)
    a = OracleRealNVP(6, 3, 4, act_norm=True)
    torch.manual_seed(5)
    b = fb.B200RealNVP(6, 3, 4, act_norm=True)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa) == list(sb)
    for k in sa:
        assert torch.allclose(sa[k], sb[k], rtol=0, atol=1e-6), k
    assert "_nf_model.flows.2.s" in sb and "_nf_model.flows.8.t" in sb
    assert float(sb["_nf_model.flows.5.data_dep_init_done"]) == 1.0
    assert sb["_nf_model.flows.2.s"].abs().max() > 1e-3          # the init did something
    assert not fb.B200RealNVP(6, 0, 4, act_norm=True).act_norm    # no layers: nothing to normalise
    # a second construction consumes the RNG the same way (randn(500, d) once)
    torch.manual_seed(5)
    fb.B200RealNVP(6, 3, 4, act_norm=True)
    r1 = torch.rand(1)
    torch.manual_seed(5)
    OracleRealNVP(6, 3, 4, act_norm=True)
    assert torch.equal(r1, torch.rand(1))


@pytest.mark.parametrize("act_norm", [False, True])
def test_torch_restatement_in_product_matches_oracle(act_norm):
    fo64, fo, fp = make_flows(6, 3, 4, device=None, act_norm=act_norm)
    x = torch.randn(20, 6)
    assert torch.allclose(fp.torch_log_prob(x), fo.log_prob(x), atol=2e-5)
    eps = torch.randn(20, 6)
    xs, lq = fp.torch_sample(eps)
    xo, lqo = fo._nf_model.sample(20, eps=eps)
    assert torch.allclose(xs, xo, atol=2e-5) and torch.allclose(lq, lqo, atol=2e-5)


@pytest.mark.parametrize("p_target,alpha", [(True, None), (False, 2.0), (False, 0.5)])
def test_gamma_coefficients_match_reference_formulas(p_target, alpha):
    pt = OPoint(torch.zeros(3, 2), torch.tensor([1.5, -2.0, 0.3]), torch.tensor([-0.7, 4.0, 2.2]),
                torch.randn(3, 2), torch.randn(3, 2))
    for beta in beta_schedule("geometric", 7):
        g = fb.make_gamma(beta, alpha, p_target)
        want = gamma(pt, beta, alpha, p_target)
        got = np.float32(g.cq) * pt.log_q + np.float32(g.cp) * pt.log_p
        assert torch.equal(got, want)
        wantg = grad_gamma(pt, beta, alpha, p_target)
        gotg = np.float32(g.gq) * pt.grad_log_q + np.float32(g.gp) * pt.grad_log_p
        assert torch.equal(gotg, wantg)


def test_beta_grid_and_operator_state_match_oracle():
    for kind in ("linear", "geometric"):
        for M in (1, 4, 16, 40):
            assert torch.equal(fb.setup_distribution_spacing(kind, M), beta_schedule(kind, M))
    with pytest.raises(Exception, match="distribution spacing incorrectly specified"):
        fb.setup_distribution_spacing("cosine", 3)
    _, fo, fp = make_flows(4, 1, 3, device=None)
    tp = fb.ManyWellEnergy(4, use_gpu=False)
    hp = fb.HamiltonianMonteCarlo(5, 4, fp.log_prob, tp.log_prob, alpha=2.0, epsilon=0.7, n_outer=2)
    ho = OracleHMC(5, 4, fo.log_prob, None, alpha=2.0, epsilon=0.7, n_outer=2)
    assert list(hp.state_dict()) == list(ho.state_dict()) == ["common_epsilon", "epsilons", "mass_vector"]
    assert all(torch.equal(hp.state_dict()[k], ho.state_dict()[k]) for k in ho.state_dict())
    mp = fb.Metropolis(5, 4, fp.log_prob, tp.log_prob, n_updates=3, max_step_size=2.0)
    mo = OracleMetropolis(5, 4, fo.log_prob, None, n_updates=3, max_step_size=2.0)
    assert list(mp.state_dict()) == ["noise_scalings"]
    assert torch.equal(mp.noise_scalings, mo.noise_scalings)
    assert hp.uses_grad_info and not mp.uses_grad_info
    mp.set_eval_mode(True)
    assert mp.eval_mode is False            # inverted on purpose (reference quirk 4)
    hp.set_eval_mode(True)
    assert hp.eval_mode is True


def test_no_cpu_fallback():
    """The product must fail loudly instead of computing on the CPU."""
    _, _, fp = make_flows(4, 1, 3, device=None)
    with pytest.raises(RuntimeError, match="CUDA"):
        fp.log_prob(torch.randn(3, 4))
    with pytest.raises(RuntimeError, match="CUDA"):
        fp.sample_and_log_prob((3,))
    tp = fb.ManyWellEnergy(4, use_gpu=False)
    with pytest.raises(RuntimeError, match="CUDA"):
        tp.log_prob(torch.randn(3, 4))
    with pytest.raises(TypeError):
        fb.HamiltonianMonteCarlo(2, 4, lambda x: x.sum(1), tp.log_prob, alpha=2.0)
    with pytest.raises(TypeError):
        fb.HamiltonianMonteCarlo(2, 4, fp.log_prob, torch.distributions.Normal(0., 1.).log_prob,
                                 alpha=2.0)


def test_point_mask_semantics():
    pt = fb.Point(torch.arange(6.).reshape(3, 2), torch.arange(3.), torch.arange(3.) + 10,
                  torch.zeros(3, 2), torch.ones(3, 2))
    m = torch.tensor([True, False, True])
    sub = pt[m]
    assert sub.x.shape == (2, 2) and torch.equal(sub.log_p, torch.tensor([10., 12.]))
    other = fb.Point(torch.full((3, 2), -1.), torch.full((3,), -1.), torch.full((3,), -1.),
                     torch.full((3, 2), -1.), torch.full((3, 2), -1.))
    pt[m] = other[m]
    assert torch.equal(pt.log_q, torch.tensor([-1., 1., -1.]))
    assert torch.equal(pt.grad_log_p[1], torch.ones(2)) and torch.equal(pt.grad_log_p[0], -torch.ones(2))


def test_state_dict_keys_follow_normflows_fixture():
    """checkpoints written by FABModel.save (core.py:222-260) must keep the key names
    normflows gives the reference's flow; the fixture lists them explicitly so that a rename fails here."""
    import json
    import os
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "normflows_state_dict_keys.json")))
    import fab_torch_b200 as fb
    flow = fb.B200RealNVP(4, fx["n_flow_layers"], 3)
    assert sorted(flow.state_dict().keys()) == fx["keys"]
    flow = fb.B200RealNVP(4, fx["n_flow_layers"], 3, act_norm=True)
    assert sorted(flow.state_dict().keys()) == fx["keys_act_norm"]
    assert "_nf_model.flows.2.s" in fx["keys_act_norm"] and "_nf_model.flows.2.data_dep_init_done" in fx["keys_act_norm"]
