"""CPU: the oracle against the committed golden vectors (written by oracle/gen_golden.py from the
UNMODIFIED reference) and against the constants the reference publishes."""
import copy
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle.noise import ReplayNoise
from oracle.realnvp import OracleRealNVP, randomize_last_layers, analytic_log_prob_and_grad
from oracle.sampler import OracleAIS, OracleHMC, OracleMetropolis, effective_sample_size
from oracle.targets import OracleGMM, OracleManyWell, to_double
from golden_util import GOLDEN_DIR, load_fixture, rebuild_flow, rebuild_target

FIXTURES = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.pt"))
                  if not os.path.basename(p).startswith(("prioritised_buffer", "replay_buffer", "eval_")))   # test_oracle_buffer.py


def test_pin_report_is_green():
    rep = json.load(open(os.path.join(GOLDEN_DIR, "pin_report.json")))
    assert len(rep["checks"]) >= 50 and all(c["ok"] for c in rep["checks"])


def test_published_constants():
    assert abs(float(OracleManyWell(32).log_Z) - 164.69567532) < 1e-7
    assert abs(float(OracleManyWell(128).log_Z) - 658.7827013) < 1e-6
    torch.manual_seed(0)
    g = OracleGMM(2, 40, 40, 1.0)
    want = torch.tensor([[-0.29947, 21.45774], [-32.92181, -29.43756], [-15.40617, 10.72629]])
    assert torch.allclose(g.locs[:3], want, atol=1e-4)
    assert abs(float(g.scale_trils[0, 0, 0]) - 1.3132616) < 1e-6


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_reproduces_golden_fp64(name):
    """Replaying the recorded noise through the fp64 oracle must give the stored ground truth."""
    fx = load_fixture(name)
    cfg = fx["config"]
    flow = rebuild_flow(fx).double()
    tgt = to_double(rebuild_target(fx))
    cls = OracleHMC if cfg["kind"] == "hmc" else OracleMetropolis
    op = cls(cfg["M"], cfg["dim"], flow.log_prob, tgt.log_prob, alpha=cfg["alpha"],
             p_target=cfg["p_target"], **cfg["opkw"])
    op.load_state_dict(fx["op_state_before"])
    op = op.double()
    op.noise = ReplayNoise(copy.deepcopy(fx["noise"]))
    ais = OracleAIS(flow, tgt.log_prob, op, p_target=cfg["p_target"], alpha=cfg["alpha"],
                    n_intermediate_distributions=cfg["M"], distribution_spacing_type=cfg["spacing"])
    flow._eps_override = fx["noise"]["base_eps"][0].double()
    pt, lw = ais.sample_and_log_weights(cfg["B"])
    want = fx["fp64"]
    assert torch.allclose(lw, want["log_w"], rtol=1e-7, atol=1e-7)
    assert torch.allclose(pt.x, want["point"]["x"], rtol=1e-7, atol=1e-7)
    lz = want["info"]["log_Z"]
    assert abs(ais.get_logging_info()["log_Z"] - lz) < 1e-9 * max(1.0, abs(lz)) + 1e-9


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_reproduces_reference_fp32(name):
    """fp32 replay of the recorded noise must reproduce the outputs the REFERENCE classes produced
    (bit-identical on the machine that generated them; a different host BLAS may reorder sums, so
    the check allows rounding-level differences and the rare accept flip they cause)."""
    fx = load_fixture(name)
    cfg = fx["config"]
    flow = rebuild_flow(fx)
    tgt = rebuild_target(fx)
    cls = OracleHMC if cfg["kind"] == "hmc" else OracleMetropolis
    op = cls(cfg["M"], cfg["dim"], flow.log_prob, tgt.log_prob, alpha=cfg["alpha"],
             p_target=cfg["p_target"], **cfg["opkw"])
    op.load_state_dict(fx["op_state_before"])
    op.noise = ReplayNoise(copy.deepcopy(fx["noise"]))
    ais = OracleAIS(flow, tgt.log_prob, op, p_target=cfg["p_target"], alpha=cfg["alpha"],
                    n_intermediate_distributions=cfg["M"], distribution_spacing_type=cfg["spacing"])
    flow._eps_override = fx["noise"]["base_eps"][0]
    pt, lw = ais.sample_and_log_weights(cfg["B"])
    ref = fx["ref"]
    same = (pt.x - ref["point"]["x"]).abs().max(dim=1).values <= 1e-3 * (1 + ref["point"]["x"].abs().max(dim=1).values)
    assert same.float().mean() >= 0.9
    assert torch.allclose(lw[same], ref["log_w"][same], rtol=1e-3, atol=1e-3)
    for s in fx["steps"]:
        assert s["after"]["x"].dtype == torch.float32 and "fp64_after" in s


def test_flow_self_consistency():
    """Flow parity is unpinned by the reference (oracle/__init__.py) -> self-consistency checks."""
    torch.manual_seed(0)
    flow = OracleRealNVP(7, 4, 6).double()
    randomize_last_layers(flow, 0.3)
    x, log_q = flow.sample_and_log_prob((200,))
    assert torch.allclose(flow.log_prob(x), log_q, atol=1e-10)              # log_prob(sample) = log_q
    z = x
    for f in reversed(flow._nf_model.flows):
        z, _ = f.inverse(z)
    y = z
    for f in flow._nf_model.flows:
        y, _ = f(y)
    assert torch.allclose(y, x, atol=1e-10)                                 # forward(inverse) = id
    xg = x[:5].detach().clone().requires_grad_(True)
    jac = torch.autograd.functional.jacobian(
        lambda t: _inverse_only(flow, t), xg[0:1]).reshape(7, 7)
    z0, ld = _inverse_with_logdet(flow, xg[0:1])
    assert abs(float(torch.linalg.slogdet(jac)[1]) - float(ld)) < 1e-9      # log-det vs Jacobian
    lq, g = analytic_log_prob_and_grad(flow, x[:20].detach())
    x2 = x[:20].detach().clone().requires_grad_(True)
    lq2 = flow.log_prob(x2)
    g2 = torch.autograd.grad(lq2.sum(), x2)[0]
    assert torch.allclose(lq, lq2, atol=1e-10) and torch.allclose(g, g2, atol=1e-9)


def _inverse_only(flow, x):
    z = x
    for f in reversed(flow._nf_model.flows):
        z, _ = f.inverse(z)
    return z


def _inverse_with_logdet(flow, x):
    z, tot = x, 0.0
    for f in reversed(flow._nf_model.flows):
        z, ld = f.inverse(z)
        tot = tot + ld
    return z, tot


def test_ess_matches_definition():
    lw = torch.randn(100, dtype=torch.float64)
    w = torch.softmax(lw, 0)
    assert abs(float(effective_sample_size(lw)) - float(1 / (w ** 2).sum() / 100)) < 1e-12
