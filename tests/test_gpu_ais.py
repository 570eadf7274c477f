"""Whole-chain `sample_and_log_weights` on the GPU vs. the fp64 oracle with injected noise."""
import copy
import math

import numpy as np
import pytest
import torch

import fab_torch_b200 as fb
from helpers import make_flows, make_manywell, make_gmm, make_aldp, rel_err
from oracle.noise import Float32RecordingNoise
from oracle.sampler import OracleAIS, OracleHMC, OracleMetropolis

pytestmark = pytest.mark.gpu


def _run_pair(dim, K, npd, tk, M, B, op_kind, spacing="linear", p_target=False, alpha=2.0,
              last_std=0.05, act_norm=False, **opkw):
    fo64, fo, fp = make_flows(dim, K, npd, last_std=last_std, act_norm=act_norm)
    if tk == "mw":
        to, tp = make_manywell(dim)
        to32 = to
    elif tk == "aldp":
        to, to32, tp = make_aldp(dim)
    else:
        to, to32, tp = make_gmm(dim, 4, 8.0)
    if op_kind == "hmc":
        op_o = OracleHMC(M, dim, fo64.log_prob, to.log_prob, alpha=alpha, p_target=p_target, **opkw).double()
        op_p = fb.HamiltonianMonteCarlo(M, dim, fp.log_prob, tp.log_prob, alpha=alpha, p_target=p_target, **opkw).cuda()
    else:
        op_o = OracleMetropolis(M, dim, fo64.log_prob, to.log_prob, alpha=alpha, p_target=p_target, **opkw).double()
        op_p = fb.Metropolis(M, dim, fp.log_prob, tp.log_prob, alpha=alpha, p_target=p_target, **opkw).cuda()
    noise = Float32RecordingNoise()
    op_o.noise = noise
    ais_o = OracleAIS(fo64, to.log_prob, op_o, p_target=p_target, alpha=alpha,
                      n_intermediate_distributions=M, distribution_spacing_type=spacing)
    torch.manual_seed(21)
    fo64._eps_override = noise.base_eps(B, dim, torch.float64, "cpu")
    pt_o, lw_o = ais_o.sample_and_log_weights(B)
    # yardstick: the reference algorithm itself in fp32 on the CPU, same noise
    if op_kind == "hmc":
        op_32 = OracleHMC(M, dim, fo.log_prob, to32.log_prob, alpha=alpha, p_target=p_target, **opkw)
    else:
        op_32 = OracleMetropolis(M, dim, fo.log_prob, to32.log_prob, alpha=alpha, p_target=p_target, **opkw)
    from oracle.noise import ReplayNoise
    op_32.noise = ReplayNoise(copy.deepcopy(noise.record))
    ais_32 = OracleAIS(fo, to32.log_prob, op_32, p_target=p_target, alpha=alpha,
                       n_intermediate_distributions=M, distribution_spacing_type=spacing)
    fo._eps_override = noise.record["base_eps"][0]
    _run_pair.cpu32 = ais_32.sample_and_log_weights(B)
    ais_p = fb.AnnealedImportanceSampler(fp, tp.log_prob, op_p, p_target=p_target, alpha=alpha,
                                         n_intermediate_distributions=M,
                                         distribution_spacing_type=spacing)
    op_p.noise = fb.InjectedNoise(noise.record)
    pt_p, lw_p = ais_p.sample_and_log_weights(B)
    return (pt_o, lw_o, ais_o, op_o), (pt_p, lw_p, ais_p, op_p)


AIS_CASES = [
    dict(dim=32, K=10, npd=10, tk="mw", M=4, B=96, op_kind="hmc", epsilon=0.05, L=3),
    dict(dim=32, K=10, npd=10, tk="mw", M=3, B=200, op_kind="hmc", epsilon=1.0, L=5),     # reference init step: mostly rejects / overflow
    dict(dim=2, K=4, npd=40, tk="gmm", M=8, B=512, op_kind="metropolis", n_updates=1,
         max_step_size=5.0, min_step_size=5.0, adjust_step_size=False),                     # config 1
    dict(dim=2, K=0, npd=1, tk="gmm", M=12, B=300, op_kind="hmc", spacing="geometric",
         p_target=True, alpha=None, epsilon=0.5, L=4, n_outer=2),
    # BASELINE config 4 architecture (Many-Well-128, 10 layers x width 1280: the 8-slot tile
    # layout, 160 MB of weights streamed from HBM/L2), shortened chain
    dict(dim=128, K=10, npd=10, tk="mw", M=2, B=24, op_kind="hmc", epsilon=0.02, L=2, last_std=0.003),
    # BASELINE config 5: 60-dof ALDP surrogate energy, 20 distributions, HMC L=4 (fab_buff.yaml:43)
    dict(dim=60, K=4, npd=5, tk="aldp", M=20, B=64, op_kind="hmc", epsilon=0.05, L=4),
    # ActNorm after every InvertibleAffine (make_normflow_model.py:28-29), folded into the packed weights
    dict(dim=32, K=5, npd=10, tk="mw", M=4, B=160, op_kind="hmc", epsilon=0.05, L=3, act_norm=True),
    dict(dim=2, K=3, npd=20, tk="gmm", M=5, B=256, op_kind="metropolis", n_updates=2, act_norm=True),
    dict(dim=32, K=5, npd=10, tk="mw", M=4, B=160, op_kind="hmc", epsilon=0.05, L=3, act_norm=True, engine="rowtile"),
]


@pytest.mark.parametrize("case", AIS_CASES)
def test_chain_parity(case, monkeypatch):
    case = dict(case)
    monkeypatch.setenv("FAB_ENGINE", case.pop("engine", "auto"))
    (pt_o, lw_o, ais_o, op_o), (pt_p, lw_p, ais_p, op_p) = _run_pair(**case)
    assert lw_p.shape == lw_o.shape and pt_p.x.shape == pt_o.x.shape
    B = lw_o.shape[0]
    dx = (pt_p.x.cpu().double() - pt_o.x).abs().max(dim=1).values
    diverged = dx > 1e-2 * (1 + pt_o.x.abs().max(dim=1).values)
    frac = diverged.float().mean().item()
    if case["op_kind"] == "metropolis":
        # exp-overflow -> reject rule is dtype dependent: the fp32 CPU run is the like-for-like truth
        pt_o, lw_o = _run_pair.cpu32
        lw_o = lw_o.double()
        dx = (pt_p.x.cpu().double() - pt_o.x.double()).abs().max(dim=1).values
        diverged = dx > 1e-2 * (1 + pt_o.x.double().abs().max(dim=1).values)
        frac = diverged.float().mean().item()
    assert frac <= 0.02, f"{frac:.3f} of the chains took a different accept branch"
    ok = ~diverged
    err_w = rel_err(lw_p.cpu()[ok], lw_o[ok])
    # the same algorithm in fp32 on the CPU (the reference's own arithmetic) sets the scale of
    # what rounding alone does to this chain; the CUDA path must not be worse than ~that.
    pt_32, lw_32 = _run_pair.cpu32
    dx32 = (pt_32.x.double() - pt_o.x).abs().max(dim=1).values
    ok32 = ~(dx32 > 1e-2 * (1 + pt_o.x.abs().max(dim=1).values))
    err_32 = rel_err(lw_32[ok32], lw_o[ok32])
    print(f"log_w rel err vs fp64 truth: cuda {err_w:.3e}, cpu-fp32 reference {err_32:.3e}; "
          f"diverged chains: cuda {int(diverged.sum())}, cpu-fp32 {int((~ok32).sum())} of {B}")
    if case["op_kind"] == "metropolis":
        assert err_w < 5e-5, f"log_w rel err {err_w:.3e} vs the fp32 reference run"
        return
    # all but 1 % of the chains meet the bar; the rest (ReLU-kink / near-threshold chains that
    # have not fully diverged) stay within 10x of it
    e_rows = ((lw_p.cpu().double()[ok] - lw_o[ok]).abs() / lw_o[ok].abs().clamp_min(1.0))
    bar = max(1e-5, 4 * err_32)
    n_bad = int((e_rows > bar).sum())
    assert n_bad <= max(1, int(0.01 * e_rows.numel())) and err_w < 10 * bar, \
        f"log_w rel err {err_w:.3e}, {n_bad} chains above {bar:.3e} (cpu fp32: {err_32:.3e})"
    info_o, info_p = ais_o.get_logging_info(), ais_p.get_logging_info()
    assert set(info_o) == set(info_p)
    assert abs(info_p["ess_base"] - info_o["ess_base"]) < 1e-4 * max(info_o["ess_base"], 1e-3) + 1e-7
    if frac == 0:
        assert abs(info_p["log_Z"] - info_o["log_Z"]) < 1e-3
        assert abs(info_p["ess_ais"] - info_o["ess_ais"]) < 1e-3 * max(info_o["ess_ais"], 1e-3) + 1e-7


def test_log_z_of_normalised_gaussians():
    """fab/sampling_methods/ais_test.py:21-64 made into an assertion: q=N(+.5,I), p=N(-.5,I),
    alpha=1, fixed Metropolis step 2.0 -> log Z = 0; the error shrinks with more distributions."""
    dim, B = 1, 10000
    dim = 2   # kernels need d >= 2; the analytic answer is unchanged
    errs = {}
    for M in (1, 8, 32):
        flow = fb.B200RealNVP(dim, 0, 1)
        with torch.no_grad():
            flow._nf_model.q0.loc.fill_(0.5)
        flow = flow.cuda()
        target = fb.DiagGaussianTarget(torch.zeros(dim) - 0.5, 1.0)
        op = fb.Metropolis(M, dim, flow.log_prob, target.log_prob, n_updates=1, alpha=1.0,
                           p_target=False, max_step_size=2.0, min_step_size=2.0,
                           adjust_step_size=False).cuda()
        ais = fb.AnnealedImportanceSampler(flow, target.log_prob, op, p_target=False, alpha=1.0,
                                           n_intermediate_distributions=M)
        torch.manual_seed(0)
        pts, log_w = ais.sample_and_log_weights(B)
        log_Z = torch.logsumexp(log_w, 0) - math.log(B)
        errs[M] = abs(log_Z.item())
        assert abs(ais.get_logging_info()["log_Z"] - log_Z.item()) < 1e-4
    assert errs[32] < 0.05 and errs[8] < 0.1 and errs[1] < 0.3, errs


def test_nan_filter_shrinks_batch():
    """ais.py:190-213: particles with non-finite log_p/log_q are dropped (stable order) and log Z
    still divides by the requested batch size (quirk 6)."""
    fo64, fo, fp = make_flows(2, 0, 1, perturb_base=False)
    with torch.no_grad():
        fp._nf_model.q0.log_scale.fill_(math.log(60.0))      # base much wider than the target mask
    to64, to, tp = make_gmm(2, 4, 8.0)
    op = fb.Metropolis(3, 2, fp.log_prob, tp.log_prob, n_updates=1, alpha=2.0, p_target=False).cuda()
    ais = fb.AnnealedImportanceSampler(fp, tp.log_prob, op, p_target=False, alpha=2.0,
                                       n_intermediate_distributions=3)
    B = 400
    torch.manual_seed(3)
    eps = torch.randn(B, 2)
    fp._eps_override = eps.cuda()
    pt, lw = ais.sample_and_log_weights(B)
    x0 = (eps * 60.0).double()
    keep = torch.isfinite(to64.log_prob(x0))
    assert 0 < keep.sum() < B
    assert lw.shape[0] == int(keep.sum()) == pt.x.shape[0]
    assert torch.isfinite(pt.log_p).all() and torch.isfinite(pt.log_q).all()
    info = ais.get_logging_info()
    lz = torch.logsumexp(lw, 0).item() - math.log(B)
    assert abs(info["log_Z"] - lz) < 1e-4


def test_all_invalid_raises():
    _, _, fp = make_flows(2, 0, 1, perturb_base=False)
    with torch.no_grad():
        fp._nf_model.q0.loc.fill_(1.0e4)
    _, _, tp = make_gmm(2, 4, 8.0)
    op = fb.Metropolis(2, 2, fp.log_prob, tp.log_prob, n_updates=1, alpha=2.0).cuda()
    ais = fb.AnnealedImportanceSampler(fp, tp.log_prob, op, p_target=False, alpha=2.0,
                                       n_intermediate_distributions=2)
    with pytest.raises(Exception, match="No valid points generated in sampling the chain init"):
        ais.sample_and_log_weights(64)


@pytest.mark.parametrize("op_kind", ["hmc", "metropolis"])
def test_cuda_graph_chain_equals_eager(op_kind):
    """use_cuda_graph=True (warm-up call, capture, replays) must reproduce the eager chain bit for
    bit on the same seed -- including after a parameter update (the weight blob is repacked in
    place, the captured graph stays valid) and after set_ais_target-style p_target flips."""
    dim, K, npd, M, B = (32, 4, 6, 5, 300) if op_kind == "hmc" else (2, 4, 40, 8, 512)

    def build(graph):
        _, _, fp = make_flows(dim, K, npd, last_std=0.02)
        if op_kind == "hmc":
            _, tp = make_manywell(dim)
            op = fb.HamiltonianMonteCarlo(M, dim, fp.log_prob, tp.log_prob, alpha=2.0, p_target=False,
                                          epsilon=0.1, L=3, n_outer=2).cuda()
        else:
            _, _, tp = make_gmm(dim, 4, 8.0)
            op = fb.Metropolis(M, dim, fp.log_prob, tp.log_prob, n_updates=2, alpha=2.0, p_target=False,
                               max_step_size=1.0, min_step_size=0.5).cuda()
        ais = fb.AnnealedImportanceSampler(fp, tp.log_prob, op, p_target=False, alpha=2.0,
                                           n_intermediate_distributions=M, use_cuda_graph=graph)
        return fp, op, ais

    (f_e, op_e, ais_e), (f_g, op_g, ais_g) = build(False), build(True)
    for step in range(5):
        if step == 3:            # "optimiser step": same parameter update on both sides
            with torch.no_grad():
                for f in (f_e, f_g):
                    f._nf_model.q0.loc.add_(0.05)
                    f._nf_model.flows[0].linears[2].weight.mul_(1.1)
        if step == 4:            # FABModel.set_ais_target(min_is_target=False)
            for a, o in ((ais_e, op_e), (ais_g, op_g)):
                a.p_target = o.p_target = True
        outs = []
        for ais in (ais_e, ais_g):
            torch.manual_seed(100 + step)
            outs.append(ais.sample_and_log_weights(B))
        (pt_e, lw_e), (pt_g, lw_g) = outs
        assert torch.equal(lw_e, lw_g) and torch.equal(pt_e.x, pt_g.x) and torch.equal(pt_e.log_q, pt_g.log_q)
        assert ais_e.get_logging_info() == ais_g.get_logging_info()
    assert any(e["graph"] is not None for e in ais_g._graphs.values())
    # pre-drawn host noise goes straight into the graph's persistent buffers
    g = torch.Generator().manual_seed(9)
    n2 = op_e.n_outer if op_kind == "hmc" else op_e.n_updates
    eps = torch.randn(B, dim, generator=g).pin_memory()
    a = torch.randn(M, n2, B, dim, generator=g).pin_memory()
    b = (torch.empty(M, n2, B).exponential_(1.0, generator=g) if op_kind == "hmc"
         else torch.rand(M, n2, B, generator=g)).pin_memory()
    outs = []
    for ais in (ais_e, ais_g):
        ais.set_next_noise(eps, a, b)
        outs.append(ais.sample_and_log_weights(B))
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][0].x, outs[1][0].x)


@pytest.mark.parametrize("engine,dim,K,npd,B,n_outer", [("warp", 32, 4, 6, 300, 2), ("rowtile", 32, 3, 10, 1280, 1),
                                                          ("warp", 6, 2, 5, 77, 1)])
def test_c_chain_equals_python_loop(engine, dim, K, npd, B, n_outer, monkeypatch):
    """`fab_ais_chain_hmc_f32` (the whole ais.py:53-87 loop as ONE C-ABI call) against the
    launch-by-launch loop of `_run_chain` on the same seeds: bit-identical particles, log-weights,
    logging record and tuner state, over several calls (tuner on), with NaN-producing particles in the
    batch (filter) and for n_outer > 1 (the proposal carried into the next outer step)."""
    monkeypatch.setenv("FAB_ENGINE", engine)
    M = 5

    def build():
        _, _, fp = make_flows(dim, K, npd, last_std=0.02)
        _, tp = make_manywell(dim)
        op = fb.HamiltonianMonteCarlo(M, dim, fp.log_prob, tp.log_prob, alpha=2.0, p_target=False,
                                      epsilon=0.1, L=3, n_outer=n_outer).cuda()
        return fp, op, fb.AnnealedImportanceSampler(fp, tp.log_prob, op, p_target=False, alpha=2.0,
                                                    n_intermediate_distributions=M)

    (f_c, op_c, ais_c), (f_p, op_p, ais_p) = build(), build()
    for step in range(4):
        outs = []
        for ais, flag in ((ais_c, "1"), (ais_p, "0")):
            monkeypatch.setenv("FAB_C_CHAIN", flag)
            torch.manual_seed(300 + step)
            if step == 2:          # a few broken base samples: the chain-init filter drops them
                eps = torch.randn(B, dim)
                eps[3] = float("nan"); eps[B // 2] = float("inf")
                ais.base_distribution._eps_override = eps.cuda()
            outs.append(ais.sample_and_log_weights(B))
        (pt_c, lw_c), (pt_p, lw_p) = outs
        assert lw_c.shape == lw_p.shape and (step != 2 or lw_c.shape[0] < B)
        assert torch.equal(lw_c, lw_p) and torch.equal(pt_c.x, pt_p.x) and torch.equal(pt_c.log_q, pt_p.log_q)
        assert torch.equal(pt_c.grad_log_p, pt_p.grad_log_p)
        assert ais_c.get_logging_info() == ais_p.get_logging_info()
        assert torch.equal(op_c.epsilons, op_p.epsilons) and torch.equal(op_c.common_epsilon, op_p.common_epsilon)
