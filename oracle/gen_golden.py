"""Pin the oracle against the UNMODIFIED reference and write the golden fixtures.  TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference):   python -m oracle.gen_golden

1. PIN: with identical seeds the restatement in oracle/ must reproduce the reference's own
   classes (imported from /root/reference via oracle/ref_loader.py) BIT FOR BIT on CPU, in fp32
   and fp64: ManyWellEnergy / GMM log_prob, effective_sample_size, the beta grids, and whole
   `AnnealedImportanceSampler.sample_and_log_weights` runs with `HamiltonianMonteCarlo` and
   `Metropolis` (tuner on), including the reference's own test fixture `setup_ais`
   (fab/sampling_methods/ais_test.py:86-126).  The flow handed to the reference classes is
   oracle.realnvp.OracleRealNVP (normflows itself is not installable here -- flow arithmetic is
   "parity unpinned", see oracle/__init__.py).
2. FIXTURES (tests/golden/*.pt): outputs of the reference classes themselves (fp32) together with
   the noise they consumed (recorded through the pinned oracle run with the same seed) and the
   fp64 ground truth obtained by replaying that noise through the oracle in float64.
   `*_steps` fixtures additionally hold, per transition, the fp32 input state of the reference
   chain and the fp64 one-transition result from exactly that state (teacher forcing).
"""
import copy
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.noise import RecordingNoise, ReplayNoise          # noqa: E402
from oracle.realnvp import OracleRealNVP, randomize_last_layers  # noqa: E402
from oracle.ref_loader import load_reference                   # noqa: E402
from oracle.sampler import (OracleAIS, OracleHMC, OracleMetropolis, Point, beta_schedule,  # noqa: E402
                            effective_sample_size)
from oracle.targets import OracleGMM, OracleManyWell, to_double  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
REPORT = []


def check(name, ok):
    REPORT.append({"check": name, "ok": bool(ok)})
    print(("PASS " if ok else "FAIL ") + name)
    assert ok, name


def build_flow(dim, K, npd, seed, last_std=0.05, base_scale=None):
    torch.manual_seed(seed)
    f = OracleRealNVP(dim, K, npd)
    if K:
        randomize_last_layers(f, last_std, seed=seed + 1)
    if base_scale is not None:
        with torch.no_grad():
            f._nf_model.q0.log_scale.fill_(float(np.log(base_scale)))
    return f


def flow_checksum(flow):
    return float(sum(p.detach().double().abs().sum() for p in flow.state_dict().values()))


def pt_dict(pt, dtype=None):
    c = lambda t: None if t is None else (t.detach().clone() if dtype is None else t.detach().to(dtype))
    return dict(x=c(pt.x), log_q=c(pt.log_q), log_p=c(pt.log_p), grad_log_q=c(pt.grad_log_q),
                grad_log_p=c(pt.grad_log_p))


def pt_equal(a, b):
    names = ("x", "log_q", "log_p", "grad_log_q", "grad_log_p")
    return all((getattr(a, n) is None and getattr(b, n) is None) or
               torch.equal(getattr(a, n), getattr(b, n)) for n in names)


def make_ops(fab, kind, M, dim, flow_ref, flow_ora, tgt_ref, tgt_ora, p_target, alpha, opkw):
    from fab.sampling_methods import HamiltonianMonteCarlo, Metropolis
    if kind == "hmc":
        op_r = HamiltonianMonteCarlo(M, dim, flow_ref.log_prob, tgt_ref.log_prob, alpha=alpha,
                                     p_target=p_target, **opkw)
        op_o = OracleHMC(M, dim, flow_ora.log_prob, tgt_ora.log_prob, alpha=alpha,
                         p_target=p_target, **opkw)
    else:
        op_r = Metropolis(M, dim, flow_ref.log_prob, tgt_ref.log_prob, alpha=alpha,
                          p_target=p_target, **opkw)
        op_o = OracleMetropolis(M, dim, flow_ora.log_prob, tgt_ora.log_prob, alpha=alpha,
                                p_target=p_target, **opkw)
    return op_r, op_o


def run_case(fab, name, *, dim, K, npd, target, M, B, kind, opkw, p_target=False, alpha=2.0,
             spacing="linear", flow_seed=0, run_seed=1234, n_calls=1, steps=False,
             base_scale=None, dtype=torch.float32):
    """Reference vs oracle (bitwise) for `n_calls` consecutive calls; returns the fixture of the
    last call."""
    from fab.sampling_methods import AnnealedImportanceSampler
    from fab.target_distributions.many_well import ManyWellEnergy
    from fab.target_distributions.gmm import GMM
    torch.set_default_dtype(dtype)
    try:
        flow = build_flow(dim, K, npd, flow_seed, base_scale=base_scale)
        if dtype == torch.float64:
            flow = flow.double()
        if target[0] == "mw":
            tgt_ref = ManyWellEnergy(dim, use_gpu=False)
            tgt_ora = OracleManyWell(dim)
        else:
            _, n_mixes, loc_scaling, lvs = target
            torch.manual_seed(0)
            tgt_ref = GMM(dim=dim, n_mixes=n_mixes, loc_scaling=loc_scaling, log_var_scaling=lvs,
                          use_gpu=False, true_expectation_estimation_n_samples=1000)
            torch.manual_seed(0)
            tgt_ora = OracleGMM(dim, n_mixes, loc_scaling, lvs)
            check(f"{name}: GMM parameters identical", torch.equal(tgt_ref.locs, tgt_ora.locs) and
                  torch.equal(tgt_ref.scale_trils, tgt_ora.scale_trils))
        op_r, op_o = make_ops(fab, kind, M, dim, flow, flow, tgt_ref, tgt_ora, p_target, alpha, opkw)
        ais_r = AnnealedImportanceSampler(flow, tgt_ref.log_prob, op_r, p_target=p_target,
                                          alpha=alpha, n_intermediate_distributions=M,
                                          distribution_spacing_type=spacing)
        ais_o = OracleAIS(flow, tgt_ora.log_prob, op_o, p_target=p_target, alpha=alpha,
                          n_intermediate_distributions=M, distribution_spacing_type=spacing)
        check(f"{name}: beta grid identical", torch.equal(ais_r.B_space, ais_o.B_space))
        fixture = None
        for call in range(n_calls):
            state_before = {k: v.clone() for k, v in op_o.state_dict().items()}
            torch.manual_seed(run_seed + call)
            pt_r, lw_r = ais_r.sample_and_log_weights(B)
            info_r = ais_r.get_logging_info()
            rec = RecordingNoise()
            op_o.noise = rec
            snaps = []
            if steps:
                orig_step = ais_o._step

                def spy(pt, log_w, j, _orig=orig_step):
                    before = (pt_dict(pt), log_w.detach().clone(),
                              {k: v.clone() for k, v in op_o.state_dict().items()})
                    out_pt, out_w = _orig(pt, log_w, j)
                    snaps.append(dict(j=j, before=before[0], log_w_before=before[1],
                                      op_state_before=before[2], after=pt_dict(out_pt),
                                      log_w_after=out_w.detach().clone()))
                    return out_pt, out_w
                ais_o._step = spy
            torch.manual_seed(run_seed + call)
            flow._eps_override = rec.base_eps(B, dim, dtype, "cpu")
            pt_o, lw_o = ais_o.sample_and_log_weights(B)
            if steps:
                ais_o._step = orig_step
            info_o = ais_o.get_logging_info()
            same_state = all(torch.equal(op_r.state_dict()[k], op_o.state_dict()[k])
                             for k in op_o.state_dict())
            check(f"{name} [{str(dtype)[6:]}, call {call}]: oracle == reference bit for bit "
                  f"(point, log_w, logging info, tuner state)",
                  pt_equal(pt_r, pt_o) and torch.equal(lw_r, lw_o) and info_r == info_o and same_state)
            fixture = dict(
                name=name, config=dict(dim=dim, K=K, npd=npd, target=list(target), M=M, B=B,
                                       kind=kind, opkw=opkw, p_target=p_target, alpha=alpha,
                                       spacing=spacing, flow_seed=flow_seed, base_scale=base_scale),
                flow_checksum=flow_checksum(flow), op_state_before=state_before,
                noise={k: v for k, v in rec.record.items() if v},
                ref=dict(point=pt_dict(pt_r), log_w=lw_r.clone(), info=info_r,
                         op_state_after={k: v.clone() for k, v in op_r.state_dict().items()}),
                steps=snaps)
        return fixture, (flow, tgt_ora)
    finally:
        torch.set_default_dtype(torch.float32)


def add_fp64_truth(fixture, flow, tgt_ora):
    """Replay the recorded fp32 noise through the oracle in float64 (ground truth)."""
    cfg = fixture["config"]
    flow64 = copy.deepcopy(flow).double()
    tgt64 = to_double(copy.deepcopy(tgt_ora))
    dim, M = cfg["dim"], cfg["M"]
    cls = OracleHMC if cfg["kind"] == "hmc" else OracleMetropolis
    def fresh_op(state):
        op = cls(M, dim, flow64.log_prob, tgt64.log_prob, alpha=cfg["alpha"],
                 p_target=cfg["p_target"], **cfg["opkw"])
        op.load_state_dict(state)
        return op.double()
    op = fresh_op(fixture["op_state_before"])
    op.noise = ReplayNoise(copy.deepcopy(fixture["noise"]))
    ais = OracleAIS(flow64, tgt64.log_prob, op, p_target=cfg["p_target"], alpha=cfg["alpha"],
                    n_intermediate_distributions=M, distribution_spacing_type=cfg["spacing"])
    flow64._eps_override = fixture["noise"]["base_eps"][0].double()
    pt, lw = ais.sample_and_log_weights(cfg["B"])
    fixture["fp64"] = dict(point=pt_dict(pt), log_w=lw.clone(), info=ais.get_logging_info())
    # teacher-forced single transitions from the reference chain's own fp32 states
    n_per = cfg["opkw"].get("n_outer", 1) if cfg["kind"] == "hmc" else cfg["opkw"]["n_updates"]
    for s in fixture["steps"]:
        j = s["j"]
        op = fresh_op(s["op_state_before"])
        lo = (j - 1) * n_per
        if cfg["kind"] == "hmc":
            rec = dict(momentum=fixture["noise"]["momentum"][lo:lo + n_per],
                       exponential=fixture["noise"]["exponential"][lo:lo + n_per])
        else:
            rec = dict(proposal=fixture["noise"]["proposal"][lo:lo + n_per],
                       uniform=fixture["noise"]["uniform"][lo:lo + n_per])
        op.noise = ReplayNoise(rec)
        ais1 = OracleAIS(flow64, tgt64.log_prob, op, p_target=cfg["p_target"], alpha=cfg["alpha"],
                         n_intermediate_distributions=M, distribution_spacing_type=cfg["spacing"])
        b = s["before"]
        d64 = lambda t: None if t is None else t.double().clone()
        pt0 = Point(d64(b["x"]), d64(b["log_q"]), d64(b["log_p"]), d64(b["grad_log_q"]),
                    d64(b["grad_log_p"]))
        pt1, lw1 = ais1._step(pt0, s["log_w_before"].double(), j)
        s["fp64_after"] = pt_dict(pt1)
        s["fp64_log_w_after"] = lw1.clone()
        s["fp64_op_state_after"] = {k: v.clone() for k, v in op.state_dict().items()}
    return fixture


def pin_pointwise(fab):
    from fab.target_distributions.many_well import ManyWellEnergy
    from fab.target_distributions.gmm import GMM
    from fab.utils.numerical import effective_sample_size as ref_ess
    from fab.sampling_methods import AnnealedImportanceSampler
    for dtype in (torch.float32, torch.float64):
        g = torch.Generator().manual_seed(7)
        for dim in (2, 32, 128):
            x = (torch.randn(64, dim, generator=g) * 1.5).to(dtype)
            check(f"ManyWellEnergy.log_prob d={dim} {dtype}",
                  torch.equal(ManyWellEnergy(dim, use_gpu=False).log_prob(x),
                              OracleManyWell(dim).log_prob(x)))
        lw = torch.randn(1000, generator=g).to(dtype) * 3
        check(f"effective_sample_size {dtype}", torch.equal(ref_ess(lw), effective_sample_size(lw)))
    check("ManyWell log_Z(d=32) = 164.69567532",
          abs(float(ManyWellEnergy(32, use_gpu=False).log_Z) - 164.69567532) < 1e-7 and
          abs(float(OracleManyWell(32).log_Z) - 164.69567532) < 1e-7)
    check("ManyWell log_Z(d=128) = 658.7827013",
          abs(float(OracleManyWell(128).log_Z) - 658.7827013) < 1e-6)
    torch.manual_seed(0)
    gr = GMM(dim=2, n_mixes=40, loc_scaling=40, log_var_scaling=1.0, use_gpu=False,
             true_expectation_estimation_n_samples=1000)
    torch.manual_seed(0)
    go = OracleGMM(2, 40, 40, 1.0)
    x = (torch.rand(500, 2) - 0.5) * 100
    check("GMM-40 locs/scale identical + first means match SURVEY §4",
          torch.equal(gr.locs, go.locs) and
          torch.allclose(go.locs[:3], torch.tensor([[-0.29947, 21.45774], [-32.92181, -29.43756],
                                                    [-15.40617, 10.72629]]), atol=1e-4) and
          abs(float(go.scale_trils[0, 0, 0]) - 1.3132616) < 1e-6)
    check("GMM.log_prob identical (incl. -inf mask)", torch.equal(gr.log_prob(x), go.log_prob(x)))
    for kind in ("linear", "geometric"):
        for M in (1, 4, 16, 40):
            class _A(AnnealedImportanceSampler):
                def __init__(self, M):
                    self.n_intermediate_distributions = M
            check(f"beta grid {kind} M={M}",
                  torch.equal(_A(M).setup_distribution_spacing(kind, M), beta_schedule(kind, M)))


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(1)          # fixed reduction order for the bitwise comparisons
    fab = load_reference()
    pin_pointwise(fab)
    hmc_c2 = dict(epsilon=1.0, n_outer=1, L=5)
    cases = [
        # name, kwargs, also run in fp64?, add teacher-forced steps?
        ("hmc_manywell32_chain", dict(dim=32, K=10, npd=10, target=("mw",), M=16, B=64, kind="hmc",
                                      opkw=hmc_c2, n_calls=2), True),
        ("hmc_manywell32_steps", dict(dim=32, K=10, npd=10, target=("mw",), M=6, B=32, kind="hmc",
                                      opkw=dict(epsilon=0.12, n_outer=1, L=5), steps=True), True),
        ("hmc_manywell8_ptarget", dict(dim=8, K=3, npd=6, target=("mw",), M=5, B=128, kind="hmc",
                                       opkw=dict(epsilon=0.2, n_outer=2, L=3), p_target=True,
                                       alpha=None, steps=True), True),
        ("metropolis_gmm40_c1", dict(dim=2, K=4, npd=40, target=("gmm", 40, 40.0, 1.0), M=8, B=512,
                                     kind="metropolis",
                                     opkw=dict(n_updates=1, max_step_size=5.0, min_step_size=5.0,
                                               adjust_step_size=False), steps=True), True),
        ("metropolis_gmm40_tuned", dict(dim=2, K=4, npd=40, target=("gmm", 40, 40.0, 1.0), M=4,
                                        B=256, kind="metropolis",
                                        opkw=dict(n_updates=3, max_step_size=5.0, min_step_size=1.0),
                                        steps=True, n_calls=2), True),
        # the reference's own fixture setup_ais (ais_test.py:86-126): GMM 4-mix loc 8, base
        # N(0, 9 I), M=40 geometric, HMC n_outer=5 L=5 eps=1 / Metropolis n_updates=5
        ("refsetup_hmc", dict(dim=2, K=0, npd=1, target=("gmm", 4, 8.0, 0.1), M=40, B=64,
                              kind="hmc", opkw=dict(n_outer=5, epsilon=1.0, L=5), p_target=True,
                              alpha=None, spacing="geometric", base_scale=3.0, steps=True), True),
        ("refsetup_metropolis", dict(dim=2, K=0, npd=1, target=("gmm", 4, 8.0, 0.1), M=40, B=64,
                                     kind="metropolis", opkw=dict(n_updates=5), p_target=True,
                                     alpha=None, spacing="geometric", base_scale=3.0, steps=True),
         True),
    ]
    for name, kw, also64 in cases:
        if also64:
            run_case(fab, name, dtype=torch.float64, **{**kw, "steps": False})
        fx, (flow, tgt) = run_case(fab, name, dtype=torch.float32, **kw)
        fx = add_fp64_truth(fx, flow, tgt)
        path = os.path.join(GOLDEN, name + ".pt")
        torch.save(fx, path)
        print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")
    with open(os.path.join(GOLDEN, "pin_report.json"), "w") as fh:
        json.dump(dict(torch=torch.__version__, reference="lollcat/fab-torch @ b1586a92",
                       checks=REPORT), fh, indent=1)
    print(f"{len(REPORT)} pin checks passed")


if __name__ == "__main__":
    main()
