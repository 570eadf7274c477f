"""Drop-in behaviour against the reference's own `FABModel` (fab/core.py).

* CPU (build container only, needs /root/reference): the B200 objects are accepted by the
  UNMODIFIED `fab.core.FABModel` at both insertion levels of INTEGRATION.md, `set_ais_target`
  flips the attributes the kernels read, and `save`/`load` round-trip with reference key names.
* GPU: what `FABModel.fab_alpha_div` does (core.py:112-128), run on the device without the
  reference tree: AIS with the alpha-divergence target, then the weighted log q loss and its
  parameter gradient.
"""
import os

import numpy as np
import pytest
import torch

import fab_torch_b200 as fb
from oracle.ref_loader import reference_available, load_reference


def _objects(device=None):
    torch.manual_seed(0)
    flow = fb.B200RealNVP(4, 2, 5)
    target = fb.ManyWellEnergy(4, use_gpu=device is not None)
    op = fb.HamiltonianMonteCarlo(3, 4, flow.log_prob, target.log_prob, alpha=2.0, p_target=False,
                                  n_outer=1, epsilon=0.2, L=2)
    if device is not None:
        flow, op = flow.to(device), op.to(device)
    return flow, target, op


@pytest.mark.skipif(not reference_available(), reason="reference tree only exists in the build container")
def test_unmodified_fabmodel_accepts_b200_objects(tmp_path):
    load_reference()
    import fab.core
    flow, target, op = _objects()
    # level 1: operator-level plugin API
    model = fab.core.FABModel(flow=flow, target_distribution=target, n_intermediate_distributions=3,
                              transition_operator=op, alpha=2.0, loss_type="fab_alpha_div")
    assert type(model.annealed_importance_sampler).__module__ == "fab.sampling_methods.ais"
    assert list(model.parameters())
    model.set_ais_target(min_is_target=False)
    assert op.p_target is True and model.annealed_importance_sampler.p_target is True
    model.set_ais_target(min_is_target=True)
    assert op.p_target is False
    # level 2: sampler-level replacement, before construction (covers FABModel.load as well)
    orig = fab.core.AnnealedImportanceSampler
    fab.core.AnnealedImportanceSampler = fb.AnnealedImportanceSampler
    try:
        model2 = fab.core.FABModel(flow=flow, target_distribution=target,
                                   n_intermediate_distributions=3, transition_operator=op,
                                   alpha=2.0, loss_type="fab_alpha_div")
        assert isinstance(model2.annealed_importance_sampler, fb.AnnealedImportanceSampler)
        assert torch.equal(model2.annealed_importance_sampler.B_space,
                           model.annealed_importance_sampler.B_space)
        path = tmp_path / "model.pt"
        model2.save(path)
        ckpt = torch.load(path, weights_only=False)
        assert set(ckpt) == {"flow", "trans_op"}
        assert "_nf_model.flows.1.log_S" in ckpt["flow"] and "common_epsilon" in ckpt["trans_op"]
        with torch.no_grad():
            flow._nf_model.q0.loc.add_(1.0)
        model2.load(path, map_location="cpu")
        assert float(flow._nf_model.q0.loc.abs().max()) == 0.0
        assert isinstance(model2.annealed_importance_sampler, fb.AnnealedImportanceSampler)
        model2.set_ais_target(min_is_target=False)
        assert model2.annealed_importance_sampler.p_target is True
    finally:
        fab.core.AnnealedImportanceSampler = orig


@pytest.mark.gpu
def test_fab_alpha_div_step_on_device():
    flow, target, op = _objects("cuda")
    ais = fb.AnnealedImportanceSampler(flow, target.log_prob, op, p_target=False, alpha=2.0,
                                       n_intermediate_distributions=3)
    # core.py:120-128
    ais.p_target = op.p_target = False
    point, log_w = ais.sample_and_log_weights(256)
    assert not log_w.requires_grad and point.x.shape == (256, 4)
    log_q_x = flow.log_prob(point.x)
    loss = -np.sign(2.0) * torch.mean(torch.softmax(log_w, dim=-1) * log_q_x)
    loss.backward()
    grads = [p.grad for p in flow.parameters()]
    assert all(g is not None and torch.isfinite(g).all() for g in grads)
    assert sum(float(g.abs().sum()) for g in grads) > 0
    ais.p_target = op.p_target = True
    info = ais.get_logging_info()
    for key in ("ess_base", "ess_ais", "log_Z", "dist0_p_accept_0", "epsilons_dist0_loop0",
                "average_distance_dist0"):
        assert key in info
    # generate_eval_data (ais.py:132-188): shapes asserted by the reference's own test
    # (fab/sampling_methods/ais_test.py:141-144)
    base_x, base_w, ais_x, ais_w = ais.generate_eval_data(200, 100)
    assert base_x.shape == (200, 4) and base_w.shape == (200,)
    assert ais_x.shape == (200, 4) and ais_w.shape == (200,)
    assert base_x.device.type == "cpu"


def _gpu_objects(dim=32, K=3, npd=10, M=4):
    torch.manual_seed(0)
    flow = fb.B200RealNVP(dim, K, npd)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for k in range(K):
            lin = flow._nf_model.flows[2 * k].linears[2]
            lin.weight.copy_(torch.randn(lin.weight.shape, generator=g) * 0.01)
            lin.bias.copy_(torch.randn(lin.bias.shape, generator=g) * 0.01)
    flow = flow.cuda()
    target = fb.ManyWellEnergy(dim, use_gpu=True)
    op = fb.HamiltonianMonteCarlo(M, dim, flow.log_prob, target.log_prob, alpha=2.0, p_target=False,
                                  n_outer=1, epsilon=0.1, L=5).cuda()
    return flow, target, op


@pytest.mark.gpu
def test_unmodified_fabmodel_runs_on_the_b200_plugin_classes():
    """The reference's own fab/core.py (baseline/_ref, unmodified) on a GPU, at both insertion levels
    of INTEGRATION.md: `loss()` = fab_alpha_div (core.py:112-128) with backward, and
    `get_eval_info()` (core.py:191-220).  Level 1 keeps the reference's Python AIS loop and calls
    the B200 operator's `transition` once per intermediate distribution; level 2 swaps the sampler."""
    from oracle.ref_loader import installed_reference_available
    if not installed_reference_available():
        pytest.skip("baseline/_ref not installed (run __graft_entry__.build() where /root/reference exists)")
    load_reference(installed=True)
    import fab.core
    M, B = 4, 1024                       # 1024 particles: the row-tile engine under FAB_ENGINE=auto
    results = {}
    for level in (1, 2):
        flow, target, op = _gpu_objects(M=M)
        orig = fab.core.AnnealedImportanceSampler
        if level == 2:
            fab.core.AnnealedImportanceSampler = fb.AnnealedImportanceSampler
        try:
            model = fab.core.FABModel(flow=flow, target_distribution=target,
                                      n_intermediate_distributions=M, transition_operator=op,
                                      alpha=2.0, loss_type="fab_alpha_div")
        finally:
            fab.core.AnnealedImportanceSampler = orig
        sampler_module = type(model.annealed_importance_sampler).__module__
        assert sampler_module == ("fab.sampling_methods.ais" if level == 1 else "fab_torch_b200.ais")
        torch.manual_seed(123)
        loss = model.loss(B)
        assert loss.ndim == 0 and torch.isfinite(loss)
        loss.backward()
        grads = [p.grad for p in model.parameters()]
        assert all(g is not None and torch.isfinite(g).all() for g in grads)
        assert sum(float(g.abs().sum()) for g in grads) > 0
        assert op._seen_first and op._seen_last            # the fused transitions ran for i = 1..M
        assert op.p_target is True                          # reset by fab_alpha_div (core.py:126-127)
        info = model.get_eval_info(outer_batch_size=512, inner_batch_size=256)
        for key in ("eval_ess_flow", "eval_ess_ais", "flow_test_set_modes_mean_log_prob",
                    "flow_test_set_exact_mean_log_prob", "flow_forward_kl", "ais_abs_MSE_log_Z_estimate"):
            assert key in info, (key, sorted(info))
        assert 0 < info["eval_ess_ais"] <= 1 and 0 < info["eval_ess_flow"] <= 1
        results[level] = (float(loss), info, model.annealed_importance_sampler.get_logging_info())
    # the two levels run the same kernels on different random draws: same order of magnitude
    (l1, i1, log1), (l2, i2, log2) = results[1], results[2]
    assert abs(log1["log_Z"] - log2["log_Z"]) < 25.0, (log1["log_Z"], log2["log_Z"])
    # the mode test set is deterministic for dim < 40 (all 2^16 modes): same flow, same number
    a, b = i1["flow_test_set_modes_mean_log_prob"], i2["flow_test_set_modes_mean_log_prob"]
    assert abs(a - b) < 1e-4 * abs(a) + 1e-4, (a, b)
    print(f"\\nlevel 1: loss {l1:.5f} log_Z {log1['log_Z']:.3f} | level 2: loss {l2:.5f} log_Z {log2['log_Z']:.3f}")
