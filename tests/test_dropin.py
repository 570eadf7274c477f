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
