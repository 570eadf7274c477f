"""GPU: the small kernels around the transitions -- ESS / log Z reduction, NaN/inf filter,
log-weight update, systematic resampling (bit-exact against oracle/resample.py)."""
import math

import numpy as np
import pytest
import torch

import fab_torch_b200 as fb
from fab_torch_b200 import _lib, dist as fdist
from oracle.resample import systematic_ancestors as oracle_ancestors, fixed_point_weights
from oracle.sampler import effective_sample_size as oracle_ess, gamma as oracle_gamma, Point as OPoint, beta_schedule

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 31, 2048, 100003])
def test_effective_sample_size(n):
    g = torch.Generator().manual_seed(n)
    lw = torch.randn(n, generator=g) * 5 + 300.0
    got = fb.effective_sample_size(lw.cuda())
    want = oracle_ess(lw.double())
    assert abs(got.item() - want.item()) <= 1e-5 * want.item()
    # normalised=True branch (numerical.py:21-23)
    w = torch.softmax(lw, 0)
    assert abs(fb.effective_sample_size(w.cuda(), normalised=True).item() - want.item()) <= 1e-4 * want.item()


def test_ess_partials_merge_like_multi_gpu():
    """Two 'ranks' on one device: partial quadruples merged by the kernel == single-shot result."""
    L = _lib.lib()
    lw = (torch.randn(4096) * 3 - 50).cuda()
    halves = [lw[:1000].contiguous(), lw[1000:].contiguous()]
    parts = torch.empty(8, device="cuda")
    for r, h in enumerate(halves):
        _lib.check(L.fab_ess_partial_f32(_lib.ptr(h), None, h.shape[0], None,
                                         _lib.ptr(parts[4 * r:4 * r + 4]), _lib.stream_ptr()))
    out = torch.empty(3, device="cuda")
    _lib.check(L.fab_ess_finalize_f32(_lib.ptr(parts), 2, _lib.ptr(out), _lib.stream_ptr()))
    want = oracle_ess(lw.cpu().double()).item()
    assert abs(out[0].item() - want) <= 1e-5 * want
    assert abs(out[1].item() - torch.logsumexp(lw.cpu().double(), 0).item()) < 1e-4
    assert out[2].item() == 4096
    ess_t, lse_t, cnt_t = fdist.merge_ess_partials(parts.cpu())       # host restatement agrees
    assert abs(float(ess_t) - out[0].item()) <= 1e-5 * want and int(cnt_t) == 4096


def test_nan_filter_compacts_in_order():
    L = _lib.lib()
    n, d = 1000, 5
    g = torch.Generator().manual_seed(3)
    x = torch.randn(n, d, generator=g)
    lq, lp, lw = torch.randn(n, generator=g), torch.randn(n, generator=g), torch.randn(n, generator=g)
    gq, gp = torch.randn(n, d, generator=g), torch.randn(n, d, generator=g)
    bad = torch.zeros(n, dtype=torch.bool)
    bad[torch.randperm(n, generator=g)[:137]] = True
    lq[bad & (torch.arange(n) % 3 == 0)] = float("nan")
    lp[bad & (torch.arange(n) % 3 == 1)] = float("inf")
    lp[bad & (torch.arange(n) % 3 == 2)] = -float("inf")
    keep = torch.isfinite(lq) & torch.isfinite(lp)
    pt = fb.Point(x.cuda(), lq.cuda(), lp.cuda(), gq.cuda(), gp.cuda())
    lwc = lw.cuda()
    ws = torch.empty(int(L.fab_filter_workspace_bytes(n, d)), dtype=torch.uint8, device="cuda")
    n_out = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.check(L.fab_nan_filter_f32(_lib.point_ptrs(pt), _lib.ptr(lwc), d, n, None, _lib.ptr(n_out),
                                    _lib.ptr(ws), _lib.stream_ptr()))
    m = int(n_out.item())
    assert m == int(keep.sum())
    assert torch.equal(pt.x[:m].cpu(), x[keep]) and torch.equal(pt.log_q[:m].cpu(), lq[keep])
    assert torch.equal(pt.log_p[:m].cpu(), lp[keep]) and torch.equal(lwc[:m].cpu(), lw[keep])
    assert torch.equal(pt.grad_log_q[:m].cpu(), gq[keep]) and torch.equal(pt.grad_log_p[:m].cpu(), gp[keep])
    # second pass on the compacted prefix (device-side count as input) is a no-op
    n_out2 = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.check(L.fab_nan_filter_f32(_lib.point_ptrs(pt), _lib.ptr(lwc), d, n, _lib.ptr(n_out),
                                    _lib.ptr(n_out2), _lib.ptr(ws), _lib.stream_ptr()))
    assert int(n_out2.item()) == m and torch.equal(pt.x[:m].cpu(), x[keep])


@pytest.mark.parametrize("p_target,alpha", [(True, None), (False, 2.0), (False, 0.5)])
def test_logw_update_is_bit_exact(p_target, alpha):
    """K11 (ais.py:93-100): same fp32 op sequence as torch -> identical bits."""
    L = _lib.lib()
    n = 5000
    lq, lp, lw = torch.randn(n) * 50 - 100, torch.randn(n) * 30 - 20, torch.randn(n) * 10
    B = beta_schedule("linear", 16)
    j = 7
    pt = OPoint(None, lq, lp)
    want = lw + (oracle_gamma(pt, B[j + 1], alpha, p_target) - oracle_gamma(pt, B[j], alpha, p_target))
    lwc, lqc, lpc = lw.cuda(), lq.cuda(), lp.cuda()     # keep the device tensors alive
    _lib.check(L.fab_logw_update_f32(fb.make_gamma(B[j], alpha, p_target),
                                     fb.make_gamma(B[j + 1], alpha, p_target),
                                     _lib.ptr(lqc), _lib.ptr(lpc), _lib.ptr(lwc), n,
                                     _lib.stream_ptr()))
    assert torch.equal(lwc.cpu(), want)


@pytest.mark.parametrize("n", [1, 7, 512, 2048, 16384])
@pytest.mark.parametrize("spread", [0.5, 5.0, 40.0])
def test_systematic_resample_bit_exact(n, spread):
    g = torch.Generator().manual_seed(n + int(spread * 10))
    lw = (torch.randn(n, generator=g) * spread + 123.0)
    if n > 10:
        lw[3] = float("nan")
        lw[5] = -float("inf")
    for u0 in (0, 1, 2 ** 31 + 12345, 2 ** 32 - 1):
        want = oracle_ancestors(lw.numpy(), u0)
        got = fb.systematic_ancestors(lw.cuda(), u0).cpu().numpy()
        assert np.array_equal(got, want), f"n={n} u0={u0}: {np.flatnonzero(got != want)[:5]}"
    # resampled particle counts follow the weights: E[count_i] = n * w_i, |count - n w| < 1
    q = fixed_point_weights(lw.numpy()).astype(np.float64)
    counts = np.bincount(want, minlength=n)
    assert np.all(np.abs(counts - n * q / q.sum()) < 1.0 + 1e-9)


def test_systematic_resample_point_and_properties():
    n, d = 4096, 32
    x = torch.randn(n, d).cuda()
    pt = fb.Point(x, torch.randn(n).cuda(), torch.randn(n).cuda(), torch.randn(n, d).cuda(),
                  torch.randn(n, d).cuda())
    lw = (torch.randn(n) * 3).cuda()
    new, anc = fb.systematic_resample(pt, lw, u0=987654321)
    assert torch.equal(new.x, x[anc]) and torch.equal(new.log_q, pt.log_q[anc])
    assert torch.equal(new.grad_log_p, pt.grad_log_p[anc])
    assert bool((anc[1:] >= anc[:-1]).all())                      # ancestors are sorted
    # uniform weights with any offset -> identity permutation (idempotence)
    anc_u = fb.systematic_ancestors(torch.zeros(n).cuda(), 12345)
    assert torch.equal(anc_u.cpu(), torch.arange(n))


def test_global_resample_and_ess_trigger_single_process():
    """resample.global_systematic_resample / resample_if_ess_below without a process group reduce
    to the single-device kernel (the multi-rank wiring is covered by tests/test_dist_gloo.py and
    tests/multi_gpu_check.py); weights are reset to the mean weight."""
    n, d = 2048, 32
    g = torch.Generator().manual_seed(5)
    lw = (torch.randn(n, generator=g) * 6 + 50).cuda()
    pt = fb.Point(torch.randn(n, d, generator=g).cuda(), torch.randn(n, generator=g).cuda(),
                  torch.randn(n, generator=g).cuda(), None, None)
    u0 = 2 ** 31 + 7
    want = oracle_ancestors(lw.cpu().numpy(), u0)
    new, anc, lw_new = fb.global_systematic_resample(pt, lw, u0)
    assert np.array_equal(anc.cpu().numpy(), want)
    assert torch.equal(new.x, pt.x[anc]) and new.grad_log_q is None
    mean_w = torch.logsumexp(lw.double(), 0).item() - np.log(n)
    assert torch.allclose(lw_new, torch.full_like(lw_new, mean_w), atol=1e-4)
    ess = float(fb.effective_sample_size(lw))
    keep = fb.resample_if_ess_below(pt, lw, ess, threshold=ess * 0.5, u0=u0)
    assert keep[2] is False and keep[0] is pt
    res = fb.resample_if_ess_below(pt, lw, ess, threshold=min(1.0, ess * 2), u0=u0)
    assert res[2] is True and torch.equal(res[0].x, new.x)
    assert float(fb.effective_sample_size(res[1])) > 0.999         # uniform weights afterwards


def _misaligned(t):
    """Same values at an address 4 bytes past a 16-byte boundary -> the scalar kernel forms."""
    buf = torch.empty(t.numel() + 4, dtype=t.dtype, device=t.device)
    v = buf[1:1 + t.numel()].view(t.shape)
    v.copy_(t)
    assert v.data_ptr() % 16 != 0
    return v


@pytest.mark.parametrize("dim", [32, 64, 128])
@pytest.mark.parametrize("n", [1, 5, 1023, 40000])
def test_streaming_target_matches_warp_per_row_form(dim, n):
    """k_target_manywell_v4 (16-byte loads, several rows per thread) sums each row in the order of
    manywell_row + warp_sum: identical bits to k_target, NaN/inf rows included."""
    L = _lib.lib()
    g = torch.Generator().manual_seed(dim + n)
    x = (torch.randn(n, dim, generator=g) * 1.9).cuda()
    if n > 4:
        x[2, 3] = float("nan")
        x[3, 0] = float("inf")
    tgt = fb.ManyWellEnergy(dim)
    desc = tgt.target_desc(x.device)
    out = []
    for xin in (x, _misaligned(x)):
        lp = torch.empty(n, device="cuda")
        gr = torch.empty(n, dim, device="cuda")
        _lib.check(L.fab_target_logprob_grad_f32(desc, _lib.ptr(xin), _lib.ptr(lp), _lib.ptr(gr), n,
                                                 _lib.stream_ptr()))
        lp2 = torch.empty(n, device="cuda")
        _lib.check(L.fab_target_logprob_grad_f32(desc, _lib.ptr(xin), _lib.ptr(lp2), None, n,
                                                 _lib.stream_ptr()))
        assert torch.equal(lp.view(torch.int32), lp2.view(torch.int32))        # value-only form
        out.append((lp, gr))
    assert torch.equal(out[0][0].view(torch.int32), out[1][0].view(torch.int32))
    assert torch.equal(out[0][1].view(torch.int32), out[1][1].view(torch.int32))


@pytest.mark.parametrize("n", [3, 4, 4099, 100001])
def test_logw_update_vector_form_matches_scalar(n):
    L = _lib.lib()
    g = torch.Generator().manual_seed(n)
    lq, lp, lw = (torch.randn(n, generator=g).cuda() * s for s in (50.0, 30.0, 10.0))
    ga, gb = fb.make_gamma(0.25, 2.0, False), fb.make_gamma(0.3125, 2.0, False)
    a, b = lw.clone(), _misaligned(lw)
    _lib.check(L.fab_logw_update_f32(ga, gb, _lib.ptr(lq), _lib.ptr(lp), _lib.ptr(a), n, _lib.stream_ptr()))
    _lib.check(L.fab_logw_update_f32(ga, gb, _lib.ptr(lq), _lib.ptr(lp), _lib.ptr(b), n, _lib.stream_ptr()))
    assert torch.equal(a, b) and not torch.equal(a, lw)


@pytest.mark.parametrize("rowf", [1, 4, 32, 60, 98, 128])
@pytest.mark.parametrize("n", [1, 777, 20011])
def test_gather_rows_all_forms(rowf, n):
    """fab_gather_rows_f32: 16-byte form (row_floats % 4 == 0) and scalar form == torch indexing."""
    L = _lib.lib()
    g = torch.Generator().manual_seed(rowf * 7 + n)
    m = max(n // 2, 1)
    src = torch.randn(m, rowf, generator=g).cuda()
    anc = torch.randint(0, m, (n,), generator=g).cuda()
    for s in (src, _misaligned(src)):
        dst = torch.empty(n, rowf, device="cuda")
        _lib.check(L.fab_gather_rows_f32(_lib.ptr(s), _lib.ptr(dst), _lib.ptr(anc), n, rowf,
                                         _lib.stream_ptr()))
        assert torch.equal(dst, src[anc])


def test_tuner_statistics_survive_a_smaller_batch():
    """the arrival counter of the cross-CTA reduction sits at a FIXED place in the
    operator workspace, so a call with a smaller batch (other grid size) after a larger one still
    produces statistics and tunes."""
    import fab_torch_b200 as fb
    from helpers import make_flows, make_manywell
    for engine in ("warp", "auto"):
        import os
        os.environ["FAB_ENGINE"] = engine
        try:
            _, _, fp = make_flows(32, 2, 10, last_std=0.02)
            _, tp = make_manywell(32)
            op = fb.HamiltonianMonteCarlo(4, 32, fp.log_prob, tp.log_prob, alpha=2.0, epsilon=0.2, L=2).cuda()
            for n in (2048, 512, 37, 1500):
                pt = op.create_new_point(torch.randn(n, 32, device="cuda"))
                eps_before, common_before = op.epsilons.clone(), op.common_epsilon.clone()
                op._stats.zero_()
                op.run(pt, 2, 0.5)
                torch.cuda.synchronize()
                assert float(op._stats[1]) == float(n), (engine, n, op._stats[:4].tolist())
                assert not torch.equal(op.epsilons, eps_before) and not torch.equal(op.common_epsilon, common_before)
        finally:
            del os.environ["FAB_ENGINE"]


def test_resample_without_any_finite_weight_raises():
    from fab_torch_b200.resample import systematic_ancestors
    lw = torch.full((64,), float("nan"), device="cuda")
    with pytest.raises(ValueError):
        systematic_ancestors(lw, 7)


def test_hmc_finish_over_peer_buffers_single_rank():
    """`fab_hmc_finish_peer_f32` (tuner statistics summed over peer-mapped exchange buffers) with a
    world of ONE: the rank stores its triple into its own buffer, reads it back and must apply exactly
    the update `fab_hmc_finish_f32` applies -- over more exchanges than the ring has slots (sequence
    counter, slot reuse).  The multi-rank path is exercised by `bench.py --gpus N` / tests/multi_gpu_check.py."""
    from fab_torch_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda", 0)

    def mk():
        _, _, fp = __import__("helpers").make_flows(4, 1, 2)
        tp = fb.ManyWellEnergy(4)
        return fb.HamiltonianMonteCarlo(3, 4, fp.log_prob, tp.log_prob, alpha=2.0, p_target=False,
                                        epsilon=0.2, L=2).cuda()
    op_a, op_b = mk(), mk()
    nbytes = int(L.fab_hmc_peer_buffer_bytes(1))
    assert nbytes == 4 * 1 * 8 * 4
    buf = torch.zeros(nbytes // 4, dtype=torch.float32, device=dev)
    ptrs = torch.tensor([buf.data_ptr()], dtype=torch.int64, device=dev)
    seq = torch.zeros(1, dtype=torch.int32, device=dev)
    g = _lib.Gamma(0.5, 0.5, 0.5, 1.0)
    gen = torch.Generator().manual_seed(4)
    for k in range(9):                      # > FAB_PEER_RING exchanges
        i = 1 + k % 3
        stats = torch.tensor([float(torch.rand((), generator=gen)) * 100, 128.0,
                              float(torch.rand((), generator=gen)) * 10, 0.0], device=dev)
        args = _lib.HmcArgs(i, 0, 2, 1, 0.65, 1000.0, g, 0, g, g, 1)
        _lib.check(L.fab_hmc_finish_f32(op_a._state(), args, _lib.ptr(stats), _lib.stream_ptr(dev)), "finish")
        _lib.check(L.fab_hmc_finish_peer_f32(op_b._state(), args, _lib.ptr(stats), _lib.ptr(ptrs), 1, 0,
                                             _lib.ptr(seq), _lib.stream_ptr(dev)), "finish_peer")
    torch.cuda.synchronize()
    assert int(seq.item()) == 9
    assert torch.equal(op_a.epsilons, op_b.epsilons) and torch.equal(op_a.common_epsilon, op_b.common_epsilon)
    assert torch.equal(op_a._log, op_b._log)
    assert not torch.equal(op_a.epsilons, mk().epsilons)          # the tuner did move
