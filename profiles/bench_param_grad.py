"""Timed FAB-loss step (fab/core.py:112-118: loss = -mean(softmax(log_w) log q(x)); backward) at
the config-2 architecture: the CUDA tape + weight-gradient kernels vs the torch-op re-evaluation
(autograd through B200RealNVP.torch_log_prob) this build used before."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("FAB_ENGINE", "warp")
import fab_torch_b200 as fb


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


if __name__ == "__main__":
    for dim, K, npd, n in ((32, 10, 10, 2048), (32, 10, 10, 512), (60, 10, 5, 1024)):
        torch.manual_seed(0)
        flow = fb.B200RealNVP(dim, K, npd).cuda()
        with torch.no_grad():
            for k in range(K):
                lin = flow._blocks()[k].linears[2]
                lin.weight.normal_(0, 0.02); lin.bias.normal_(0, 0.02)
        x = torch.randn(n, dim, device="cuda")
        w = torch.softmax(torch.randn(n, device="cuda"), 0)

        def step_cuda():
            flow.zero_grad(set_to_none=True)
            (-(w * flow.log_prob(x)).mean()).backward()

        def step_torch():
            flow.zero_grad(set_to_none=True)
            (-(w * flow.torch_log_prob(x)).mean()).backward()

        # device time of the two kernel calls alone, and the per-optimiser-step repack of the weight blobs
        lq, _, tape = flow.cuda_log_prob_tape(x, with_grad=False)
        t_tape = timed(lambda: flow.cuda_log_prob_tape(x, with_grad=False))
        t_pg = timed(lambda: flow.cuda_param_grad(tape, w))

        def repack():
            with torch.no_grad():
                flow._nf_model.q0.loc.add_(0.0)          # version bump = "an optimiser step happened"
            flow.blob()
            if flow.rowtile_supported():
                flow.umma_blob()
        t_pack = timed(repack)
        print(f"   kernels only: forward + tape {t_tape:.3f} ms, weight-gradient GEMMs + host chain rule {t_pg:.3f} ms; "
              f"repack of the weight blob(s) after a parameter update {t_pack:.3f} ms")
        a, b = timed(step_cuda), timed(step_torch)
        step_cuda(); g1 = [p.grad.clone() for p in flow.parameters()]
        step_torch(); g2 = [p.grad.clone() for p in flow.parameters()]
        err = max(((u - v).abs().max() / v.abs().max().clamp_min(1e-12)).item() for u, v in zip(g1, g2))
        print(f"dim {dim} K {K} W {dim * npd} n {n}: loss+backward  cuda kernels {a:.3f} ms   torch ops {b:.3f} ms   "
              f"speed-up {b / a:.2f}x   max grad diff (rel. to tensor max) {err:.2e}")
