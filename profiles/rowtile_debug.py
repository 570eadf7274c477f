"""Bring-up check of the row-tile (tcgen05) engine: flow log-density + input-gradient against the
fp64 torch-op restatement on the GPU and against the warp-level engine.  Run under gpurun."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import fab_torch_b200 as fb
from fab_torch_b200 import _lib


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs() / b.abs().clamp_min(1.0)).max().item()


def prof(tag, calls=1):
    import ctypes as C
    L = _lib.lib()
    if not hasattr(L, "fab_umma_prof_read"):
        return
    buf = (C.c_ulonglong * 16)()
    L.fab_umma_prof_read(buf)
    names = ["cw wait acc", "-", "mma wait aready", "mma wait stages", "mma issue", "prod wait slot", "cw barrier", "cw total"]
    print(f"   [prof {tag}] " + ", ".join(f"{n}={buf[i] / calls / 1e3:.1f}k" for i, n in enumerate(names) if n != "-"))


def stats(name, got, ref):
    got, ref = got.double(), ref.double()
    ok = torch.isfinite(ref) & torch.isfinite(got)
    if ok.ndim == 2:
        ok = ok.all(dim=1, keepdim=True).expand_as(got)
    got, ref = got[ok], ref[ok]
    e = (got - ref).abs() / ref.abs().clamp_min(1.0)
    print(f"   {name:10s} max {e.max().item():.3e}  p99 {e.flatten().quantile(0.99).item():.3e}  "
          f"median {e.flatten().median().item():.3e}  finite {got.numel()}")


def case(dim, K, npd, n, seed=0, last_std=0.05, xs=1.5):
    torch.manual_seed(seed)
    flow = fb.B200RealNVP(dim, K, npd)
    with torch.no_grad():
        for k in range(K):
            lin = flow._nf_model.flows[2 * k].linears[2]
            lin.weight.normal_(0, last_std)
            lin.bias.normal_(0, last_std)
        flow._nf_model.q0.loc.normal_(0, 0.2)
        flow._nf_model.q0.log_scale.normal_(0, 0.1)
    flow = flow.cuda()
    f64 = fb.B200RealNVP(dim, K, npd)
    f64.load_state_dict(flow.state_dict())
    f64 = f64.cuda().double()
    x = torch.randn(n, dim, device="cuda") * xs
    xd = x.double().requires_grad_(True)
    lq64 = f64.torch_log_prob(xd)
    g64 = torch.autograd.grad(lq64.sum(), xd)[0]
    os.environ["FAB_ENGINE"] = "warp"
    lq_w, g_w = flow.cuda_log_prob(x, with_grad=True)
    os.environ["FAB_ENGINE"] = "rowtile"
    lq_r, g_r = flow.cuda_log_prob(x, with_grad=True)
    lq_r0, _ = flow.cuda_log_prob(x, with_grad=False)
    torch.cuda.synchronize()
    print(f"dim={dim} K={K} W={dim * npd} n={n}")
    stats("warp lq", lq_w, lq64.detach()); stats("rowtile lq", lq_r, lq64.detach())
    stats("warp grad", g_w, g64); stats("rowtile g", g_r, g64)
    fin = torch.isfinite(lq_r) & torch.isfinite(lq_w)
    lq_w, lq_r, lq_r0, lq64 = lq_w[fin], lq_r[fin], lq_r0[fin], lq64[fin]
    prof("(discard)")
    print(f"   value-only == value+grad: {torch.equal(lq_r, lq_r0)}   mean signed lq err: warp "
          f"{(lq_w.double() - lq64.detach()).mean().item():+.3e}  rowtile {(lq_r.double() - lq64.detach()).mean().item():+.3e}")
    # timing
    for eng in ("warp", "rowtile"):
        os.environ["FAB_ENGINE"] = eng
        for _ in range(3):
            flow.cuda_log_prob(x, with_grad=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            flow.cuda_log_prob(x, with_grad=True)
        torch.cuda.synchronize()
        print(f"   {eng:8s} log_prob+grad: {(time.perf_counter() - t0) * 100:.3f} ms per call")
        if eng == "rowtile":
            prof("rowtile, per call", 13)


def calib(dim=32, K=10, npd=10, n=4096):
    """Sweep the truncation compensation constant (FAB_UE_TRUNC) and print bias / error of log q."""
    torch.manual_seed(0)
    flow = fb.B200RealNVP(dim, K, npd)
    with torch.no_grad():
        for k in range(K):
            lin = flow._nf_model.flows[2 * k].linears[2]
            lin.weight.normal_(0, 0.02); lin.bias.normal_(0, 0.02)
    flow = flow.cuda()
    f64 = fb.B200RealNVP(dim, K, npd)
    f64.load_state_dict(flow.state_dict())
    f64 = f64.cuda().double()
    x = torch.randn(n, dim, device="cuda")
    xd = x.double().requires_grad_(True)
    lq64 = f64.torch_log_prob(xd)
    g64 = torch.autograd.grad(lq64.sum(), xd)[0]
    lq64 = lq64.detach()
    os.environ["FAB_ENGINE"] = "warp"
    lq_w, g_w = flow.cuda_log_prob(x, with_grad=True)
    print(f"calibration, K={K} W={dim * npd} n={n}; |log q| median {lq64.abs().median().item():.1f}")
    print(f"   warp        : lq mean signed {(lq_w.double() - lq64).mean().item():+.3e} rms {(lq_w.double() - lq64).pow(2).mean().sqrt().item():.3e}; "
          f"grad rel rms {((g_w.double() - g64).norm() / g64.norm()).item():.3e}")
    os.environ["FAB_ENGINE"] = "rowtile"
    for tr in ((0.0, 0.0), (1.67e-8, 1.67e-8), (3e-8, 1.67e-8), (4.3e-8, 1.67e-8), (6e-8, 1.67e-8), (4.3e-8, 1.0e-8), (4.3e-8, 2.4e-8)):
        os.environ["FAB_UE_TRUNC"] = f"{tr[0]:.3e},{tr[1]:.3e}"
        lq_r, g_r = flow.cuda_log_prob(x, with_grad=True)
        print(f"   trunc {tr[0]:.2e},{tr[1]:.2e}: lq mean signed {(lq_r.double() - lq64).mean().item():+.3e} rms {(lq_r.double() - lq64).pow(2).mean().sqrt().item():.3e}; "
              f"grad rel rms {((g_r.double() - g64).norm() / g64.norm()).item():.3e}  "
              f"grad signed {(((g_r.double() - g64) * g64.sign()).sum() / g64.abs().sum()).item():+.3e}")
    del os.environ["FAB_UE_TRUNC"]
    flow._ublob_key = None


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    if which == "calib":
        calib()
    elif which == "small":
        case(32, 1, 2, 64)
        case(32, 1, 10, 128)
        case(32, 2, 10, 200)
    else:
        case(32, 10, 10, 2048, last_std=0.02, xs=1.0)
        case(32, 10, 10, 16384, last_std=0.02, xs=1.0)
