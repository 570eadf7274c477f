"""Debug helper (not a test): are the large isolated gradient errors ReLU-kink flips?
For the worst particle of the flow test, perturb x by ~1e-6 in fp64 and see whether the fp64
gradient itself jumps by a comparable amount."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from helpers import make_flows

dim, K, npd, n = 32, 10, 10, 512
fo64, fo, fp = make_flows(dim, K, npd, last_std=0.05)
g = torch.Generator().manual_seed(5)
x = torch.randn(n, dim, generator=g) * 1.5


def grad64(xx):
    xx = xx.double().requires_grad_(True)
    lq = fo64.log_prob(xx)
    return lq.detach(), torch.autograd.grad(lq.sum(), xx)[0]


lq_ref, g_ref = grad64(x)
x32 = x.clone().requires_grad_(True)
lq32 = fo.log_prob(x32)
g32 = torch.autograd.grad(lq32.sum(), x32)[0]
lq, grad = fp.cuda_log_prob(x.cuda(), with_grad=True)
err = ((grad.cpu().double() - g_ref).abs() / g_ref.abs().clamp_min(1)).max(dim=1).values
err32 = ((g32.double() - g_ref).abs() / g_ref.abs().clamp_min(1)).max(dim=1).values
top = torch.argsort(err, descending=True)[:5]
print("worst particles (cuda):", [(int(i), f"{err[i]:.2e}", f"cpu32 {err32[i]:.2e}") for i in top])
print("worst particles (cpu32):", [(int(i), f"{err32[i]:.2e}") for i in torch.argsort(err32, descending=True)[:5]])
print("median err cuda %.2e cpu32 %.2e" % (err.median(), err32.median()))
for i in top[:3]:
    xi = x[i:i + 1]
    jumps = []
    for t in range(20):
        d = torch.randn(1, dim, generator=g).double() * 1e-6
        _, gi = grad64(xi.double() + d)
        jumps.append(((gi - g_ref[i:i + 1]).abs() / g_ref[i:i + 1].abs().clamp_min(1)).max().item())
    print(f"particle {int(i)}: cuda err {err[i]:.2e}; fp64 gradient change under 1e-6 perturbations: "
          f"max {max(jumps):.2e} median {sorted(jumps)[10]:.2e};  log_q err cuda "
          f"{abs(lq[i].item() - lq_ref[i].item()):.2e}  |log_q| {abs(lq_ref[i].item()):.1f}  |g|max {g_ref[i].abs().max():.1f}")
