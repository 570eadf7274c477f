import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch
from helpers import make_flows, rel_err

def run(dim, K, npd, last_std, n=64):
    fo64, fo, fp = make_flows(dim, K, npd, last_std=last_std)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(n, dim, generator=g) * 1.5
    x64 = x.double().requires_grad_(True)
    lq_ref = fo64.log_prob(x64); g_ref = torch.autograd.grad(lq_ref.sum(), x64)[0]
    x32 = x.clone().requires_grad_(True)
    lq_32 = fo.log_prob(x32); g_32 = torch.autograd.grad(lq_32.sum(), x32)[0]
    lq, grad = fp.cuda_log_prob(x.cuda(), with_grad=True)
    eps = torch.randn(n, dim, generator=g)
    xs_ref, lqs_ref = fo64._nf_model.sample(n, eps=eps.double())
    xs_32, lqs_32 = fo._nf_model.sample(n, eps=eps)
    xs, lqs = fp.cuda_sample(eps.cuda())
    print(f"d={dim} K={K} W={dim*npd} std={last_std}: log_q cuda {rel_err(lq, lq_ref):.2e} cpu32 {rel_err(lq_32, lq_ref):.2e} | "
          f"grad cuda {rel_err(grad, g_ref):.2e} cpu32 {rel_err(g_32, g_ref):.2e} | "
          f"sample x cuda {rel_err(xs, xs_ref):.2e} cpu32 {rel_err(xs_32, xs_ref):.2e} | lq_s cuda {rel_err(lqs, lqs_ref):.2e} cpu32 {rel_err(lqs_32, lqs_ref):.2e}")

for cfg in [(32, 1, 10, 0.0), (32, 1, 10, 0.05), (32, 2, 10, 0.05), (32, 4, 10, 0.05), (32, 10, 10, 0.0), (32, 10, 10, 0.01), (32, 10, 10, 0.05), (32, 10, 1, 0.05), (8, 10, 4, 0.05)]:
    run(*cfg)
