"""Whole AIS chain (BASELINE config 2) on the row-tile engine vs the warp-level engine: same
weights, same injected noise; prints log-weight agreement and per-launch k_hmc_step times."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import fab_torch_b200 as fb


def run(engine, B, noise, eps0=None, graph=False):
    os.environ["FAB_ENGINE"] = engine
    cfg = dict(bench.CFG)
    if eps0 is not None:
        cfg["epsilon"] = eps0
    flow, target, op, ais = bench.build_gpu(cfg, torch.device("cuda", 0), None)
    ais.use_cuda_graph = graph
    h_eps, h_mom, h_exp = noise
    ais.set_next_noise(h_eps, h_mom, h_exp)
    pt, lw = ais.sample_and_log_weights(B)
    torch.cuda.synchronize()
    info = ais.get_logging_info()
    k_ms = ais.time_transitions(B, repeats=2)
    return pt, lw, info, k_ms, op


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    eps0 = float(sys.argv[2]) if len(sys.argv) > 2 else None
    cfg = bench.CFG
    g = torch.Generator().manual_seed(7)
    noise = (torch.randn(B, cfg["dim"], generator=g).pin_memory(),
             torch.randn(cfg["M"], cfg["n_outer"], B, cfg["dim"], generator=g).pin_memory(),
             torch.empty(cfg["M"], cfg["n_outer"], B).exponential_(1.0, generator=g).pin_memory())
    res = {}
    for eng in ("warp", "rowtile"):
        pt, lw, info, k_ms, op = run(eng, B, noise, eps0)
        res[eng] = (pt, lw, info, op)
        print(f"{eng:8s} B={B}: k_hmc_step {k_ms:.4f} ms/launch   log_Z {info['log_Z']:.6f}  ess {info['ess_ais']:.5f}  "
              f"n={lw.shape[0]}  eps {op.epsilons[:3, 0].tolist()} common {op.common_epsilon.item():.5f}")
    (pa, la, ia, oa), (pb, lb, ib, ob) = res["warp"], res["rowtile"]
    if la.shape == lb.shape:
        e = ((la - lb).abs() / la.abs().clamp_min(1.0)).double()
        dx = (pa.x - pb.x).abs().max(dim=1).values
        div = dx > 1e-2 * (1 + pa.x.abs().max(dim=1).values)
        print(f"log_w rel diff rowtile vs warp: median {e.median().item():.3e} p99 {e.quantile(0.99).item():.3e} "
              f"max(same branch) {e[~div].max().item():.3e}; chains on another accept branch: {int(div.sum())}")
        print(f"tuner state equal: {torch.equal(oa.epsilons, ob.epsilons)} {torch.equal(oa.common_epsilon, ob.common_epsilon)}")
