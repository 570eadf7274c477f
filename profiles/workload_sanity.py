"""How does the config-2 bench workload behave over 25 consecutive calls?  (log_Z ran
away to 5e37 with eps0 = 1, tuner on.)  Prints log_Z / ESS / p_accept per call for several
(eps0, last-layer std, tuner) settings."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import fab_torch_b200 as fb

os.environ.setdefault("FAB_ENGINE", "warp")


def run(eps0, tune, reset, calls=25, B=2048):
    cfg = dict(bench.CFG)
    cfg["epsilon"] = eps0
    flow, target, op, ais = bench.build_gpu(cfg, torch.device("cuda", 0), None)
    op.set_eval_mode(not tune)
    eps_init, common_init = op.epsilons.clone(), op.common_epsilon.clone()
    torch.manual_seed(1234)
    out = []
    for c in range(calls):
        if reset:
            op.epsilons.copy_(eps_init); op.common_epsilon.copy_(common_init)
        pt, lw = ais.sample_and_log_weights(B)
        info = ais.get_logging_info()
        out.append((info["log_Z"], info["ess_ais"], info.get("dist0_p_accept_0", float("nan")),
                    info.get(f"dist{cfg['M'] - 1}_p_accept_0", float("nan")), lw.shape[0]))
    print(f"eps0={eps0} tune={tune} reset={reset}: final eps[0]={op.get_epsilon(1, 0).item():.4f}")
    for c in (0, 1, 2, 4, 9, 14, 19, 24):
        if c < len(out):
            z, e, p0, pM, n = out[c]
            print(f"   call {c:2d}: log_Z {z:12.4f}  ess {e:.5f} (x N = {e * n:7.1f})  p_acc first/last {p0:.3f}/{pM:.3f}  n={n}")


if __name__ == "__main__":
    run(1.0, True, False)
    run(1.0, True, True)
    run(0.2, True, False)
    run(0.1, True, False)
    run(0.1, False, False)
    run(0.05, True, False)
