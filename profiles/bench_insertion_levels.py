"""`sample_and_log_weights(2048)` at BASELINE config 2 through the three ways of INTEGRATION.md:
  level 1: the UNMODIFIED reference sampler (baseline/_ref: fab/sampling_methods/ais.py, its Python loop over
           the intermediate distributions, its host-side filters and ESS) calling the B200 flow / target /
           fused HMC `transition` as plugins;
  level 2: fab_torch_b200.AnnealedImportanceSampler (one C-ABI chain call), eager;
  level 2 + CUDA graph (what bench.py times).
Wall clock per call (synchronised), median of 10 after 3 warm-ups; tuner on."""
import os
import statistics
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                    # noqa: E402
import fab_torch_b200 as fb                                      # noqa: E402
from oracle.ref_loader import installed_reference_available, load_reference   # noqa: E402


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(ts)


if __name__ == "__main__":
    cfg = dict(bench.CFG)
    B = 2048
    dev = torch.device("cuda", 0)
    out = {}
    flow, target, op, ais = bench.build_gpu(cfg, dev, None)
    ais.use_cuda_graph = False
    out["level 2 (B200 sampler, eager)"] = timed(lambda: ais.sample_and_log_weights(B))
    ais.use_cuda_graph = True
    out["level 2 (B200 sampler, CUDA graph)"] = timed(lambda: ais.sample_and_log_weights(B))
    if installed_reference_available():
        load_reference(installed=True)
        from fab.sampling_methods.ais import AnnealedImportanceSampler as RefAIS
        flow1, target1, op1, _ = bench.build_gpu(cfg, dev, None)
        ref = RefAIS(base_distribution=flow1, target_log_prob=target1.log_prob, transition_operator=op1,
                     p_target=cfg["p_target"], alpha=cfg["alpha"], n_intermediate_distributions=cfg["M"])
        out["level 1 (reference's Python AIS loop + B200 plugins)"] = timed(lambda: ref.sample_and_log_weights(B))
    for k, v in out.items():
        print(f"{k:58s} {v:8.3f} ms per call   {B / v:8.1f} k particles/s")
