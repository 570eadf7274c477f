"""Per-launch time of the fused HMC step on both engines for small batches (where should
FAB_ENGINE=auto switch?): config-2 architecture, B = 64 .. 2048."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

if __name__ == "__main__":
    for B in (64, 128, 256, 512, 1024, 2048):
        row = [f"B={B:5d}"]
        for eng in ("warp", "rowtile"):
            os.environ["FAB_ENGINE"] = eng
            flow, target, op, ais = bench.build_gpu(dict(bench.CFG), torch.device("cuda", 0), None)
            ais.use_cuda_graph = False
            ais.sample_and_log_weights(B)
            row.append(f"{eng} {ais.time_transitions(B, repeats=2):.4f} ms")
        print("   ".join(row), flush=True)
