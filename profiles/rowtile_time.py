"""k_hmc_step_u time per launch at BASELINE config 2 for the library in FAB_B200_LIB (A/B of builds)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["FAB_ENGINE"] = "rowtile"
import bench

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    flow, target, op, ais = bench.build_gpu(dict(bench.CFG), torch.device("cuda", 0), None)
    ais.use_cuda_graph = False
    torch.manual_seed(0)
    pt, lw = ais.sample_and_log_weights(B)
    info = ais.get_logging_info()
    print(f"{os.environ.get('FAB_B200_LIB', 'default'):32s} B={B}: k_hmc_step_u {ais.time_transitions(B, repeats=3):.4f} ms/launch  "
          f"log_Z {info['log_Z']:.5f}")
