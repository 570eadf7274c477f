"""Cycle breakdown of k_hmc_step_u (row-tile engine) at BASELINE config 2 from a -DUE_PROF build:
FAB_B200_LIB=build/libfab_prof.so python profiles/rowtile_prof.py [B].  Counters: umma_engine.cuh."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["FAB_ENGINE"] = "rowtile"
import bench
from fab_torch_b200 import _lib

NAMES = ["cw wait D0", "cw wait D1", "mma wait A0/A1", "mma wait stages", "mma issue", "prod wait slot",
         "cw barrier", "cw flow evals", "wait G1", "wait G2", "wait G3", "wait G3T", "wait G2T", "wait G1T"]


def read():
    buf = (C.c_ulonglong * 16)()
    _lib.lib().fab_umma_prof_read(buf)
    return list(buf)


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    cfg = dict(bench.CFG)
    flow, target, op, ais = bench.build_gpu(cfg, torch.device("cuda", 0), None)
    ais.use_cuda_graph = False
    ais.sample_and_log_weights(B)
    read()
    calls = 3
    for _ in range(calls):
        ais.sample_and_log_weights(B)
    v = read()
    launches = calls * cfg["M"] * cfg["n_outer"]
    print(f"B={B}: cycles of CTA 0 per k_hmc_step_u launch ({launches} launches)")
    for n, x in zip(NAMES, v):
        print(f"   {n:22s} {x / launches / 1e3:9.1f}k")
    print(f"   k_hmc_step_u {ais.time_transitions(B, repeats=2):.4f} ms/launch")
