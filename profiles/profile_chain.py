"""Workload for ncu: one warm-up + `--chains` BASELINE-config-2 AIS calls (the bench.py step).

  launch list : ncu --metrics gpu__time_duration.sum --clock-control none --csv \
                    --log-file gpurun_out/launches.csv python profiles/profile_chain.py
  top kernel  : ncu --set full --clock-control none --import-source on -k regex:k_hmc_step \
                    -s 20 -c 1 -o gpurun_out/hmc_step python profiles/profile_chain.py
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch          # noqa: E402
import bench          # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=1)
ap.add_argument("--batch", type=int, default=bench.CFG["batch_per_gpu"])
args = ap.parse_args()
device = torch.device("cuda", 0)
flow, target, op, ais = bench.build_gpu(bench.CFG, device, None)
torch.manual_seed(1234)
ais.sample_and_log_weights(args.batch)          # warm-up (also packs the weight blob)
torch.cuda.synchronize()
for _ in range(args.chains):
    ais.sample_and_log_weights(args.batch)
torch.cuda.synchronize()
print(ais.get_logging_info())
