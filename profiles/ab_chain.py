"""A/B timing helper: ms per BASELINE-config-2 chain (sample_and_log_weights(2048), CUDA graph)
for the build named by FAB_B200_LIB (default: the in-tree library).  Experiment helper only.

    for v in build/var/*.so; do FAB_B200_LIB=$v python profiles/ab_chain.py; done
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch          # noqa: E402
import bench          # noqa: E402

device = torch.device("cuda", 0)
flow, target, op, ais = bench.build_gpu(bench.CFG, device, None)
torch.manual_seed(1234)
B = bench.CFG["batch_per_gpu"]
for _ in range(3):
    ais.sample_and_log_weights(B)
torch.cuda.synchronize()
ts = []
for _ in range(int(os.environ.get("AB_STEPS", "12"))):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    ais.sample_and_log_weights(B)
    e.record()
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
ts.sort()
info = ais.get_logging_info()
print(f"{os.environ.get('FAB_B200_LIB', 'in-tree')}: median {ts[len(ts) // 2]:.3f} ms  min {ts[0]:.3f} ms "
      f"per chain; log_Z {info['log_Z']:.6g}", flush=True)
