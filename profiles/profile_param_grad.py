"""One FAB-loss step (loss = -mean(softmax(log_w) log q(x)), backward; fab/core.py:112-118) at the config-2
architecture for an ncu launch list: three warm-up steps (the parameter-space chain rule is captured into
a CUDA graph on the second), then `cudaProfilerStart` .. one step .. `cudaProfilerStop`.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/pg_launches.csv python profiles/profile_param_grad.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("FAB_ENGINE", "warp")
import fab_torch_b200 as fb

if __name__ == "__main__":
    dim, K, npd, n = 32, 10, 10, 2048
    torch.manual_seed(0)
    flow = fb.B200RealNVP(dim, K, npd).cuda()
    with torch.no_grad():
        for blk in flow._blocks():
            lin = blk.linears[2]
            lin.weight.normal_(0, 0.02); lin.bias.normal_(0, 0.02)
    x = torch.randn(n, dim, device="cuda")
    w = torch.softmax(torch.randn(n, device="cuda"), 0)

    def step():
        flow.zero_grad(set_to_none=True)
        (-(w * flow.log_prob(x)).mean()).backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
