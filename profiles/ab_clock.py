"""Is the slower chain after an L2 flush a clock effect or a memory-system effect?  FAB_PROF build:
cycles (clock64) of the slowest CTA of the last k_hmc_step launch vs the event time of the chain."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import fab_torch_b200 as fb

device = torch.device("cuda", 0)
flow, target, op, ais = bench.build_gpu(bench.CFG, device, None)
torch.manual_seed(1234)
B = bench.CFG["batch_per_gpu"]
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=device)
big = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=device)
lib = fb._lib.lib()
cyc = (ctypes.c_ulonglong * 1024)(); smid = (ctypes.c_uint * 1024)()
lib.fab_debug_cta_cycles.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.POINTER(ctypes.c_uint)]
for _ in range(3):
    ais.sample_and_log_weights(B)
torch.cuda.synchronize()

def run(mode, steps=8):
    out = []
    for _ in range(steps):
        if mode == "flush_write":
            flush.fill_(1.0)
        elif mode == "flush_read":
            big.sum()
        elif mode == "flush_write_then_read":
            flush.fill_(1.0); big.sum()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ais.sample_and_log_weights(B); e.record()
        torch.cuda.synchronize()
        assert lib.fab_debug_cta_cycles(cyc, smid) == 0
        mx = max(cyc[i] for i in range(147))
        out.append((s.elapsed_time(e), mx))
    print(f"{mode:24s}", " ".join(f"{t:.2f}ms/{c/1e6:.3f}Mcyc" for t, c in out), flush=True)

for mode in os.environ.get("AB_MODES", "none,flush_write,none,flush_read,none").split(","):
    run(mode)
