"""Why does a chain inside bench.py's timed loop take longer than an isolated one?  Times the
config-2 chain (a) isolated with a sync after each, (b) back to back, (c) back to back with the
L2 flush between chains (bench.py's loop), and prints per-step times + SM clock samples."""
import os, sys, subprocess, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

device = torch.device("cuda", 0)
flow, target, op, ais = bench.build_gpu(bench.CFG, device, None)
torch.manual_seed(1234)
B = bench.CFG["batch_per_gpu"]
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=device)
for _ in range(3):
    ais.sample_and_log_weights(B)
torch.cuda.synchronize()

def run(mode, steps=20):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    clk = []
    stop = False
    def pump():
        while not stop:
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu", "--format=csv,noheader,nounits", "-i", "0"],
                                 capture_output=True, text=True).stdout.strip()
            clk.append(out)
            time.sleep(0.05)
    th = threading.Thread(target=pump); th.start()
    torch.cuda.synchronize()
    for s, e in ev:
        if mode == "flush":
            flush.fill_(1.0)
        s.record()
        ais.sample_and_log_weights(B)
        e.record()
        if mode == "isolated":
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    stop = True; th.join()
    ts = [s.elapsed_time(e) for s, e in ev]
    print(mode, "ms/step:", " ".join(f"{t:.2f}" for t in ts), flush=True)
    print("   clocks(sm MHz, W, C):", " | ".join(clk[:12]), flush=True)

for mode in ("isolated", "back2back", "flush", "isolated"):
    run(mode)
