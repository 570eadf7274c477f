"""Achieved HBM bandwidth of the elementwise / copy kernels of the path (the un-fused forms the
operator-level plugin API uses; inside `sample_and_log_weights` the same work is fused into the
tile kernels), at sizes that exceed L2, against MEASURED_PEAKS.json `hbm_gbs`.

    python profiles/bench_hbm_kernels.py          # one JSON object per kernel
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                        # noqa: E402
import fab_torch_b200 as fb         # noqa: E402
from fab_torch_b200 import _lib     # noqa: E402

dev = torch.device("cuda", 0)
L = _lib.lib()
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, warmup=3, steps=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for s, e in ev:
        s.record(); fn(); e.record()
    torch.cuda.synchronize()
    return min(s.elapsed_time(e) for s, e in ev)


def report(name, nbytes, ms, note=""):
    gbs = nbytes / ms / 1e6
    print(json.dumps(dict(kernel=name, algorithmic_bytes=nbytes, ms=round(ms, 4), achieved_gbs=round(gbs, 1),
                          peak_gbs=peak, frac=round(gbs / peak, 3), note=note)), flush=True)


n, d = 1 << 22, 32
g = fb.make_gamma(0.3, 2.0, False)
gn = fb.make_gamma(0.4, 2.0, False)
s = _lib.stream_ptr(dev)
nw = 1 << 26                        # 3 x 256 MB: larger than the 126 MB L2
lq, lp, lw = (torch.randn(nw, device=dev) for _ in range(3))
report("k_logw_update_v4 (ais.py:93-100)", nw * 16,
       timed(lambda: L.fab_logw_update_f32(g, gn, _lib.ptr(lq), _lib.ptr(lp), _lib.ptr(lw), nw, s)),
       "reads log_q, log_p, log_w; writes log_w; n = 2^26")
del lq, lp, lw
lq, lp, lw = (torch.randn(n, device=dev) for _ in range(3))

x = torch.randn(n, d, device=dev)
tgt = fb.ManyWellEnergy(d)
out_lp = torch.empty(n, device=dev)
out_g = torch.empty(n, d, device=dev)
desc = tgt.target_desc(dev)
report("k_target_manywell_v4 value+grad (many_well.py:81-90)", n * (2 * d + 1) * 4,
       timed(lambda: L.fab_target_logprob_grad_f32(desc, _lib.ptr(x), _lib.ptr(out_lp), _lib.ptr(out_g), n, s)),
       "reads x[n,32]; writes log_p[n], grad[n,32]")

anc = torch.randint(0, n, (n,), device=dev, dtype=torch.int64).sort().values
dst = torch.empty_like(x)
report("k_gather_rows_v4 (resample gather, sorted ancestors)", n * (2 * d * 4 + 8),
       timed(lambda: L.fab_gather_rows_f32(_lib.ptr(x), _lib.ptr(dst), _lib.ptr(anc), n, d, s)),
       "reads anc[n] + x rows; writes rows")

bx = torch.empty(n, d, device=dev)
blw, blq = torch.empty(n, device=dev), torch.empty(n, device=dev)
report("k_buffer_add_v4 (prioritised_replay_buffer.py:71-85)", n * (2 * d + 4) * 4,
       timed(lambda: L.fab_buffer_add_f32(_lib.ptr(bx), _lib.ptr(blw), _lib.ptr(blq), n, d, 12345,
                                          _lib.ptr(x), _lib.ptr(lq), _lib.ptr(lp), n, s)),
       "reads batch rows; writes ring rows (wrapping)")

gum = torch.randn(n, device=dev)
ws = torch.empty(int(L.fab_buffer_topk_workspace_bytes(n)), dtype=torch.uint8, device=dev)
idx = torch.empty(4096, dtype=torch.int64, device=dev)
report("k_buffer_keys + k_buffer_select_* (Gumbel-top-k, k=4096 of 2^22)", n * 8 + n * 4 + 5 * n * 4,
       timed(lambda: L.fab_buffer_topk_f32(_lib.ptr(lw), _lib.ptr(gum), n, 4096, _lib.ptr(idx), _lib.ptr(ws), s)),
       "6 launches at full-chip width: keys+digit 0, digits 1-3, count, ordered scatter (48 MB of keys: L2 resident)")
