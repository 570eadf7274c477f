"""Per-phase cycle profile of k_hmc_step (CTA 0), experiment helper.

  FAB_NVCC_FLAGS="-DFAB_PROF" python fab_torch_b200/csrc/build.py --force
  python profiles/phase_profile.py            # prints cycles per phase for one config-2 chain
  python fab_torch_b200/csrc/build.py --force # restore the product build

The counters only exist in -DFAB_PROF builds (common.cuh: prof_mark); thread 0 of CTA 0 charges
the cycles since the previous mark to a phase id after the barrier that ends the phase.
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch          # noqa: E402
import bench          # noqa: E402
import fab_torch_b200 as fb   # noqa: E402

NAMES = {0: "hmc prologue (loads, momentum)", 1: "leapfrog elementwise", 2: "rows->operand",
         3: "inv g1 z->[v|h1] gemm+epilogue", 5: "inv g2 gemm+epilogue (WxW)",
         7: "inv g3 k-split gemm (W->2*d2)", 8: "inv coupling + logdet", 9: "base gaussian",
         10: "bwd coupling prep", 11: "bwd g3t gemm+epilogue (2*d2->W)",
         13: "bwd g2t gemm+epilogue (WxW)", 15: "bwd g1mt k-split gemm (W+d->d)",
         16: "bwd g1mt reduce", 17: "operand->rows + target", 18: "accept / store"}

device = torch.device("cuda", 0)
flow, target, op, ais = bench.build_gpu(bench.CFG, device, None)
torch.manual_seed(1234)
B = bench.CFG["batch_per_gpu"]
ais.sample_and_log_weights(B)
torch.cuda.synchronize()
lib = fb._lib.lib()
buf = (ctypes.c_ulonglong * 32)()
lib.fab_debug_prof.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
assert lib.fab_debug_prof(buf, 1) == 0
ais.sample_and_log_weights(B)
torch.cuda.synchronize()
assert lib.fab_debug_prof(buf, 0) == 0
tot = sum(buf)
launches = bench.CFG["M"] * bench.CFG["n_outer"]
print(f"CTA 0, {launches} k_hmc_step launches (+ k_ais_init's eval): {tot} cycles "
      f"= {tot / 1.965e6:.2f} ms at 1965 MHz")
for i in sorted(NAMES):
    print(f"{i:2d} {NAMES[i]:34s} {buf[i]:12d} cyc  {100.0 * buf[i] / tot:5.1f} %")

# per-CTA duration of the last k_hmc_step launch (SM-to-SM spread)
cyc = (ctypes.c_ulonglong * 1024)()
smid = (ctypes.c_uint * 1024)()
lib.fab_debug_cta_cycles.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.POINTER(ctypes.c_uint)]
assert lib.fab_debug_cta_cycles(cyc, smid) == 0
n_cta = (B + fb._lib.lib().fab_tile_particles(flow.desc(), B) - 1) // fb._lib.lib().fab_tile_particles(flow.desc(), B)
v = sorted((cyc[i], smid[i], i) for i in range(n_cta))
med = v[len(v) // 2][0]
print(f"per-CTA cycles of the last launch ({n_cta} CTAs): min {v[0][0]}  median {med}  max {v[-1][0]} "
      f"(max/median {v[-1][0] / med:.3f})")
print("slowest (cycles, smid, cta):", v[-6:])
print("fastest (cycles, smid, cta):", v[:4])
