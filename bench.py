#!/usr/bin/env python
"""Benchmark of the AIS hot path: BASELINE.json metric "AIS particles/sec (Many-Well-32, 16 dists,
HMC L=5)".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A *step* is one `AnnealedImportanceSampler.sample_and_log_weights(batch)` call (ais.py:53-87) on
BASELINE config 2: Many-Well d=32, RealNVP 10 layers x width 320, 16 intermediate distributions,
HMC L=5 / n_outer=1 (tuner on), 2048 particles per GPU (weak scaling: N GPUs carry 2048*N
particles sharded over ranks; the only exchanges are the scalar ESS all-gather and the tuner's
all-reduce).  Prints ONE JSON line on rank 0 (contract in the task statement); see DESIGN.md §6.

Initial HMC step size eps0 = 0.1 (the reference's own many_well.yaml value is 1.0): with eps0 = 1
and a randomly initialised flow every proposal is rejected for the first ~10 calls and, once the
tuner has found an accepting step size, chains escape along directions where q decays faster than
p^2 and log Z reaches 1e37 (profiles/r02_workload_sanity.log; the reference does the same).  With
eps0 = 0.1 the chain accepts from the first call and log Z stays finite over the whole run; the
work per step (flow passes, launches) is identical.

`--impl reference` times the UNMODIFIED reference (AnnealedImportanceSampler, HamiltonianMonteCarlo,
ManyWellEnergy from baseline/_ref, installed by __graft_entry__.build()) on the host cores; the
flow object handed to it is oracle.realnvp.OracleRealNVP, because the reference's flow lives in the
third-party package `normflows`, which is not installable here (flow parity: unpinned).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG = dict(dim=32, n_layers=10, nodes_per_dim=10, M=16, L=5, n_outer=1, epsilon=0.1,
           batch_per_gpu=2048, alpha=2.0, p_target=False)
METRIC = "ais_particles_per_sec"
UNIT = "particles/s"
WORKLOAD = "manywell32_realnvp10x320_M16_hmcL5_b2048_per_gpu"


def config_block(cfg, world):
    """Identical in both arms (the driver compares them)."""
    return dict(workload=WORKLOAD, global_batch=cfg["batch_per_gpu"] * world, eps0=cfg["epsilon"],
                tuner="on", l2_flush_between_steps=True, flow_parity="unpinned",
                **{k: cfg[k] for k in ("dim", "n_layers", "M", "L", "n_outer")})


def flops_per_particle_flow_pass(dim, K, W):
    d1 = int(dim / 2 + 0.5)
    d2 = dim - d1
    return 2 * K * (d1 * W + W * W + 2 * W * d2 + dim * dim)        # F_f


def algorithmic_flops_per_particle(cfg):
    Ff = flops_per_particle_flow_pass(cfg["dim"], cfg["n_layers"], cfg["dim"] * cfg["nodes_per_dim"])
    return Ff * (1 + 2 * (1 + cfg["M"] * cfg["L"] * cfg["n_outer"]))   # 163 * F_f at config 2


def profiled_traffic():
    """dram__bytes_read+write per k_hmc_step launch from the latest committed ncu capture."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_metrics.json")))
    if not files:
        return None
    return json.load(open(files[-1])).get("dram_bytes_per_launch")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"],
                    bf16_tflops_sustained=p["bf16_tflops_sustained"],
                    sm_max_mhz=p.get("sm_max_mhz", 1965.0), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0,
                sm_max_mhz=1965.0, source="fallback")


# ------------------------------------------------------------------------------ CPU reference arm
def build_cpu_port(cfg, batch):
    """The reference path on the CPU: oracle port of AIS + HMC + Many-Well with the restated flow,
    same architecture/initialisation as the GPU arm."""
    from oracle.realnvp import OracleRealNVP, randomize_last_layers
    from oracle.sampler import OracleAIS, OracleHMC
    from oracle.targets import OracleManyWell
    torch.manual_seed(0)
    flow = OracleRealNVP(cfg["dim"], cfg["n_layers"], cfg["nodes_per_dim"])
    randomize_last_layers(flow, 0.01, seed=1)
    target = OracleManyWell(cfg["dim"])
    op = OracleHMC(cfg["M"], cfg["dim"], flow.log_prob, target.log_prob, alpha=cfg["alpha"],
                   p_target=cfg["p_target"], epsilon=cfg["epsilon"], n_outer=cfg["n_outer"],
                   L=cfg["L"])
    ais = OracleAIS(flow, target.log_prob, op, p_target=cfg["p_target"], alpha=cfg["alpha"],
                    n_intermediate_distributions=cfg["M"])
    return ais


def build_cpu_reference(cfg, dtype=torch.float32):
    """The reference's own classes (baseline/_ref, unmodified) around the restated flow.  Returns
    (sampler, kind): kind = "reference" or, if baseline/_ref is absent, "port" (oracle/sampler.py,
    pinned bit-for-bit against the reference: tests/golden/pin_report.json)."""
    from oracle.ref_loader import installed_reference_available, load_reference
    if not installed_reference_available():
        return build_cpu_port(cfg, cfg["batch_per_gpu"]), "port"
    load_reference(installed=True)
    from fab.sampling_methods import AnnealedImportanceSampler, HamiltonianMonteCarlo
    from fab.target_distributions.many_well import ManyWellEnergy
    from oracle.realnvp import OracleRealNVP, randomize_last_layers
    torch.manual_seed(0)
    flow = OracleRealNVP(cfg["dim"], cfg["n_layers"], cfg["nodes_per_dim"])
    randomize_last_layers(flow, 0.01, seed=1)
    target = ManyWellEnergy(cfg["dim"], a=-0.5, b=-6.0, use_gpu=False)
    if dtype == torch.float64:
        flow, target = flow.double(), target.double()
    op = HamiltonianMonteCarlo(cfg["M"], cfg["dim"], flow.log_prob, target.log_prob, alpha=cfg["alpha"],
                               p_target=cfg["p_target"], epsilon=cfg["epsilon"], n_outer=cfg["n_outer"],
                               L=cfg["L"])
    if dtype == torch.float64:
        op = op.double()
    ais = AnnealedImportanceSampler(flow, target.log_prob, op, p_target=cfg["p_target"],
                                    alpha=cfg["alpha"], n_intermediate_distributions=cfg["M"])
    return ais, "reference"


def time_cpu(cfg, batch, steps, warmup, dtype=torch.float32):
    """Wall-clock of full `sample_and_log_weights(batch)` calls of the CPU arm on all host threads."""
    torch.set_num_threads(os.cpu_count() or 1)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)          # the reference follows the default dtype (many_well.yaml:41)
    try:
        ais, kind = build_cpu_reference(cfg, dtype)
        torch.manual_seed(1234)
        for _ in range(warmup):
            ais.sample_and_log_weights(batch)
        ts = []
        for _ in range(steps):
            t0 = time.perf_counter()
            ais.sample_and_log_weights(batch)
            ts.append(time.perf_counter() - t0)
        return ts, ais.get_logging_info(), kind
    finally:
        torch.set_default_dtype(prev)


def parity_leg(cfg, batch, device):
    """log-Z / log-weight agreement of the CUDA chain with the CPU arm on IDENTICAL noise: one
    fresh-state `sample_and_log_weights(batch)` on each side (same weights; the CPU side -- the
    oracle port, pinned bit-for-bit against the unmodified reference -- records every variate it
    draws and the CUDA side replays them).  BASELINE target: |dlogZ| <= 1e-3.  The chain is chaotic
    (80 leapfrog steps), so the per-particle errors are reported as a distribution plus the number
    of rows above the 1e-5 bar; with an untrained flow a handful of particles carry log Z."""
    import copy
    import fab_torch_b200 as fb
    from oracle.noise import Float32RecordingNoise
    ais_c = build_cpu_port(cfg, batch)
    noise = Float32RecordingNoise()
    ais_c.transition_operator.noise = noise
    flow_c = ais_c.base_distribution
    torch.manual_seed(4321)
    flow_c._eps_override = noise.base_eps(batch, cfg["dim"], torch.float32, "cpu")
    pt_c, lw_c = ais_c.sample_and_log_weights(batch)
    info_c = ais_c.get_logging_info()
    flow_g = fb.B200RealNVP(cfg["dim"], cfg["n_layers"], cfg["nodes_per_dim"])
    flow_g.load_state_dict(flow_c.state_dict())
    flow_g = flow_g.to(device)
    target = fb.ManyWellEnergy(cfg["dim"])
    op = fb.HamiltonianMonteCarlo(cfg["M"], cfg["dim"], flow_g.log_prob, target.log_prob,
                                  alpha=cfg["alpha"], p_target=cfg["p_target"], epsilon=cfg["epsilon"],
                                  n_outer=cfg["n_outer"], L=cfg["L"]).to(device)
    ais_g = fb.AnnealedImportanceSampler(flow_g, target.log_prob, op, p_target=cfg["p_target"],
                                         alpha=cfg["alpha"], n_intermediate_distributions=cfg["M"])
    op.noise = fb.InjectedNoise(copy.deepcopy(noise.record))
    pt_g, lw_g = ais_g.sample_and_log_weights(batch)
    info_g = ais_g.get_logging_info()
    out = dict(log_Z_cuda=info_g["log_Z"], log_Z_cpu_port=info_c["log_Z"],
               log_Z_abs_err=abs(info_g["log_Z"] - info_c["log_Z"]),
               ess_ais_cuda=info_g["ess_ais"], ess_ais_cpu_port=info_c["ess_ais"],
               n_cuda=int(lw_g.shape[0]), n_cpu_port=int(lw_c.shape[0]),
               note="one fresh-state call each on identical injected noise (fp32 both sides)")
    if lw_g.shape[0] == lw_c.shape[0]:
        a, b = lw_g.detach().cpu().double(), lw_c.detach().double()
        e = ((a - b).abs() / b.abs().clamp_min(1.0))
        dx = (pt_g.x.detach().cpu().double() - pt_c.x.double()).abs().max(dim=1).values
        div = dx > 1e-2 * (1 + pt_c.x.double().abs().max(dim=1).values)
        out.update(log_w_rel_err_median=float(e.median()), log_w_rel_err_p99=float(e.quantile(0.99)),
                   log_w_rel_err_max_same_branch=float(e[~div].max()) if (~div).any() else None,
                   rows_above_1e5_bar=int((e[~div] > 1e-5).sum()), rows_compared=int((~div).sum()),
                   chains_on_other_accept_branch=int(div.sum()))
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = CFG["batch_per_gpu"]
    ts, info, kind = time_cpu(CFG, batch, args.steps, args.warmup)
    total = float(np.sum(ts))
    value = batch * len(ts) / total
    cores = torch.get_num_threads()
    sample = (f"each step = one full sample_and_log_weights({batch}) call of the workload through the "
              f"{'unmodified reference classes (baseline/_ref)' if kind == 'reference' else 'oracle port'} "
              f"(fp32, torch CPU, {cores} threads, ONE process); particles/s is per-particle, so the same "
              f"figure is reported for the {args.gpus}-GPU global batch")
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * total / len(ts), higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=config_block(CFG, args.gpus), impl="reference",
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind=kind, sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0, log_Z_last_timed_call=info["log_Z"], ess_ais_last_timed_call=info["ess_ais"])
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except Exception:
                continue
            for name, cell in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                   "sw_power_cap"), r[4:8]):
                if cell.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(np.max(mx)),
                    power_w_max=float(np.max(pw)), samples=len(sm), reasons=sorted(reasons))


# ------------------------------------------------------------------------------ GPU arm
def build_gpu(cfg, device, group):
    import fab_torch_b200 as fb
    torch.manual_seed(0)                     # identical weights on every rank
    flow = fb.B200RealNVP(cfg["dim"], cfg["n_layers"], cfg["nodes_per_dim"])
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():                    # non-zero last layers, N(0, 0.01^2)
        for k in range(cfg["n_layers"]):
            lin = flow._nf_model.flows[2 * k].linears[2]
            lin.weight.copy_(torch.randn(lin.weight.shape, generator=g) * 0.01)
            lin.bias.copy_(torch.randn(lin.bias.shape, generator=g) * 0.01)
    flow = flow.to(device)
    target = fb.ManyWellEnergy(cfg["dim"])
    op = fb.HamiltonianMonteCarlo(cfg["M"], cfg["dim"], flow.log_prob, target.log_prob,
                                  alpha=cfg["alpha"], p_target=cfg["p_target"],
                                  epsilon=cfg["epsilon"], n_outer=cfg["n_outer"], L=cfg["L"]).to(device)
    ais = fb.AnnealedImportanceSampler(flow, target.log_prob, op, p_target=cfg["p_target"],
                                       alpha=cfg["alpha"], n_intermediate_distributions=cfg["M"],
                                       process_group=group, use_cuda_graph=True)
    return flow, target, op, ais


def run_gpu_arm(args):
    import torch.distributed as dist
    import fab_torch_b200 as fb
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        group = dist.group.WORLD
    cfg = CFG
    B_local = cfg["batch_per_gpu"]
    B_global = B_local * world
    flow, target, op, ais = build_gpu(cfg, device, group)
    torch.manual_seed(1234 + rank)           # independent particles per rank
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=device)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident metric -------------------------------------------------------------
    for _ in range(args.warmup):
        ais.sample_and_log_weights(B_global)
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    barrier()
    for s, e in ev:
        flush.fill_(1.0)                     # evict L2 between timed iterations (untimed)
        s.record()
        ais.sample_and_log_weights(B_global)
        e.record()
    barrier()
    ms = torch.tensor([sum(s.elapsed_time(e) for s, e in ev)], dtype=torch.float64, device=device)
    info = ais.get_logging_info()
    # ---- end-to-end: host noise -> H2D -> chain -> D2H of the result ------------------------
    M, d, no = cfg["M"], cfg["dim"], cfg["n_outer"]
    h_eps = torch.randn(B_local, d).pin_memory()
    h_mom = torch.randn(M, no, B_local, d).pin_memory()
    h_exp = torch.empty(M, no, B_local).exponential_(1.0).pin_memory()
    h_x = torch.empty(B_local, d).pin_memory()
    h_w = torch.empty(B_local).pin_memory()
    h2d = (h_eps.numel() + h_mom.numel() + h_exp.numel()) * 4
    d2h = (h_x.numel() + h_w.numel()) * 4 + 40

    def e2e_step():
        ais.set_next_noise(h_eps, h_mom, h_exp)          # pinned host -> device inside the call
        pt, lw = ais.sample_and_log_weights(B_global)
        n = lw.shape[0]
        h_x[:n].copy_(pt.x, non_blocking=True)
        h_w[:n].copy_(lw, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(h_w[:n].max())

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    for s, e in ev2:
        flush.fill_(1.0)
        s.record()
        e2e_step()
        e.record()
    barrier()
    ms_e2e = torch.tensor([sum(s.elapsed_time(e) for s, e in ev2)], dtype=torch.float64,
                          device=device)
    clock_rec = clocks.stop() if rank == 0 else None
    # ---- tuner off (set_eval_mode(True)): same chain without the step-size update ----
    op.set_eval_mode(True)
    for _ in range(2):
        ais.sample_and_log_weights(B_global)
    barrier()
    ev3 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for s_, e_ in ev3:
        flush.fill_(1.0)
        s_.record()
        ais.sample_and_log_weights(B_global)
        e_.record()
    barrier()
    ms_eval = torch.tensor([sum(a.elapsed_time(b) for a, b in ev3) / len(ev3)], dtype=torch.float64, device=device)
    op.set_eval_mode(False)
    # ---- roofline of the dominant kernel (the fused HMC step), timed per launch on its stream ----
    op.chain_noise_override = None
    k_ms = ais.time_transitions(B_global, repeats=2)
    k_ms_t = torch.tensor([k_ms], dtype=torch.float64, device=device)
    # ---- N > 1: global ESS trigger + global systematic resample, checked against one device -------
    multi = None
    if world > 1:
        multi = multi_gpu_parity(ais, B_global, device, group, rank, world)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
        dist.all_reduce(k_ms_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ms_eval, op=dist.ReduceOp.MAX)
    total_ms, total_e2e_ms, kernel_ms = float(ms), float(ms_e2e), float(k_ms_t)
    if rank == 0:
        peaks = measured_peaks()
        value = B_global * args.steps / (total_ms * 1e-3)
        e2e_value = B_global * args.steps / (total_e2e_ms * 1e-3)
        Ff = flops_per_particle_flow_pass(d, cfg["n_layers"], d * cfg["nodes_per_dim"])
        flops_per_launch = 2.0 * Ff * cfg["L"] * B_local          # value+grad per leapfrog
        achieved = flops_per_launch / (kernel_ms * 1e-3) / 1e12
        sm_mhz = (clock_rec or {}).get("sm_max_mhz") or peaks["sm_max_mhz"]
        n_sm = torch.cuda.get_device_properties(device).multi_processor_count
        ffma_peak = n_sm * 128 * 2 * sm_mhz * 1e6 / 1e12
        rowtile = flow.use_rowtile(B_local)
        if rowtile:
            # tcgen05 kind::f16, cta_group::2: 8192 FLOP/cycle/SM measured (profiles/r02_mb_umma_rates.log);
            # the fp32-grade split spends three f16 MMAs per algorithmic one, and a CTA pair carries 128
            # particles, so B_local / 64 SMs can be busy (sequential chain per particle)
            busy_sm = min(n_sm, 2 * ((B_local + 127) // 128))
            pipe_peak = busy_sm * 8192 * sm_mhz * 1e6 / 1e12 / 3
            kernel, pipe = "k_hmc_step_u", (f"tcgen05.mma kind::f16 cta_group::2, f16 hi/lo split x3 (fp32-grade), "
                                            f"{busy_sm} of {n_sm} SMs carry row tiles at this batch")
        else:
            pipe_peak = n_sm * 4 * 2048 / 8 * sm_mhz * 1e6 / 1e12 / 3
            kernel, pipe = "k_hmc_step", "mma.sync m16n8k8 tf32 x3 (fp32-grade split), 14 particles per SM"
        roof = dict(bound="tensor", achieved=achieved, peak=peaks["bf16_tflops_sustained"],
                    unit="TFLOP/s", frac=achieved / peaks["bf16_tflops_sustained"], traffic=profiled_traffic(),
                    kernel=kernel, kernel_ms=kernel_ms, peak_source=peaks["source"] +
                    " bf16_tflops_sustained (kernel timed inside a long step)",
                    pipe=pipe, pipe_peak=pipe_peak, pipe_frac=achieved / pipe_peak, fp32_ffma_peak=ffma_peak,
                    flops_per_launch=flops_per_launch,
                    note="algorithmic FLOPs = 2 F_f L B (SURVEY 8d); frac is vs the measured bf16 cuBLAS peak as "
                         "required; pipe_frac is vs the ceiling of the engine and of the SMs this batch can "
                         "occupy (DESIGN.md 4: the chain is sequential per particle and 2048 particles are 16 "
                         "row tiles of 128)")
        # chain init (1 fused launch on the warp engine; sample + row-tile log q / gradient + target + tail on the
        # row-tile engine), filter 3, ESS 2, M transitions (+ the tuner exchange launch at N > 1), filter 3, ESS 2
        launches_per_step = (4 if rowtile else 1) + 3 + 2 + cfg["M"] * cfg["n_outer"] * (2 if world > 1 else 1) + 3 + 2
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=total_ms / args.steps, higher_is_better=True,
                    scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                    config=config_block(cfg, world),
                    engine=dict(hmc_step="row-tile tcgen05 (f16 hi/lo operands = 22 significant bits of every fp32 weight "
                                         "and activation, fp32 accumulate)" if rowtile
                                else "warp-level mma.sync 3xTF32 (weights rounded to 22 bits)",
                                cuda_graph=bool(ais.use_cuda_graph),
                                tuner_exchange=("none (one rank)" if world == 1 else
                                                "NVLink peer memory (fab_hmc_finish_peer_f32)" if getattr(op, "_peer", None) is not None
                                                else "NCCL all-reduce")),
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d * world,
                             d2h_bytes_per_step=d2h * world, ms_per_step=total_e2e_ms / args.steps),
                    gpu_launches=launches_per_step * args.steps, clocks=clock_rec, roofline=roof,
                    eval_mode=dict(ms_per_step=float(ms_eval), value=B_global / (float(ms_eval) * 1e-3),
                                   note="set_eval_mode(True): tuner off, 5 timed steps"),
                    log_Z_last_timed_call=info["log_Z"], ess_ais_last_timed_call=info["ess_ais"])
        if multi is not None:
            line["multi_gpu_parity"] = multi
        if world == 1:
            # BASELINE config 3's batch (16 384 particles) on this ONE GPU: the batch at which every SM carries
            # row tiles (the headline batch of 2 048 per GPU fills 32 of 148) -- same step, same kernels, extra leg
            # outside the timed region of the headline
            try:
                B_big = 8 * B_local
                _, _, _, ais_big = build_gpu(cfg, device, None)
                for _ in range(2):
                    ais_big.sample_and_log_weights(B_big)
                ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
                for a_, b_ in ev:
                    a_.record(); ais_big.sample_and_log_weights(B_big); b_.record()
                torch.cuda.synchronize()
                ms_big = sum(a_.elapsed_time(b_) for a_, b_ in ev) / len(ev)
                k_big = ais_big.time_transitions(B_big, repeats=1)
                tf_big = 2.0 * Ff * cfg["L"] * B_big / (k_big * 1e-3) / 1e12
                line["full_chip_batch"] = dict(
                    particles=B_big, ms_per_step=ms_big, value=B_big / (ms_big * 1e-3), unit=UNIT,
                    kernel_ms=k_big, achieved_tflops=tf_big, frac=tf_big / peaks["bf16_tflops_sustained"],
                    pipe_frac=tf_big / (n_sm * 8192 * sm_mhz * 1e6 / 1e12 / 3),
                    note="config 3's 16 384 particles on one GPU (strong-scaling reference point of the 8-GPU line); "
                         "frac vs the measured bf16 cuBLAS peak (three f16 products per fp32-grade product, so the "
                         "ceiling of this arithmetic is a third of it), pipe_frac vs the tcgen05 rate of all SMs "
                         "(8192 FLOP/cycle/SM measured) / 3")
                del ais_big
            except Exception as e:           # noqa: BLE001 -- an extra leg must never cost the headline line
                line["full_chip_batch"] = dict(error=f"{type(e).__name__}: {e}")
        if world == 1 and not args.no_cpu_baseline:
            ts, cinfo, kind = time_cpu(cfg, B_local, steps=7, warmup=1)      # ~10 s of CPU work
            cv = B_local * len(ts) / float(np.sum(ts))
            ts64, _, _ = time_cpu(cfg, B_local, steps=1, warmup=0, dtype=torch.float64)
            line["cpu_baseline"] = dict(
                value=cv, unit=UNIT, cores=torch.get_num_threads(), kind=kind,
                sample=f"{len(ts)} timed + 1 warm-up full sample_and_log_weights({B_local}) calls through the "
                       f"{'unmodified reference classes (baseline/_ref)' if kind == 'reference' else 'oracle port'}, "
                       f"fp32, {float(np.sum(ts)):.1f} s of CPU work",
                fp64_value=B_local / float(ts64[0]),
                fp64_note="the reference's default dtype (many_well.yaml:41): 1 timed call, no warm-up")
            line["parity"] = parity_leg(cfg, B_local, device)
        print(json.dumps(line), flush=True)
    if world > 1:
        # captured graphs hold NCCL kernels: release them before the communicator goes away
        ais.release_graphs()
        torch.cuda.synchronize()
        dist.barrier()
        # (belt and braces: a teardown that does not return must not hold the GPU box)
        import threading
        guard = threading.Timer(30.0, lambda: os._exit(0))
        guard.daemon = True
        guard.start()
        dist.destroy_process_group()
        guard.cancel()


def multi_gpu_parity(ais, B_global, device, group, rank, world):
    """Untimed check on N > 1 GPUs (BASELINE config 3): the rank-sharded chain with the all-reduced
    ESS and the global systematic resample must equal the single-device answer bit for bit.  Every
    rank holds the same pre-drawn noise for the GLOBAL batch; rank 0 also runs the whole batch alone."""
    import torch.distributed as dist
    import fab_torch_b200 as fb
    from fab_torch_b200.resample import systematic_resample
    cfg = CFG
    M, d, no = cfg["M"], cfg["dim"], cfg["n_outer"]
    g = torch.Generator().manual_seed(99)
    eps = torch.randn(B_global, d, generator=g)
    mom = torch.randn(M, no, B_global, d, generator=g)
    exp = torch.empty(M, no, B_global).exponential_(1.0, generator=g)
    B_local = B_global // world
    sl = slice(rank * B_local, (rank + 1) * B_local)
    op = ais.transition_operator
    op.set_eval_mode(True)                      # identical tuner state on both sides
    ais.set_next_noise(eps[sl].contiguous(), mom[:, :, sl].contiguous(), exp[:, :, sl].contiguous())
    pt, lw = ais.sample_and_log_weights(B_global)
    info = ais.get_logging_info()
    pt_r, lw_r, did = ais.resample_if_ess_below(pt, lw, threshold=1.1, u0=12345)   # always triggers
    out = dict(ess_ais_global=info["ess_ais"], log_Z_global=info["log_Z"], resampled=bool(did))
    # gather the sharded result on rank 0 and compare with a single-device run of the global batch
    sizes = torch.tensor([lw.shape[0]], device=device)
    all_sizes = [torch.empty_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    if any(int(t.item()) != B_local for t in all_sizes):       # NaN filter shrank a shard: not expected here
        out["skipped"] = "ragged shards after the NaN filter"
        op.set_eval_mode(False)
        return out
    xs = [torch.empty_like(pt_r.x) for _ in range(world)]
    ws = [torch.empty_like(lw) for _ in range(world)]
    dist.all_gather(xs, pt_r.x.contiguous(), group=group)
    dist.all_gather(ws, lw.contiguous(), group=group)
    if rank == 0:
        flow1, target1, op1, ais1 = build_gpu(cfg, device, None)
        ais1.use_cuda_graph = False
        op1.set_eval_mode(True)
        op1.epsilons.copy_(op.epsilons)                 # same (tuned) step sizes on both sides
        op1.common_epsilon.copy_(op.common_epsilon)
        ais1.set_next_noise(eps, mom, exp)
        pt1, lw1 = ais1.sample_and_log_weights(B_global)
        info1 = ais1.get_logging_info()
        pt1_r, lw1_r, _ = ais1.resample_if_ess_below(pt1, lw1, threshold=1.1, u0=12345)
        out.update(log_w_bit_equal=bool(torch.equal(torch.cat(ws), lw1)),
                   resampled_x_bit_equal=bool(torch.equal(torch.cat(xs), pt1_r.x)),
                   ess_abs_diff=abs(info["ess_ais"] - info1["ess_ais"]),
                   log_Z_abs_diff=abs(info["log_Z"] - info1["log_Z"]))
    op.set_eval_mode(False)
    dist.barrier()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
