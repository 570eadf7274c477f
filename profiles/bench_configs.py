"""Device timings of the other BASELINE.json configs (parity-test cases, not bench.py lines):

  C1  GMM-40 d=2, RealNVP 4x80, 8 distributions, Metropolis x1, batch 512       (+ CPU port)
  C3s Many-Well-32 as config 3 (batch 16384) on ONE GPU: the strong-scaling reference point for the
      8-GPU line of bench.py (16384 = 8 x 2048) and the batch where every SM carries row tiles
  C4  Many-Well-128, RealNVP 10x1280, 32 distributions, HMC L=10, the 512-particle shard one of
      8 GPUs carries for batch 4096
  C5  ALDP surrogate d=60, RealNVP 10x300, 20 distributions, HMC L=4, batch 1024, feeding the
      device prioritised replay buffer (add, Gumbel-top-k sample, adjust)

    python profiles/bench_configs.py            # prints one JSON object per config
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                      # noqa: E402
import fab_torch_b200 as fb       # noqa: E402

dev = torch.device("cuda", 0)


def randomize(flow, std, seed=1):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for k in range(flow.n_flow_layers):
            lin = flow._nf_model.flows[2 * k].linears[2]
            lin.weight.copy_(torch.randn(lin.weight.shape, generator=g) * std)
            lin.bias.copy_(torch.randn(lin.bias.shape, generator=g) * std)


def timed(fn, warmup, steps):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for s, e in ev:
        s.record(); fn(); e.record()
    torch.cuda.synchronize()
    return sum(s.elapsed_time(e) for s, e in ev) / steps


def c1():
    torch.manual_seed(0)
    target = fb.GMM(2, 40, 40.0, 1.0)
    torch.manual_seed(0)
    flow = fb.B200RealNVP(2, 4, 40); randomize(flow, 0.01); flow = flow.to(dev)
    op = fb.Metropolis(8, 2, flow.log_prob, target.log_prob, n_updates=1, alpha=2.0, p_target=False,
                       max_step_size=5.0, min_step_size=5.0, adjust_step_size=False).to(dev)
    ais = fb.AnnealedImportanceSampler(flow, target.log_prob, op, p_target=False, alpha=2.0,
                                       n_intermediate_distributions=8, use_cuda_graph=True)
    ms = timed(lambda: ais.sample_and_log_weights(512), 5, 50)
    out = dict(config="C1 gmm40_d2_realnvp4x80_M8_metropolis_b512", ms_per_call=ms, particles_per_s=512 / ms * 1e3)
    try:
        from oracle.realnvp import OracleRealNVP, randomize_last_layers
        from oracle.sampler import OracleAIS, OracleMetropolis
        from oracle.targets import OracleGMM
        torch.manual_seed(0)
        to = OracleGMM(2, 40, 40.0, 1.0)
        torch.manual_seed(0)
        fo = OracleRealNVP(2, 4, 40); randomize_last_layers(fo, 0.01, seed=1)
        oo = OracleMetropolis(8, 2, fo.log_prob, to.log_prob, n_updates=1, alpha=2.0, p_target=False,
                              max_step_size=5.0, min_step_size=5.0, adjust_step_size=False)
        ao = OracleAIS(fo, to.log_prob, oo, p_target=False, alpha=2.0, n_intermediate_distributions=8)
        torch.set_num_threads(os.cpu_count() or 1)
        for _ in range(3):
            ao.sample_and_log_weights(512)
        t0 = time.perf_counter()
        for _ in range(20):
            ao.sample_and_log_weights(512)
        cpu_ms = (time.perf_counter() - t0) / 20 * 1e3
        out.update(cpu_port_ms_per_call=cpu_ms, cpu_port_particles_per_s=512 / cpu_ms * 1e3,
                   cpu_threads=torch.get_num_threads())
    except Exception as e:        # the oracle is optional here
        out["cpu_port"] = f"unavailable: {e}"
    return out


def c3_strong():
    import bench
    out = []
    for B in (2048, 16384):
        flow, target, op, ais = bench.build_gpu(dict(bench.CFG), dev, None)
        ms = timed(lambda: ais.sample_and_log_weights(B), 3, 10)
        k_ms = ais.time_transitions(B, repeats=2)
        Ff = bench.flops_per_particle_flow_pass(32, 10, 320)
        out.append(dict(config=f"C3s manywell32_realnvp10x320_M16_hmcL5 batch {B} on ONE GPU", ms_per_call=ms,
                        particles_per_s=B / ms * 1e3, k_hmc_step_ms=k_ms,
                        k_hmc_step_algorithmic_tflops=2 * Ff * 5 * B / k_ms / 1e9,
                        engine="rowtile" if flow.use_rowtile(B) else "warp"))
    return out


def c4():
    torch.manual_seed(0)
    flow = fb.B200RealNVP(128, 10, 10); randomize(flow, 0.003); flow = flow.to(dev)
    target = fb.ManyWellEnergy(128)
    op = fb.HamiltonianMonteCarlo(32, 128, flow.log_prob, target.log_prob, alpha=2.0, p_target=False,
                                  epsilon=0.05, n_outer=1, L=10).to(dev)
    ais = fb.AnnealedImportanceSampler(flow, target.log_prob, op, p_target=False, alpha=2.0,
                                       n_intermediate_distributions=32)
    ms = timed(lambda: ais.sample_and_log_weights(512), 1, 2)
    Ff = 2 * 10 * (64 * 1280 + 1280 * 1280 + 2 * 1280 * 64 + 128 * 128)
    flops = Ff * (1 + 2 * (1 + 32 * 10)) * 512
    return dict(config="C4 manywell128_realnvp10x1280_M32_hmcL10, 512-particle shard of batch 4096 / 8 GPUs",
                ms_per_call=ms, particles_per_s=512 / ms * 1e3, algorithmic_tflops=flops / ms / 1e9,
                tile_particles=int(fb._lib.lib().fab_tile_particles(flow.desc(), 512)),
                blob_mb=flow.desc().total_floats * 4 / 1e6)


def c5():
    torch.manual_seed(0)
    dim, M, B = 60, 20, 1024
    flow = fb.B200RealNVP(dim, 10, 5); randomize(flow, 0.01); flow = flow.to(dev)
    target = fb.AldpSurrogateEnergy(dim)
    op = fb.HamiltonianMonteCarlo(M, dim, flow.log_prob, target.log_prob, alpha=2.0, p_target=False,
                                  epsilon=0.1, n_outer=1, L=4).to(dev)
    ais = fb.AnnealedImportanceSampler(flow, target.log_prob, op, p_target=False, alpha=2.0,
                                       n_intermediate_distributions=M, use_cuda_graph=True)

    def sampler():
        pt, lw = ais.sample_and_log_weights(B, logging=False)
        return pt.x, lw, pt.log_q
    buf = fb.PrioritisedReplayBuffer(dim, 64 * B, 8 * B, sampler, device="cuda")
    ms_chain = timed(lambda: ais.sample_and_log_weights(B), 2, 10)
    data = sampler()
    ms_add = timed(lambda: buf.add(*data), 3, 20)
    ms_sample = timed(lambda: buf.sample_n_batches(B, 8), 3, 20)
    x, lw, lq, idx = buf.sample(B)
    adj = torch.randn(B, device=dev) * 0.1
    ms_adjust = timed(lambda: buf.adjust(adj, lq, idx), 3, 20)
    return dict(config="C5 aldp_surrogate60_realnvp10x300_M20_hmcL4_b1024 + prioritised buffer (8192 live rows)",
                ms_per_chain=ms_chain, particles_per_s=B / ms_chain * 1e3, ms_buffer_add_1024=ms_add,
                ms_buffer_sample_8x1024_of_8192=ms_sample, ms_buffer_adjust_1024=ms_adjust)


if __name__ == "__main__":
    for fn in (c1, c3_strong, c4, c5):
        r = fn()
        for item in (r if isinstance(r, list) else [r]):
            print(json.dumps(item), flush=True)
