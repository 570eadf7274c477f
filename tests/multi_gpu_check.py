"""Multi-GPU check (run with torchrun, one rank per GPU; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py

The particle-sharded run (NCCL all-gather of ESS quadruples, all-reduce of tuner statistics) must
reproduce the single-GPU run on the same global batch with the same noise: per-particle log-weights
and states identical, ESS / log Z / tuner state equal to rounding of the reductions.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fab_torch_b200 as fb      # noqa: E402


def build(device, group):
    torch.manual_seed(0)
    flow = fb.B200RealNVP(32, 10, 10)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for k in range(10):
            lin = flow._nf_model.flows[2 * k].linears[2]
            lin.weight.copy_(torch.randn(lin.weight.shape, generator=g) * 0.01)
            lin.bias.copy_(torch.randn(lin.bias.shape, generator=g) * 0.01)
    flow = flow.to(device)
    target = fb.ManyWellEnergy(32)
    M = 6
    op = fb.HamiltonianMonteCarlo(M, 32, flow.log_prob, target.log_prob, alpha=2.0, p_target=False,
                                  epsilon=0.05, n_outer=2, L=3).to(device)
    ais = fb.AnnealedImportanceSampler(flow, target.log_prob, op, p_target=False, alpha=2.0,
                                       n_intermediate_distributions=M, process_group=group)
    return flow, op, ais, M


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    device = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=device)
    all_ok = True
    # the two engines round differently, so the comparison is made per engine (FAB_ENGINE=auto picks
    # by the LOCAL batch size, which differs between the sharded and the single-device run)
    for engine in ("warp", "rowtile"):
        os.environ["FAB_ENGINE"] = engine
        all_ok = check(rank, world, device, engine) and all_ok
    dist.destroy_process_group()
    sys.exit(0 if all_ok else 1)


def check(rank, world, device, engine):
    B = 512 * world
    n_loc = B // world
    g = torch.Generator().manual_seed(99)
    eps = torch.randn(B, 32, generator=g)
    mom = torch.randn(6, 2, B, 32, generator=g)
    ex = torch.empty(6, 2, B).exponential_(1.0, generator=g)
    sl = slice(rank * n_loc, (rank + 1) * n_loc)

    # sharded run
    flow, op, ais, M = build(device, dist.group.WORLD)
    flow._eps_override = eps[sl].to(device)
    op.chain_noise_override = [(mom[j][:, sl].contiguous().to(device), ex[j][:, sl].contiguous().to(device))
                               for j in range(M)]
    pt_s, lw_s = ais.sample_and_log_weights(B)
    info_s = ais.get_logging_info()
    eps_s = (op.epsilons.clone(), op.common_epsilon.clone())

    # single-device run of the whole batch (every rank does it; no collectives)
    flow1, op1, ais1, _ = build(device, None)
    flow1._eps_override = eps.to(device)
    op1.chain_noise_override = [(mom[j].contiguous().to(device), ex[j].contiguous().to(device)) for j in range(M)]
    pt_1, lw_1 = ais1.sample_and_log_weights(B)
    info_1 = ais1.get_logging_info()

    def same(a, b):          # bitwise equality that treats NaN == NaN
        return torch.equal(torch.nan_to_num(a, nan=12345.0), torch.nan_to_num(b, nan=12345.0))

    ok = True
    checks = {}
    checks["log_w"] = same(lw_s, lw_1[sl])
    checks["x"] = same(pt_s.x, pt_1.x[sl])
    checks["log_q"] = same(pt_s.log_q, pt_1.log_q[sl])
    for k in ("ess_base", "ess_ais", "log_Z", "dist0_p_accept_0", "dist0_p_accept_1",
              "average_distance_dist0"):
        checks[k] = abs(info_s[k] - info_1[k]) <= 1e-5 * max(1.0, abs(info_1[k]))
    checks["tuner"] = torch.allclose(eps_s[0], op1.epsilons) and torch.allclose(eps_s[1], op1.common_epsilon)
    # global ESS trigger + systematic resample over the ranks (BASELINE config 3): every rank must
    # receive its slice of the single-device resample of the whole batch, bit for bit
    u0 = 987654321
    pt_r, lw_r, did = ais.resample_if_ess_below(pt_s, lw_s, threshold=1.1, u0=u0)      # always fires
    pt_w, lw_w, did1 = ais1.resample_if_ess_below(pt_1, lw_1, threshold=1.1, u0=u0)
    checks["resample fired"] = bool(did) and bool(did1)
    checks["resampled x"] = same(pt_r.x, pt_w.x[sl])
    checks["resampled grad_log_q"] = same(pt_r.grad_log_q, pt_w.grad_log_q[sl])
    checks["resampled log_w"] = torch.allclose(lw_r, lw_w[sl])
    ok = all(bool(v) for v in checks.values())
    if not ok:
        print(f"rank {rank}: failed checks: {[k for k, v in checks.items() if not v]}", flush=True)
    flag = torch.tensor([1.0 if ok else 0.0], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"[{engine}] world={world} B={B}: sharded == single-device: {'PASS' if flag.item() == 1.0 else 'FAIL'}"
              f"  (log_Z {info_s['log_Z']:.6f} vs {info_1['log_Z']:.6f}, ess_ais {info_s['ess_ais']:.6e} "
              f"vs {info_1['ess_ais']:.6e})", flush=True)
    return flag.item() == 1.0


if __name__ == "__main__":
    main()
