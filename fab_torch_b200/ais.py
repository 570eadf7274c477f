"""AnnealedImportanceSampler on the B200 kernels -- same constructor, attributes and return values
as fab/sampling_methods/ais.py:20-213, so it can replace `FABModel.annealed_importance_sampler`
(fab/core.py:65-73) or be bound over `fab.core.AnnealedImportanceSampler`.

One `sample_and_log_weights` call enqueues, without any host synchronisation in between:
  init kernel (flow sample + Point + log_w, ais.py:56-64)  ->  NaN/inf filter (ais.py:65,190-213,
  stable compaction on the device, live count kept in device memory)  ->  base ESS partial
  ->  M fused transitions (each also applies the log-weight update ais.py:93-100)
  ->  filter  ->  ESS / log Z reduction (ais.py:80-86, numerical.py:18-23)
and then reads back ONE small record (live counts, ESS, logsumexp) -- the only sync of the call.
With a process group, particles are sharded over ranks; the exchanges are the scalar triples of
SURVEY §8e (an all-gather of 4 floats per ESS, an all-reduce of 4 floats per HMC outer step when
the step-size tuner is on).
"""
from typing import Any, Dict, NamedTuple, Optional, Tuple

import numpy as np
import torch

from fab_torch_b200 import _lib
from fab_torch_b200 import dist as fdist
from fab_torch_b200.point import Point
from fab_torch_b200.transition_operators import TransitionOperator, make_gamma


class LoggingInfo(NamedTuple):
    ess_base: float
    ess_ais: float
    log_Z: float


def setup_distribution_spacing(distribution_spacing_type: str,
                               n_intermediate_distributions: int) -> torch.Tensor:
    """beta grid of M+2 points, float64 (ais.py:108-129)."""
    assert n_intermediate_distributions > 0
    M = n_intermediate_distributions
    if distribution_spacing_type == "geometric":
        n_lin = int(M / 4)
        n_geo = M - n_lin - 1
        B_space = np.concatenate([np.linspace(0, 0.01, n_lin + 2)[:-1],
                                  np.geomspace(0.01, 1, n_geo + 2)])
    elif distribution_spacing_type == "linear":
        B_space = np.linspace(0.0, 1.0, M + 2)
    else:
        raise Exception(f"distribution spacing incorrectly specified:"
                        f" '{distribution_spacing_type}',"
                        f"options are 'geometric' or 'linear'")
    assert B_space.shape == (M + 2,)
    return torch.tensor(B_space)


class AnnealedImportanceSampler:
    def __init__(self, base_distribution, target_log_prob, transition_operator: TransitionOperator,
                 p_target: bool, alpha: Optional[float] = None,
                 n_intermediate_distributions: int = 1,
                 distribution_spacing_type: str = "linear",
                 process_group=None, use_cuda_graph: bool = False):
        if not p_target:
            assert alpha is not None, "Must specify alpha if AIS target is not p."
        self.base_distribution = base_distribution
        self.target_log_prob = target_log_prob
        self.transition_operator = transition_operator
        self.p_target = p_target
        self.alpha = alpha
        self.n_intermediate_distributions = n_intermediate_distributions
        self.distribution_spacing_type = distribution_spacing_type
        self.B_space = setup_distribution_spacing(distribution_spacing_type,
                                                  n_intermediate_distributions)
        self._logging_info: LoggingInfo
        self.process_group = process_group
        if not (hasattr(base_distribution, "desc") and hasattr(base_distribution, "blob")):
            raise TypeError("base_distribution must be a B200RealNVP")
        target = getattr(target_log_prob, "__self__", None)
        if target is None or not hasattr(target, "target_desc"):
            raise TypeError("target_log_prob must be the log_prob of a fab_torch_b200 target")
        self._target = target
        self._host = None      # pinned read-back record
        self._filter_ws = None
        # use_cuda_graph: capture the whole chain (init, filters, M transitions, ESS) once per
        # (batch, mode) and replay it -- removes the per-kernel launch cost that dominates small
        # configurations (BASELINE config 1: 27 launches for ~0.3 ms of device work).  Single-GPU
        # only; the noise lives in persistent buffers that are refilled before every replay.
        self.use_cuda_graph = use_cuda_graph
        self._graphs = {}
        self._graph_warm = set()
        self._next_noise = None

    # ------------------------------------------------------------------------------------------
    def get_logging_info(self) -> Dict[str, Any]:
        logging_info = self._logging_info._asdict()
        logging_info.update(self.transition_operator.get_logging_info())
        return logging_info

    def _world(self):
        return fdist.world(self.process_group)

    def _filter(self, pt: Point, log_w, n_in, n_out):
        L = _lib.lib()
        n, d = pt.x.shape
        nbytes = int(L.fab_filter_workspace_bytes(n, d))
        if self._filter_ws is None or self._filter_ws.numel() < nbytes or \
                self._filter_ws.device != pt.x.device:
            self._filter_ws = torch.empty(nbytes, dtype=torch.uint8, device=pt.x.device)
        rc = L.fab_nan_filter_f32(_lib.point_ptrs(pt), _lib.ptr(log_w), d, n,
                                  _lib.ptr(n_in) if n_in is not None else None, _lib.ptr(n_out),
                                  _lib.ptr(self._filter_ws), _lib.stream_ptr(pt.x.device))
        _lib.check(rc, "fab_nan_filter_f32")

    def _ess(self, values, sub, n_active, out3):
        """ESS / logsumexp of (values - sub) over the live particles of ALL ranks -> out3."""
        L = _lib.lib()
        dev = values.device
        world, rank = self._world()
        part = torch.empty(4, dtype=torch.float32, device=dev)
        rc = L.fab_ess_partial_f32(_lib.ptr(values), _lib.ptr(sub) if sub is not None else None,
                                   values.shape[0], _lib.ptr(n_active), _lib.ptr(part),
                                   _lib.stream_ptr(dev))
        _lib.check(rc, "fab_ess_partial_f32")
        parts = fdist.gather_partials(part, self.process_group)
        rc = L.fab_ess_finalize_f32(_lib.ptr(parts), world, _lib.ptr(out3), _lib.stream_ptr(dev))
        _lib.check(rc, "fab_ess_finalize_f32")

    def _w_update(self, j: int):
        """(gamma_j, gamma_{j+1}) with the SAMPLER's alpha/p_target, or None when beta does not
        move (ais.py:93-105)."""
        if bool(self.B_space[j + 1] == self.B_space[j]):
            return None
        return (make_gamma(self.B_space[j], self.alpha, self.p_target),
                make_gamma(self.B_space[j + 1], self.alpha, self.p_target))

    def _run_chain(self, batch_size: int, logging: bool, timings=None, noise=None):
        """Enqueue the whole chain; returns device tensors + the device record (no sync).
        `noise` = (base eps [n,d], per-transition noise pairs) replaces the draws."""
        flow, op, target = self.base_distribution, self.transition_operator, self._target
        dev = flow._device()
        L = _lib.lib()
        d = flow.dim
        n = batch_size
        with_grad = bool(op.uses_grad_info)
        eps = noise[0] if noise is not None else getattr(flow, "_eps_override", None)
        if noise is not None:
            pass
        elif eps is not None:
            flow._eps_override = None
        else:
            eps = op.noise.base_eps(n, d, dev)
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        pt = Point(f(n, d), f(n), f(n), f(n, d) if with_grad else None,
                   f(n, d) if with_grad else None)
        log_w, log_q0 = f(n), f(n)
        valid = torch.empty(n, dtype=torch.uint8, device=dev)
        counts = torch.zeros(2, dtype=torch.int32, device=dev)
        rec = torch.zeros(8, dtype=torch.float32, device=dev)
        M = self.n_intermediate_distributions
        if self._c_chain_ok(op, timings):
            # single rank + HMC: the whole chain through ONE C-ABI call (fab_ais_chain_hmc_f32)
            chain_noise = noise[1] if noise is not None else op.chain_noise(M, n, d, dev)
            self._run_chain_c(pt, log_w, log_q0, valid, counts, rec, eps, chain_noise, logging)
            return pt, log_w, counts, rec
        g1 = make_gamma(self.B_space[1], self.alpha, self.p_target)
        tdesc = target.target_desc(dev)
        if with_grad and tdesc.kind == _lib.FAB_TARGET_MANYWELL and flow.use_rowtile(n):
            # same engine choice as the transitions (and as fab_ais_chain_hmc_f32): inverse-pass log q and
            # its gradient on the row-tile engine
            ws = op._workspace(int(L.fab_umma_workspace_bytes(flow.desc(), n)), dev)
            rc = L.fab_ais_init_umma_f32(flow.desc(), _lib.ptr(flow.blob()), _lib.ptr(flow.umma_blob()), tdesc,
                                         _lib.ptr(eps.contiguous()), g1, _lib.point_ptrs(pt), _lib.ptr(log_w),
                                         _lib.ptr(log_q0), _lib.ptr(valid), _lib.ptr(ws), n, _lib.stream_ptr(dev))
            _lib.check(rc, "fab_ais_init_umma_f32")
        else:
            rc = L.fab_ais_init_f32(flow.desc(), _lib.ptr(flow.blob()), tdesc,
                                    _lib.ptr(eps.contiguous()), g1, 1 if with_grad else 0,
                                    _lib.point_ptrs(pt), _lib.ptr(log_w), _lib.ptr(log_q0),
                                    _lib.ptr(valid), n, _lib.stream_ptr(dev))
            _lib.check(rc, "fab_ais_init_f32")
        self._filter(pt, log_w, None, counts[0:1])                 # "chain init"
        if logging:
            self._ess(pt.log_p, pt.log_q, counts[0:1], rec[0:3])   # ESS over base weights
        chain_noise = noise[1] if noise is not None else op.chain_noise(M, n, d, dev)
        kernel_timing = timings is not None and "timings" in op.run.__code__.co_varnames
        for j in range(1, M + 1):
            if timings is not None and not kernel_timing:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            if kernel_timing:       # events around the fused kernel only (not the tuner's collective)
                op.run(pt, j, self.B_space[j], log_w, self._w_update(j), n_active=counts[0:1],
                       noise=chain_noise[j - 1], timings=timings)
            else:
                op.run(pt, j, self.B_space[j], log_w, self._w_update(j), n_active=counts[0:1],
                       noise=chain_noise[j - 1])
            if timings is not None and not kernel_timing:
                e1.record()
                timings.append((e0, e1))
        self._filter(pt, log_w, counts[0:1], counts[1:2])          # "chain end"
        if logging:
            self._ess(log_w, None, counts[1:2], rec[3:6])
        if self._world()[0] > 1:
            # whole-batch live counts (the reference decides "no valid points" on the whole batch,
            # ais.py:202-204): summed on the device inside the chain, read back with the record
            import torch.distributed as dist
            gcounts = counts.to(torch.float32)
            dist.all_reduce(gcounts, group=self.process_group)
            rec[6:8] = gcounts
        return pt, log_w, counts, rec

    def _c_chain_ok(self, op, timings) -> bool:
        import os
        from fab_torch_b200.transition_operators import HamiltonianMonteCarlo
        return (timings is None and self._world()[0] == 1 and type(op) is HamiltonianMonteCarlo
                and self._target is op._target and op._flow is self.base_distribution
                and os.environ.get("FAB_C_CHAIN", "1") != "0")

    def _run_chain_c(self, pt, log_w, log_q0, valid, counts, rec, eps, chain_noise, logging):
        """ais.py:53-87 as one call into the library (same launches, in the same order, as the loop in
        `_run_chain`; bit-identical results -- tests/test_gpu_ais.py::test_c_chain_equals_python_loop)."""
        import ctypes as C
        flow, op, target = self.base_distribution, self.transition_operator, self._target
        dev = pt.x.device
        n, d = pt.x.shape
        L = _lib.lib()
        M = self.n_intermediate_distributions
        tdesc = target.target_desc(dev)
        rowtile = tdesc.kind == _lib.FAB_TARGET_MANYWELL and flow.use_rowtile(n)
        og = (_lib.Gamma * (M + 2))(*[make_gamma(self.B_space[j], op.alpha, op.p_target) for j in range(M + 2)])
        wg = (_lib.Gamma * (M + 2))(*[make_gamma(self.B_space[j], self.alpha, self.p_target) for j in range(M + 2)])
        wu = (C.c_uint8 * (M + 1))(*[0 if bool(self.B_space[j + 1] == self.B_space[j]) else 1 for j in range(M + 1)])
        moms = [a.contiguous() for a, _ in chain_noise]
        exps = [b.contiguous() for _, b in chain_noise]
        pm = (C.c_void_p * M)(*[_lib.ptr(a) for a in moms])
        pe = (C.c_void_p * M)(*[_lib.ptr(b) for b in exps])
        args = _lib.ChainHmcArgs(M, op.n_outer, op.L, 0 if op.eval_mode else 1, 1 if logging else 0,
                                 1 if rowtile else 0, op.target_p_accept, op.max_grad, og, wg, wu, pm, pe)
        nbytes = int(L.fab_ais_chain_workspace_bytes(flow.desc(), n, op.n_outer, 1 if rowtile else 0))
        _lib.check(nbytes, "fab_ais_chain_workspace_bytes")
        ws = op._workspace(nbytes, dev)
        rc = L.fab_ais_chain_hmc_f32(flow.desc(), _lib.ptr(flow.blob()),
                                     _lib.ptr(flow.umma_blob()) if rowtile else None, tdesc, op._state(),
                                     C.byref(args), _lib.ptr(eps.contiguous()), _lib.point_ptrs(pt), _lib.ptr(log_w),
                                     _lib.ptr(log_q0), _lib.ptr(valid), _lib.ptr(counts), _lib.ptr(rec),
                                     _lib.ptr(op._stats), _lib.ptr(ws), n, _lib.stream_ptr(dev))
        _lib.check(rc, "fab_ais_chain_hmc_f32")
        op._seen_first = True
        op._seen_last = True

    def release_graphs(self):
        """Drop the captured CUDA graphs (and their persistent buffers).  With a process group, call
        this before `destroy_process_group`: the graphs hold the communicator's kernels."""
        torch.cuda.synchronize()
        self._graphs.clear()
        self._graph_warm.clear()

    def set_next_noise(self, eps: torch.Tensor, noise_a: torch.Tensor, noise_b: torch.Tensor):
        """Pre-drawn randomness for the NEXT `sample_and_log_weights` call, e.g. pinned host
        tensors (copied host->device inside the call): eps [n,d]; HMC: momentum [M,n_outer,n,d] +
        exponential [M,n_outer,n]; Metropolis: proposal [M,n_updates,n,d] + uniform [M,n_updates,n]."""
        self._next_noise = (eps, noise_a, noise_b)

    def _run_graphed(self, local: int, logging: bool):
        """Chain through a captured CUDA graph: first call with a given key runs eagerly (warm-up:
        workspaces, weight blob, kernel attributes), the second captures, later ones replay."""
        flow, op = self.base_distribution, self.transition_operator
        dev = flow._device()
        d, M = flow.dim, self.n_intermediate_distributions
        blob = flow.blob()                                   # repacked in place if parameters changed
        # everything the captured launches bake in: raw pointers of the weight images, of the
        # operator's state buffers and of the target tables, and the scalar hyper-parameters
        state_ptrs = tuple(b.data_ptr() for b in op.buffers())
        tdesc = self._target.target_desc(dev)
        tkey = tuple(getattr(tdesc, f) for f, _ in tdesc._fields_)
        hyper = tuple(getattr(op, k, None) for k in ("L", "n_outer", "max_grad", "target_p_accept",
                                                     "n_updates", "target_prob_accept"))
        ukey = flow.umma_blob().data_ptr() if flow.use_rowtile(local) else 0
        key = (local, self._world()[0], logging, self.p_target, self.alpha, op.p_target, op.alpha,
               bool(getattr(op, "eval_mode", False)), getattr(op, "adjust_step_size", None),
               blob.data_ptr(), ukey, id(op), str(dev), state_ptrs, tkey, hyper,
               _lib.engine_choice())
        ent = self._graphs.get(key)
        if ent is None:
            ent = dict(eps=torch.empty(local, d, dtype=torch.float32, device=dev),
                       noise=op.chain_noise_static(M, local, d, dev), graph=None)
            self._graphs[key] = ent
        # this call's randomness -> the persistent buffers (same generator calls as the eager path)
        if self._next_noise is not None:
            (e, a, b), self._next_noise = self._next_noise, None
            ent["eps"].copy_(e, non_blocking=True)
            ent["noise"][0].copy_(a, non_blocking=True)
            ent["noise"][1].copy_(b, non_blocking=True)
        else:
            eps = getattr(flow, "_eps_override", None)
            if eps is not None:
                flow._eps_override = None
                ent["eps"].copy_(eps)
            elif type(op.noise).__name__ != "DeviceNoise":
                ent["eps"].copy_(op.noise.base_eps(local, d, dev))
            else:
                ent["eps"].normal_()
            op.fill_chain_noise(ent["noise"])
        pairs = [(ent["noise"][0][j], ent["noise"][1][j]) for j in range(M)]
        if key not in self._graph_warm:
            self._graph_warm.add(key)
            return self._run_chain(local, logging, noise=(ent["eps"], pairs))
        if ent["graph"] is None:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                pt, log_w, counts, rec = self._run_chain(local, logging, noise=(ent["eps"], pairs))
            ent.update(graph=g, out=(pt, log_w, counts, rec))
        ent["graph"].replay()
        pt, log_w, counts, rec = ent["out"]
        c = lambda t: None if t is None else t.clone()
        return (Point(c(pt.x), c(pt.log_q), c(pt.log_p), c(pt.grad_log_q), c(pt.grad_log_p)),
                log_w.clone(), counts, rec)

    def time_transitions(self, batch_size: int, repeats: int = 1) -> float:
        """Mean device time (ms) of one fused transition launch, measured with CUDA events on the
        launching stream around every transition of `repeats` chains (bench.py roofline)."""
        world, rank = self._world()
        local = fdist.shard_size(batch_size, self.process_group)
        self.transition_operator.process_group = self.process_group
        total, count = 0.0, 0
        for _ in range(repeats):
            timings = []
            self._run_chain(local, False, timings=timings)
            torch.cuda.current_stream().synchronize()
            total += sum(a.elapsed_time(b) for a, b in timings)
            count += len(timings)
        return total / max(count, 1)

    def sample_and_log_weights(self, batch_size: int, logging: bool = True
                               ) -> Tuple[Point, torch.Tensor]:
        world, rank = self._world()
        op = self.transition_operator
        op.process_group = self.process_group
        local = fdist.shard_size(batch_size, self.process_group)
        if self.use_cuda_graph:
            # (with a process group the tuner / ESS collectives are captured with the kernels: NCCL
            # supports stream capture; every rank captures and replays the same sequence)
            pt, log_w, counts, rec = self._run_graphed(local, logging)
        else:
            if self._next_noise is not None:
                (e, a, b), self._next_noise = self._next_noise, None
                dev_ = self.base_distribution._device()
                a, b = a.to(dev_, non_blocking=True), b.to(dev_, non_blocking=True)
                pt, log_w, counts, rec = self._run_chain(
                    local, logging, noise=(e.to(dev_, non_blocking=True),
                                           [(a[j], b[j]) for j in range(a.shape[0])]))
            else:
                pt, log_w, counts, rec = self._run_chain(local, logging)
        dev = log_w.device
        if self._host is None:
            self._host = torch.empty(10, dtype=torch.float32).pin_memory()
        both = torch.cat([counts.to(torch.float32), rec])
        self._host.copy_(both, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()               # the one sync of the call
        h = self._host.numpy()
        n_init, n_end = int(h[0]), int(h[1])
        g_init, g_end = n_init, n_end
        if world > 1:
            # the reference decides on the WHOLE batch (ais.py:202-204); deciding per shard would let
            # one rank raise while the others wait in the next collective
            g_init, g_end = int(h[2 + 6]), int(h[2 + 7])
        if g_init == 0:
            raise Exception("No valid points generated in sampling the chain init")
        if g_end == 0:
            raise Exception("No valid points generated in sampling the chain end")
        if n_init != local:
            print(f"{local - n_init} nan/inf samples/log-probs/log-weights encountered at chain init.")
        if n_end != n_init:
            print(f"{n_init - n_end} nan/inf samples/log-probs/log-weights encountered at chain end.")
        if logging:
            log_Z = np.float32(h[2 + 4]) - np.log(np.float32(batch_size))     # ais.py:83-84 (quirk 6)
            self._logging_info = LoggingInfo(ess_base=float(h[2 + 0]), ess_ais=float(h[2 + 3]),
                                             log_Z=float(log_Z))
        if n_end != local:
            pt = pt[slice(0, n_end)]
            log_w = log_w[:n_end]
        return pt, log_w.detach()

    def resample_if_ess_below(self, point: Point, log_w: torch.Tensor, threshold: float,
                              u0: Optional[int] = None):
        """Global ESS / resample trigger (build-side extension): uses the
        all-reduced `ess_ais` of the last `sample_and_log_weights` call; below `threshold` the
        (rank-sharded) particle set is resampled systematically with the bit-exact integer kernel
        and the weights are reset to the mean weight.  Returns (point, log_w, resampled?)."""
        from fab_torch_b200.resample import resample_if_ess_below
        return resample_if_ess_below(point, log_w, self._logging_info.ess_ais, threshold, u0,
                                     self.process_group)

    # kept for API parity with the reference sampler (ais.py:90-105); runs one fused transition
    def perform_transition(self, x_new: Point, log_w: torch.Tensor, j: int):
        x_new = self.transition_operator.run(x_new, j, self.B_space[j], log_w, self._w_update(j))
        return x_new, log_w

    def generate_eval_data(self, outer_batch_size: int, inner_batch_size: int):
        """ais.py:132-188: batched chains for evaluation; results concatenated on the CPU."""
        base_samples, base_log_w_s, ais_samples, ais_log_w = [], [], [], []
        assert outer_batch_size % inner_batch_size == 0
        for _ in range(outer_batch_size // inner_batch_size):
            flow, op = self.base_distribution, self.transition_operator
            with torch.no_grad():
                x, log_q0 = flow.sample_and_log_prob((inner_batch_size,))
            point = op.create_new_point(x)
            if not op.uses_grad_info:
                point.log_q = log_q0.detach()
            base_log_w = point.log_p - log_q0
            ok = torch.isfinite(point.log_p) & torch.isfinite(point.log_q)
            if not bool(ok.any()):              # ais.py:202-204 (raise_exception=True at chain init)
                raise Exception("No valid points generated in sampling the chain init")
            if not bool(ok.all()):
                print(f"{int((~ok).sum())} nan/inf samples/log-probs/log-weights encountered at chain init.")
                point, base_log_w = point[ok], base_log_w[ok]
                point = Point(*(None if t is None else t.contiguous() for t in
                                (point.x, point.log_q, point.log_p, point.grad_log_q,
                                 point.grad_log_p)))
            base_samples.append(point.x.detach().cpu())
            base_log_w_s.append(base_log_w.detach().cpu())
            g1 = make_gamma(self.B_space[1], self.alpha, self.p_target)
            log_w = (g1.cq * point.log_q + g1.cp * point.log_p - point.log_q).contiguous()
            for j in range(1, self.n_intermediate_distributions + 1):
                point, log_w = self.perform_transition(point, log_w, j)
            ok = torch.isfinite(point.log_p) & torch.isfinite(point.log_q)
            if not bool(ok.any()):              # raise_exception=False at chain end (ais.py:177-178)
                print("No valid points generated in sampling the chain end")
            elif not bool(ok.all()):
                print(f"{int((~ok).sum())} nan/inf samples/log-probs/log-weights encountered at chain end.")
                point, log_w = point[ok], log_w[ok]
            ais_samples.append(point.x.detach().cpu())
            ais_log_w.append(log_w.detach().cpu())
        return (torch.cat(base_samples, dim=0), torch.cat(base_log_w_s, dim=0),
                torch.cat(ais_samples, dim=0), torch.cat(ais_log_w, dim=0))
