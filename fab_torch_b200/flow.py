"""B200RealNVP: the reference's RealNVP-style coupling flow behind the `TrainableDistribution`
surface (fab/wrappers/normflows.py:8-31), evaluated by the sm_100a tile kernels.

Architecture = experiments/make_flow/make_normflow_model.py:11-30,82-96 (normflows):
DiagGaussian base, then per layer AffineCouplingBlock(MLP[d1, W, W, 2*d2], exp scale) followed by
an LU-parameterised InvertibleAffine.  Parameter/module names follow normflows so checkpoints
written by `FABModel.save` (fab/core.py:222-260) keep the reference's state-dict keys
(`_nf_model.q0.loc`, `_nf_model.flows.0.flows.1.param_map.net.0.weight`,
`_nf_model.flows.1.{P,L,U,log_S,sign_S,eye}`), and default initialisation consumes the torch RNG
in the same order as normflows (so `torch.manual_seed(s)` gives the same weights).

Hot path: `sample_and_log_prob`, `log_prob` and d log_prob / dx run as CUDA kernels on a packed
weight blob (layout: include/fab_b200.h) that is rebuilt lazily whenever a parameter changes.
Gradients of `log_prob` w.r.t. the flow parameters (the training loss, fab/core.py:112-118 -- SURVEY
§8f row 2) are CUDA kernels too: the forward pass of `log_prob` records an activation tape
(`fab_flow_logprob_tape_f32`) whenever a parameter requires grad, and `backward` turns it into the
weight gradients with batch-contraction GEMMs (`fab_flow_param_grad_f32`, csrc/param_grad.cuh); what
is left for torch is the chain rule through the merged / LU-parameterised matrices in PARAMETER
space ([W x d]-sized).  `sample_and_log_prob` (not on the FAB-loss path: the AIS points are detached,
fab/core.py:120-128) still differentiates w.r.t. parameters by re-running the torch-op restatement.
"""
import math
from typing import Tuple

import torch
import torch.nn as nn

from fab_torch_b200 import _lib
from fab_torch_b200.types_ import TrainableDistribution


# --------------------------------------------------------------------------- parameter containers
class _DiagGaussian(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.shape = (dim,)
        self.loc = nn.Parameter(torch.zeros(1, dim))
        self.log_scale = nn.Parameter(torch.zeros(1, dim))


class _MLP(nn.Module):
    def __init__(self, sizes, init_zeros=True):
        super().__init__()
        mods = []
        for a, b in zip(sizes[:-2], sizes[1:-1]):
            mods += [nn.Linear(a, b), nn.LeakyReLU(0.0)]
        mods.append(nn.Linear(sizes[-2], sizes[-1]))
        if init_zeros:
            nn.init.zeros_(mods[-1].weight)
            nn.init.zeros_(mods[-1].bias)
        self.net = nn.Sequential(*mods)


class _AffineCoupling(nn.Module):
    def __init__(self, param_map):
        super().__init__()
        self.add_module("param_map", param_map)


class _CouplingBlock(nn.Module):
    """flows[1] holds the conditioner, like normflows' [Split, AffineCoupling, Merge]."""

    def __init__(self, param_map):
        super().__init__()
        self.flows = nn.ModuleList([nn.Identity(), _AffineCoupling(param_map), nn.Identity()])

    @property
    def linears(self):
        net = self.flows[1].param_map.net
        return net[0], net[2], net[4]


class _InvertibleAffine(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        Q, _ = torch.linalg.qr(torch.randn(dim, dim))
        LU, piv = torch.linalg.lu_factor(Q)
        P, L, U = torch.lu_unpack(LU, piv)
        self.register_buffer("P", P)
        self.L = nn.Parameter(L)
        S = U.diag()
        self.register_buffer("sign_S", torch.sign(S))
        self.log_S = nn.Parameter(torch.log(torch.abs(S)))
        self.U = nn.Parameter(torch.triu(U, diagonal=1))
        self.register_buffer("eye", torch.diag(torch.ones(dim)))


class _ActNorm(nn.Module):
    """normflows' ActNorm(dim): z * exp(s) + t, parameters of shape [1, d]; `data_dep_init_done`
    flips to 1 once s, t have been set from the statistics of the first batch."""

    def __init__(self, dim: int):
        super().__init__()
        self.s = nn.Parameter(torch.zeros(1, dim))
        self.t = nn.Parameter(torch.zeros(1, dim))
        self.register_buffer("data_dep_init_done", torch.tensor(0.0))


class _NFModel(nn.Module):
    def __init__(self, dim, n_layers, width, act_norm=False):
        super().__init__()
        self.q0 = _DiagGaussian(dim)
        d1 = int((dim / 2) + 0.5)
        flows = []
        for _ in range(n_layers):
            flows.append(_CouplingBlock(_MLP([d1, width, width, 2 * (dim - d1)])))
            flows.append(_InvertibleAffine(dim))
            if act_norm:
                flows.append(_ActNorm(dim))
        self.flows = nn.ModuleList(flows)


# --------------------------------------------------------------------------- packing helpers
def _r4(v: int) -> int:
    return (v + 3) // 4 * 4


def _r8(v: int) -> int:
    return (v + 7) // 8 * 8


def _r16(v: int) -> int:
    return (v + 15) // 16 * 16


def _round22(M: torch.Tensor) -> torch.Tensor:
    """Round fp32 values to 22 significant bits = tf32 hi (top 11 bits, truncated) + tf32 lo (the
    remainder rounded to nearest on the tf32 grid), so that the kernels' hi/lo operand split is
    exact (mma_gemm.cuh: split_w).  Relative change <= 2^-23, unbiased."""
    mask = -8192                                          # 0xffffe000 as int32
    hi = (M.contiguous().view(torch.int32) & mask).view(torch.float32)
    lo = M - hi
    lo_r = ((lo.view(torch.int32) + 0x1000) & mask).view(torch.float32)
    return hi + lo_r


def _pack_frag(M: torch.Tensor, K16: int, N8: int) -> torch.Tensor:
    """M [Lyr, K, N] -> MMA fragment order [Lyr, (K16/16)*(N8/8)*128] (see include/fab_b200.h):
    Wf[kp][nt][lane = 4g+t] = (M[16kp+t][8nt+g], M[16kp+t+4][.], M[16kp+8+t][.], M[16kp+12+t][.])."""
    Lyr, K, N = M.shape
    out = M.new_zeros(Lyr, K16, N8)
    out[:, :K, :N] = _round22(M)
    # k = 16*kp + 8*h + 4*half + t ; n = 8*nt + g  ->  [kp][nt][g][t][h][half]
    v = out.view(Lyr, K16 // 16, 2, 2, 4, N8 // 8, 8)
    return v.permute(0, 1, 5, 6, 4, 2, 3).reshape(Lyr, -1)


def _pad_last(v: torch.Tensor, n: int) -> torch.Tensor:
    out = v.new_zeros(*v.shape[:-1], n)
    out[..., :v.shape[-1]] = v
    return out


class B200RealNVP(TrainableDistribution):
    """Drop-in for `make_wrapped_normflow_realnvp(dim, n_flow_layers, layer_nodes_per_dim,
    act_norm)` (make_normflow_model.py:82-96).

    `act_norm=True` appends an ActNorm layer (z * exp(s) + t) to every block (make_normflow_model.py:
    28-29) and, like the reference factory (:94-95), draws 500 samples once to set s, t from the
    batch statistics.  The kernels never see the layer: the host folds it into the block's linear
    part (Wmix := diag(exp(-s)) W, bias c = -t @ Wmix, log|det| = sum(log_S) - sum(s); include/
    fab_b200.h) when it packs the weights (both engines)."""

    def __init__(self, dim: int, n_flow_layers: int = 5, layer_nodes_per_dim: int = 10,
                 act_norm: bool = False):
        super().__init__()
        self.dim = dim
        self.n_flow_layers = n_flow_layers
        self.width = dim * layer_nodes_per_dim
        self.act_norm = bool(act_norm) and n_flow_layers > 0
        self._nf_model = _NFModel(dim, n_flow_layers, self.width, self.act_norm)
        self._desc = _lib.FlowDesc()
        self._blob = None
        self._blob_key = None
        self._ublob = None              # row-tile engine weight images (f16 hi/lo planes)
        self._ublob_key = None
        self._uws = None
        self._plist = None              # cached parameter list (see _param_key)
        self._perm_dev = None           # device copy of the [shifts | scales] row permutation of W3
        self._pg_layouts = {}           # fab_flow_param_grad_layout per batch size
        self._pg_states = {}            # persistent buffers of cuda_param_grad per batch size
        self._pack_graphs = {}          # captured repack launches (see _repack)
        self._eps_override = None       # test hook: next sample uses this base noise
        # filled lazily: needs the .so
        self._desc_ready = False
        if self.act_norm:
            self._init_act_norm(500)

    # ---- layer access ------------------------------------------------------------------
    def _blocks(self):
        st = 3 if self.act_norm else 2
        return [self._nf_model.flows[st * k] for k in range(self.n_flow_layers)]

    def _mixes(self):
        st = 3 if self.act_norm else 2
        return [self._nf_model.flows[st * k + 1] for k in range(self.n_flow_layers)]

    def _acts(self):
        return [self._nf_model.flows[3 * k + 2] for k in range(self.n_flow_layers)] if self.act_norm else []

    @torch.no_grad()
    def _init_act_norm(self, n: int):
        """Data-dependent initialisation of the ActNorm layers (normflows ActNorm.forward on its first
        call, triggered by `sample((500,))` in make_normflow_model.py:94-95): walking the layers in the
        sampling direction, s = -log(std + 1e-6), t = -mean * exp(s) of the layer's input batch.
        One-time parameter initialisation in torch ops on the parameters' device (consumes
        randn(n, d) from the global generator, like the reference's call)."""
        q0 = self._nf_model.q0
        d1 = int((self.dim / 2) + 0.5)
        z = q0.loc + torch.exp(q0.log_scale) * torch.randn((n, self.dim), dtype=q0.loc.dtype, device=q0.loc.device)
        for blk, mix, act in zip(self._blocks(), self._mixes(), self._acts()):
            l1, l2, l3 = blk.linears
            z1, z2 = z[:, :d1], z[:, d1:]
            par = l3(torch.relu(l2(torch.relu(l1(z1)))))
            z = torch.cat([z1, z2 * torch.exp(par[:, 1::2]) + par[:, 0::2]], dim=1)
            Lf = torch.tril(mix.L, diagonal=-1) + mix.eye
            Uf = torch.triu(mix.U, diagonal=1) + torch.diag(mix.sign_S * torch.exp(mix.log_S))
            W_inv = torch.inverse(Uf.double()).to(z.dtype) @ torch.inverse(Lf.double()).to(z.dtype) @ mix.P.t()
            z = z @ W_inv
            if not act.data_dep_init_done > 0.0:
                act.s.copy_(-torch.log(z.std(dim=0, keepdim=True) + 1e-6))
                act.t.copy_(-z.mean(dim=0, keepdim=True) * torch.exp(act.s))
                act.data_dep_init_done.fill_(1.0)
            z = z * torch.exp(act.s) + act.t

    # ---- descriptor / blob -------------------------------------------------------------
    def desc(self) -> "_lib.FlowDesc":
        if not self._desc_ready:
            n = _lib.lib().fab_flow_desc_init(self._desc, self.dim, max(self.width, 1),
                                              self.n_flow_layers)
            _lib.check(n, "fab_flow_desc_init")
            self._desc_ready = True
        return self._desc

    def _param_key(self):
        """Changes whenever a parameter is updated in place (optimiser step, load_state_dict: version
        counters) or moved (.to(): the cached list is dropped in `_apply`, and three storage addresses
        are part of the key).  Walking `self.parameters()` and reading every data_ptr cost 0.3 ms per
        call at config 2 -- twice per chain call -- so the list is cached and only versions are read."""
        ps = self._plist
        if ps is None:
            ps = self._plist = list(self.parameters())
        if not ps:
            return ()
        return (tuple(p._version for p in ps), ps[0].data_ptr(), ps[len(ps) // 2].data_ptr(), ps[-1].data_ptr())

    def _apply(self, fn, *args, **kwargs):
        self._plist = None
        self._pack_graphs = {}          # parameter storage may move: the captured pointers go stale
        self._pg_states = {}
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._plist = None
        return super().load_state_dict(*args, **kwargs)

    def _repack(self, which: str, fn):
        """Run the repack `fn` (writes a persistent blob from the current parameters).  On CUDA the
        ~100 small torch launches of a repack are captured once and replayed on every later parameter
        update (3.2 -> ~0.2 ms per optimiser step at config 2): first call eager, second captures.
        FAB_PACK_GRAPH=0 keeps it eager."""
        import os
        st = self._pack_graphs.setdefault(which, dict(warm=False, graph=None))
        on_cuda = self._device().type == "cuda" and os.environ.get("FAB_PACK_GRAPH", "1") != "0" \
            and not torch.cuda.is_current_stream_capturing()
        if not on_cuda or not st["warm"]:
            fn()
            st["warm"] = True
            return
        if st["graph"] is None:
            # (thread-local capture mode: `backward` runs this on the autograd thread while other
            # threads of the process -- data loaders, loggers -- may be calling into CUDA)
            try:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    fn()
                st["graph"] = g
            except Exception as e:               # noqa: BLE001 -- a failed capture only costs the replay speed-up
                import warnings
                warnings.warn(f"fab_torch_b200: CUDA-graph capture of '{which}' failed ({type(e).__name__}: {e}); "
                              "running it eagerly from now on")
                st["graph"] = False
                torch.cuda.synchronize()
        if st["graph"] is False:
            fn()
            return
        st["graph"].replay()

    def blob(self) -> torch.Tensor:
        """Packed fp32 weights on the parameters' device, rebuilt only when a parameter changed."""
        key = self._param_key()
        if self._blob is None or key != self._blob_key:
            with torch.no_grad():
                if self._blob is None or self._blob.device != self._device():
                    self._blob = self._pack()
                    self._pack_graphs.pop("blob", None)
                else:                          # same storage: captured CUDA graphs stay valid
                    self._repack("blob", lambda: self._blob.copy_(self._pack()))
            self._blob_key = key
        return self._blob

    def _act_st(self):
        """(s, t) of the ActNorm layers, [K, d] each; None without ActNorm."""
        if not self.act_norm:
            return None, None
        acts = self._acts()
        return torch.stack([a.s.reshape(-1) for a in acts]), torch.stack([a.t.reshape(-1) for a in acts])

    def _fold_act(self, W, W_inv, logs):
        """ActNorm folded into the block's linear part: in the log_prob direction the block starts with
        ((z - t) exp(-s)) @ W = z @ W' + c, W' = diag(exp(-s)) W, c = -t @ W'; when sampling it ends with
        (z @ W^-1) exp(s) + t = z @ W'^-1 + t.  Returns (W', W'^-1, log|det|, c, t); c = t = None
        without ActNorm."""
        s_, t_ = self._act_st()
        if s_ is None:
            return W, W_inv, logs, None, None
        s_, t_ = s_.to(W.dtype), t_.to(W.dtype)
        We = torch.exp(-s_)[:, :, None] * W
        We_inv = W_inv * torch.exp(s_)[:, None, :] if W_inv is not None else None
        c = -(t_[:, None, :] @ We)[:, 0, :]
        return We, We_inv, logs - s_.sum(dim=1), c, t_

    def _mixing(self, dtype=torch.float32):
        """Effective W [K,d,d], W^-1 and log|det| [K] of every block's linear part plus the biases
        (c, t) of `_fold_act` (differentiable)."""
        mixes = self._mixes()
        P = torch.stack([m.P for m in mixes])
        eye = mixes[0].eye
        L = torch.tril(torch.stack([m.L for m in mixes]), diagonal=-1) + eye
        log_S = torch.stack([m.log_S for m in mixes])
        sign_S = torch.stack([m.sign_S for m in mixes])
        U = torch.triu(torch.stack([m.U for m in mixes]), diagonal=1) + \
            torch.diag_embed(sign_S * torch.exp(log_S))
        W = P @ L @ U
        L_inv = torch.inverse(L.double()).to(dtype)
        U_inv = torch.inverse(U.double()).to(dtype)
        W_inv = U_inv @ L_inv @ P.transpose(1, 2)
        return self._fold_act(W, W_inv, log_S.sum(dim=1))

    def _mixing_pack(self, need_inverse: bool = True):
        """`_mixing` for the weight packers: no autograd, and the float64 inverses of the triangular
        factors by triangular solves (no pivoting, no host-side info check: capturable in a CUDA graph)."""
        mixes = self._mixes()
        P = torch.stack([m.P for m in mixes])
        eye = mixes[0].eye
        L = torch.tril(torch.stack([m.L for m in mixes]), diagonal=-1) + eye
        log_S = torch.stack([m.log_S for m in mixes])
        sign_S = torch.stack([m.sign_S for m in mixes])
        U = torch.triu(torch.stack([m.U for m in mixes]), diagonal=1) + \
            torch.diag_embed(sign_S * torch.exp(log_S))
        W = P @ L @ U
        W_inv = None
        if need_inverse:
            eye64 = eye.double().expand_as(L)
            L_inv = torch.linalg.solve_triangular(L.double(), eye64, upper=False).float()
            U_inv = torch.linalg.solve_triangular(U.double(), eye64, upper=True).float()
            W_inv = U_inv @ L_inv @ P.transpose(1, 2)
        return self._fold_act(W, W_inv, log_S.sum(dim=1))

    def _pack(self) -> torch.Tensor:
        d = self.desc()
        dev = self._nf_model.q0.loc.device
        if self._nf_model.q0.loc.dtype != torch.float32:
            raise RuntimeError("B200RealNVP kernels are fp32; keep the flow in float32")
        DP, D8, D16 = _r4(d.dim), _r8(d.dim), _r16(d.dim)
        D1K, P8, P16 = _r16(d.d1), _r8(2 * d.d2), _r16(2 * d.d2)
        W8, W16 = d.width_pad, d.width_kpad
        base = torch.cat([_pad_last(self._nf_model.q0.loc.reshape(-1), DP),
                          _pad_last(self._nf_model.q0.log_scale.reshape(-1), DP)])
        K = self.n_flow_layers
        tail = base.new_zeros(512)
        if K == 0:
            return torch.cat([base, tail]).contiguous()
        blocks = self._blocks()
        W1 = torch.stack([b.linears[0].weight for b in blocks])     # [K, W, d1]
        b1 = torch.stack([b.linears[0].bias for b in blocks])
        W2 = torch.stack([b.linears[1].weight for b in blocks])     # [K, W, W]
        b2 = torch.stack([b.linears[1].bias for b in blocks])
        W3 = torch.stack([b.linears[2].weight for b in blocks])     # [K, 2*d2, W]
        b3 = torch.stack([b.linears[2].bias for b in blocks])
        perm = torch.cat([torch.arange(0, 2 * d.d2, 2, device=dev), torch.arange(1, 2 * d.d2, 2, device=dev)])
        W3 = W3[:, perm, :]
        b3 = b3[:, perm]
        Wm, Wm_inv, logs, cmix, tmix = self._mixing_pack()
        t = lambda M: M.transpose(1, 2)
        dd, d1, W = d.dim, d.d1, d.width
        z = lambda r, c: W1.new_zeros(K, r, c)
        # merged operands (products in float64, rounded once)
        mw1 = z(dd, D8 + W)                            # o_mw1: z -> [v | h1pre]
        mw1[:, :, :dd] = Wm
        mw1[:, :, D8:D8 + W] = (Wm[:, :, :d1].double() @ t(W1).double()).float()
        w1mt = z(W16 + dd, dd)                         # o_w1mt: [gh1 | gv] -> g_u
        w1mt[:, :W, :] = (W1.double() @ t(Wm[:, :, :d1]).double()).float()
        w1mt[:, W16:W16 + dd, :] = t(Wm)
        b1e = z(1, D8 + W8)[:, 0]                      # o_b1: [c | b1 + c[:d1] @ W1^T]   (c = 0 without ActNorm)
        b1e[:, D8:D8 + W] = b1
        if cmix is not None:
            b1e[:, :dd] = cmix
            b1e[:, D8:D8 + W] = (b1.double() + (cmix[:, None, :d1].double() @ t(W1).double())[:, 0, :]).float()
        parts = [
            _pack_frag(mw1, D16, D8 + W8),
            _pack_frag(t(W2), W16, W8),                # o_w2      M[k][n] = W2[n][k]
            _pack_frag(t(W3), W16, P8),                # o_w3      M[k][n] = W3p[n][k]
            _pack_frag(W3, P16, W8),                   # o_w3t     M[k][n] = W3p[k][n]
            _pack_frag(W2, W16, W8),                   # o_w2t     M[k][n] = W2[k][n]
            _pack_frag(w1mt, W16 + D16, D8),
            _pack_frag(t(W1), D1K, W8),                # o_w1      M[k][n] = W1[n][k]
            _pack_frag(Wm_inv, D16, D8),               # o_mix_inv
            b1e,
            _pad_last(b2, W8),
            _pad_last(b3, P8),
            _pad_last(logs[:, None], 4),
            _pad_last(b1, W8),                         # o_b1s
            _pad_last(tmix, D8) if tmix is not None else z(1, D8)[:, 0],   # o_tmix
        ]
        layers = torch.cat(parts, dim=1)
        assert layers.shape[1] == d.layer_stride, (layers.shape, d.layer_stride)
        blob = torch.cat([base, layers.reshape(-1), tail]).contiguous()
        assert blob.numel() == d.total_floats
        return blob

    # ---- row-tile engine (tcgen05): weight images and engine choice ------------------------------
    def rowtile_supported(self) -> bool:
        return bool(_lib.lib().fab_umma_supported(self.desc()))

    def use_rowtile(self, n: int) -> bool:
        mode = _lib.engine_choice()
        if mode == "warp":
            return False
        if not self.rowtile_supported():
            if mode == "rowtile":
                raise RuntimeError("FAB_ENGINE=rowtile: this flow shape is not covered by the row-tile "
                                   "engine (needs dim 32, width 64..320 in steps of 64, <= 10 layers)")
            return False
        return mode == "rowtile" or n >= _lib.rowtile_min_n()

    def umma_blob(self) -> torch.Tensor:
        """Weight images of the row-tile engine (include/fab_b200.h), rebuilt on the device when a
        parameter changed: merged fp32 matrices in a plain buffer -> fab_umma_pack_f32."""
        key = self._param_key()
        if self._ublob is None or key != self._ublob_key:
            with torch.no_grad():
                L = _lib.lib()
                dev = self._device()
                nbytes = int(L.fab_umma_blob_bytes(self.desc()))
                _lib.check(nbytes, "fab_umma_blob_bytes")
                if self._ublob is None or self._ublob.numel() != nbytes or self._ublob.device != dev:
                    self._ublob = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
                    self._pack_graphs.pop("ublob", None)

                def pack_images():
                    plain = self._plain_umma()
                    rc = L.fab_umma_pack_f32(self.desc(), _lib.ptr(plain), _lib.ptr(self._ublob),
                                             _lib.stream_ptr(dev))
                    _lib.check(rc, "fab_umma_pack_f32")
                self._repack("ublob", pack_images)
            self._ublob_key = key
        return self._ublob

    def _plain_umma(self) -> torch.Tensor:
        """[loc | log_scale | per layer: the seven operand matrices M[k][n] (+ bias rows), sum(log_S)]
        in the float layout `fab_umma_plain_layout` reports.  Merged products in float64."""
        import ctypes as C
        d = self.desc()
        offs = (C.c_int64 * 32)()
        _lib.check(_lib.lib().fab_umma_plain_layout(d, offs), "fab_umma_plain_layout")
        total, off_layers, per_layer, logs_off = (int(offs[i]) for i in range(4))
        K, dd, d1, W = self.n_flow_layers, d.dim, d.d1, d.width
        blocks = self._blocks()
        W1 = torch.stack([b.linears[0].weight for b in blocks])     # [K, W, d1]
        b1 = torch.stack([b.linears[0].bias for b in blocks])
        W2 = torch.stack([b.linears[1].weight for b in blocks])     # [K, W, W]
        b2 = torch.stack([b.linears[1].bias for b in blocks])
        W3 = torch.stack([b.linears[2].weight for b in blocks])     # [K, 2*d2, W]
        b3 = torch.stack([b.linears[2].bias for b in blocks])
        dev = W1.device
        perm = torch.cat([torch.arange(0, 2 * d.d2, 2, device=dev), torch.arange(1, 2 * d.d2, 2, device=dev)])
        W3, b3 = W3[:, perm, :], b3[:, perm]
        Wm, _, logs, cmix, _ = self._mixing_pack(need_inverse=False)
        t = lambda M: M.transpose(1, 2)
        mw1 = torch.cat([Wm, (Wm[:, :, :d1].double() @ t(W1).double()).float()], dim=2)   # [K, d, d+W]
        if cmix is None:
            b1e = torch.cat([b1.new_zeros(K, dd), b1], dim=1)
        else:                                     # folded ActNorm: v = z @ Wmix + c, h1pre bias b1 + c[:d1] @ W1^T
            b1e = torch.cat([cmix, (b1.double() + (cmix[:, None, :d1].double() @ t(W1).double())[:, 0, :]).float()], dim=1)
        w1mt = (W1.double() @ t(Wm[:, :, :d1]).double()).float()                         # [K, W, d]
        mats = [(mw1, b1e), (t(W2), b2), (t(W3), b3), (W3, None), (W2, None), (w1mt, None), (t(Wm), None)]
        layer = W1.new_zeros(K, per_layer)
        for i, (M, bias) in enumerate(mats):
            mo, bo, kk, nn_ = int(offs[4 + 2 * i]), int(offs[5 + 2 * i]), int(offs[18 + 2 * i]), int(offs[19 + 2 * i])
            assert tuple(M.shape[1:]) == (kk, nn_), (i, M.shape, kk, nn_)
            layer[:, mo:mo + kk * nn_] = M.reshape(K, -1)
            if bias is not None:
                layer[:, bo:bo + nn_] = bias
        layer[:, logs_off] = logs
        plain = torch.cat([self._nf_model.q0.loc.reshape(-1), self._nf_model.q0.log_scale.reshape(-1),
                           layer.reshape(-1)]).contiguous()
        assert plain.numel() == total and off_layers == 2 * dd
        return plain

    def _umma_ws(self, n: int, dev) -> torch.Tensor:
        nbytes = int(_lib.lib().fab_umma_workspace_bytes(self.desc(), n))
        _lib.check(nbytes, "fab_umma_workspace_bytes")
        if self._uws is None or self._uws.numel() < nbytes or self._uws.device != dev:
            self._uws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        return self._uws

    # ---- Distribution surface ----------------------------------------------------------------
    @property
    def event_shape(self) -> Tuple[int, ...]:
        return self._nf_model.q0.shape

    def _device(self):
        return self._nf_model.q0.loc.device

    def sample_and_log_prob(self, shape: Tuple[int, ...]) -> Tuple[torch.Tensor, torch.Tensor]:
        assert len(shape) == 1
        eps, self._eps_override = self._eps_override, None
        if eps is None:
            eps = torch.randn((shape[0], self.dim), dtype=torch.float32, device=self._device())
        return _SampleFn.apply(self, eps, *self.parameters())

    def sample(self, shape: Tuple) -> torch.Tensor:
        return self.sample_and_log_prob(shape)[0]

    def log_prob(self, x: torch.Tensor) -> torch.Tensor:
        return _LogProbFn.apply(self, x, *self.parameters())

    # raw kernel launches (no autograd), used by the fused sampler too
    def cuda_sample(self, eps: torch.Tensor):
        eps = _lib.f32(eps).contiguous()
        n = eps.shape[0]
        x = torch.empty_like(eps)
        log_q = torch.empty(n, dtype=torch.float32, device=eps.device)
        rc = _lib.lib().fab_flow_sample_f32(self.desc(), _lib.ptr(self.blob()), _lib.ptr(eps),
                                            _lib.ptr(x), _lib.ptr(log_q), n,
                                            _lib.stream_ptr(eps.device))
        _lib.check(rc, "fab_flow_sample_f32")
        return x, log_q

    def cuda_log_prob(self, x: torch.Tensor, with_grad: bool):
        x = _lib.f32(x).contiguous()
        n = x.shape[0]
        log_q = torch.empty(n, dtype=torch.float32, device=x.device)
        grad = torch.empty_like(x) if with_grad else None
        if n > 0 and self.use_rowtile(n):
            rc = _lib.lib().fab_flow_logprob_grad_umma_f32(
                self.desc(), _lib.ptr(self.umma_blob()), _lib.ptr(x), _lib.ptr(log_q), _lib.ptr(grad),
                _lib.ptr(self._umma_ws(n, x.device)), n, _lib.stream_ptr(x.device))
            _lib.check(rc, "fab_flow_logprob_grad_umma_f32")
            return log_q, grad
        rc = _lib.lib().fab_flow_logprob_grad_f32(self.desc(), _lib.ptr(self.blob()), _lib.ptr(x),
                                                  _lib.ptr(log_q), _lib.ptr(grad), n,
                                                  _lib.stream_ptr(x.device))
        _lib.check(rc, "fab_flow_logprob_grad_f32")
        return log_q, grad

    # ---- parameter gradient (csrc/param_grad.cuh) ------------------------------------------------
    def _pg_layout(self, n: int):
        import ctypes as C
        if n not in self._pg_layouts:
            offs = (C.c_int64 * 20)()
            _lib.check(_lib.lib().fab_flow_param_grad_layout(self.desc(), n, offs), "fab_flow_param_grad_layout")
            self._pg_layouts[n] = [int(v) for v in offs]
        return self._pg_layouts[n]

    def cuda_log_prob_tape(self, x: torch.Tensor, with_grad: bool):
        """log q, (optionally) d log q / dx and the activation tape of the parameter gradient."""
        x = _lib.f32(x).contiguous()
        n = x.shape[0]
        offs = self._pg_layout(n)
        log_q = torch.empty(n, dtype=torch.float32, device=x.device)
        grad = torch.empty_like(x) if with_grad else None
        tape = torch.empty(max(offs[0], 1), dtype=torch.float32, device=x.device)
        rc = _lib.lib().fab_flow_logprob_tape_f32(self.desc(), _lib.ptr(self.blob()), _lib.ptr(x),
                                                  _lib.ptr(log_q), _lib.ptr(grad), _lib.ptr(tape), n,
                                                  _lib.stream_ptr(x.device))
        _lib.check(rc, "fab_flow_logprob_tape_f32")
        return log_q, grad, tape

    def _pg_state(self, n: int, dev):
        """Per batch size: the kernels' output / workspace buffers and the flat gradient buffer (in
        `self.parameters()` order) the chain rule writes.  The buffers are persistent so that the chain
        rule -- ~40 small torch launches -- can be replayed as one CUDA graph (`_repack`)."""
        st = self._pg_states.get(n)
        if st is not None and st["out"].device == dev:
            return st
        offs = self._pg_layout(n)
        ps = list(self.parameters())
        sizes = [p.numel() for p in ps]
        starts = [0]
        for sz in sizes:
            starts.append(starts[-1] + sz)
        st = dict(offs=offs, out=torch.empty(offs[10], dtype=torch.float32, device=dev),
                  ws=torch.empty(max(offs[17], 1), dtype=torch.float32, device=dev),
                  res=torch.zeros(starts[-1], dtype=torch.float32, device=dev),
                  slices=[(starts[i], starts[i + 1], tuple(p.shape)) for i, p in enumerate(ps)],
                  start={id(p): starts[i] for i, p in enumerate(ps)})
        if len(self._pg_states) >= 8:             # training uses one or two batch sizes; bound the rest
            old = next(iter(self._pg_states))
            del self._pg_states[old]
            self._pack_graphs.pop(f"pg{old}", None)
        self._pg_states[n] = st
        self._pack_graphs.pop(f"pg{n}", None)
        return st

    def _pg_dst(self, st, params):
        """View of the flat gradient buffer covering the same parameter of every layer, [K, *shape]."""
        K = len(params)
        o = [st["start"][id(p)] for p in params]
        shape = tuple(params[0].shape)
        step = o[1] - o[0] if K > 1 else params[0].numel()
        if any(o[k] != o[0] + k * step for k in range(K)) or any(tuple(p.shape) != shape for p in params):
            raise RuntimeError("flow layers do not share one parameter layout")
        strides, acc = [], 1
        for sz in reversed(shape):
            strides.append(acc)
            acc *= sz
        return torch.as_strided(st["res"], (K,) + shape, (step,) + tuple(reversed(strides)), o[0])

    def _pg_chain(self, st):
        """Chain rule in parameter space, batched over the layers: kernel output (`param_grad.cuh`:
        Ga, Gb, Gc, Gd, base tail) -> gradients of the module's parameters, written into st['res'].
        W = P Lf Uf with Lf = tril(L, -1) + I, Uf = triu(U, 1) + diag(sign_S exp(log_S)):
          dL = tril(P^T dW Uf^T, -1),  dU = triu(Lf^T P^T dW, 1),
          dlog_S = diag(Lf^T P^T dW) sign_S exp(log_S) + sum_i g_i    (sum(log_S) enters log q directly)."""
        offs, out, res = st["offs"], st["out"], st["res"]
        LS, oa, ob, oc, od, tail_off = offs[11], offs[12], offs[13], offs[14], offs[15], offs[16]
        K, d, W = self.n_flow_layers, self.dim, self.width
        dsc = self.desc()
        d1, p2 = dsc.d1, 2 * dsc.d2
        q0 = self._nf_model.q0
        s0 = st["start"][id(q0.loc)]
        res[s0:s0 + d].copy_(out[tail_off:tail_off + d])
        s0 = st["start"][id(q0.log_scale)]
        res[s0:s0 + d].copy_(out[tail_off + d:tail_off + 2 * d])
        if not K:
            return
        dev = out.device
        lay = out[:K * LS].view(K, LS)
        Ga = lay[:, oa:oa + (d + 1) * W].view(K, d + 1, W)
        Gb = lay[:, ob:ob + (d + 1) * d].view(K, d + 1, d)
        Gc = lay[:, oc:oc + W * (W + 1)].view(K, W, W + 1)
        Gd = lay[:, od:od + p2 * (W + 1)].view(K, p2, W + 1)
        blocks, mixes, acts = self._blocks(), self._mixes(), self._acts()
        sum_g = out[tail_off + 2 * d]
        P = torch.stack([m.P for m in mixes])
        Lf = torch.tril(torch.stack([m.L.detach() for m in mixes]), diagonal=-1) + mixes[0].eye
        sdiag = torch.stack([m.sign_S for m in mixes]) * torch.exp(torch.stack([m.log_S.detach() for m in mixes]))
        Uf = torch.triu(torch.stack([m.U.detach() for m in mixes]), diagonal=1) + torch.diag_embed(sdiag)
        Wm = P @ Lf @ Uf                                                                # the kernels' Wmix ...
        if acts:                                                                        # ... = diag(exp(-s)) W with ActNorm
            s_ = torch.stack([a.s.detach().reshape(-1) for a in acts])
            t_ = torch.stack([a.t.detach().reshape(-1) for a in acts])
            es = torch.exp(-s_)
            Wm = es[:, :, None] * Wm
        W1 = torch.stack([b.linears[0].weight.detach() for b in blocks])              # [K, W, d1]
        dM1 = Ga[:, :d, :]                                                              # [K, d, W]
        db1 = Ga[:, d, :]                                                               # [K, W]
        dW1 = dM1.transpose(1, 2) @ Wm[:, :, :d1]
        dWm = Gb[:, :d, :].clone()
        dWm[:, :, :d1] += dM1 @ W1
        if acts:
            # v = z @ Wmix + c and h1pre = z @ M1 + (b1 + c[:d1] @ W1^T) with c = -t @ Wmix
            c = -(t_[:, None, :] @ Wm)[:, 0, :]
            dW1 = dW1 + db1[:, :, None] * c[:, None, :d1]
            dc = Gb[:, d, :].clone()
            dc[:, :d1] += (db1[:, None, :] @ W1)[:, 0, :]
            dWm = dWm - t_[:, :, None] * dc[:, None, :]
            self._pg_dst(st, [a.t for a in acts]).copy_((-(dc[:, None, :] @ Wm.transpose(1, 2)))[:, 0, :].unsqueeze(1))
            self._pg_dst(st, [a.s for a in acts]).copy_((-(dWm * Wm).sum(dim=2) - sum_g).unsqueeze(1))
            dWm = es[:, :, None] * dWm                                                  # d / d (P Lf Uf)
        self._pg_dst(st, [b.linears[0].weight for b in blocks]).copy_(dW1)
        self._pg_dst(st, [b.linears[0].bias for b in blocks]).copy_(db1)
        self._pg_dst(st, [b.linears[1].weight for b in blocks]).copy_(Gc[:, :, :W])
        self._pg_dst(st, [b.linears[1].bias for b in blocks]).copy_(Gc[:, :, W])
        if self._perm_dev is None or self._perm_dev.device != dev:
            self._perm_dev = torch.cat([torch.arange(0, p2, 2, device=dev), torch.arange(1, p2, 2, device=dev)])
        perm = self._perm_dev                     # kernel rows: shifts then scales; torch rows interleave them
        self._pg_dst(st, [b.linears[2].weight for b in blocks])[:, perm, :] = Gd[:, :, :W]
        self._pg_dst(st, [b.linears[2].bias for b in blocks])[:, perm] = Gd[:, :, W]
        A = P.transpose(1, 2) @ dWm
        dUf = Lf.transpose(1, 2) @ A
        self._pg_dst(st, [m.L for m in mixes]).copy_(torch.tril(A @ Uf.transpose(1, 2), diagonal=-1))
        self._pg_dst(st, [m.U for m in mixes]).copy_(torch.triu(dUf, diagonal=1))
        self._pg_dst(st, [m.log_S for m in mixes]).copy_(torch.diagonal(dUf, dim1=1, dim2=2) * sdiag + sum_g)

    def cuda_param_grad(self, tape: torch.Tensor, g: torch.Tensor):
        """Gradients of sum_i g_i log q(x_i) for every parameter, in `self.parameters()` order."""
        n = g.shape[0]
        dev = tape.device
        st = self._pg_state(n, dev)
        g = _lib.f32(g).contiguous()
        rc = _lib.lib().fab_flow_param_grad_f32(self.desc(), _lib.ptr(self.blob()), _lib.ptr(tape), _lib.ptr(g), n,
                                                _lib.ptr(st["out"]), _lib.ptr(st["ws"]), _lib.stream_ptr(dev))
        _lib.check(rc, "fab_flow_param_grad_f32")
        with torch.no_grad():
            self._repack(f"pg{n}", lambda: self._pg_chain(st))
            flat = st["res"].clone()              # the persistent buffer is overwritten by the next call
        return [flat[a:b].view(shape) for a, b, shape in st["slices"]]

    # ---- the same maths in torch ops (GPU), differentiable w.r.t. parameters -----------------
    def _coupling_params(self, k, v1):
        l1, l2, l3 = self._blocks()[k].linears
        h = torch.relu(l1(v1))
        h = torch.relu(l2(h))
        par = l3(h)
        return par[:, 0::2], par[:, 1::2]

    def torch_log_prob(self, x: torch.Tensor) -> torch.Tensor:
        d1 = int((self.dim / 2) + 0.5)
        q0 = self._nf_model.q0
        log_q = torch.zeros(len(x), dtype=x.dtype, device=x.device)
        z = x
        if self.n_flow_layers:
            W, _, logs, c, _ = self._mixing(x.dtype)
        for k in range(self.n_flow_layers - 1, -1, -1):
            v = z @ W[k]
            if c is not None:
                v = v + c[k]
            v1, v2 = v[:, :d1], v[:, d1:]
            shift, scale = self._coupling_params(k, v1)
            z = torch.cat([v1, (v2 - shift) * torch.exp(-scale)], dim=1)
            log_q = log_q + logs[k] - scale.sum(dim=1)
        return log_q - 0.5 * self.dim * math.log(2 * math.pi) - torch.sum(
            q0.log_scale + 0.5 * ((z - q0.loc) / torch.exp(q0.log_scale)) ** 2, dim=1)

    def torch_sample(self, eps: torch.Tensor):
        d1 = int((self.dim / 2) + 0.5)
        q0 = self._nf_model.q0
        z = q0.loc + torch.exp(q0.log_scale) * eps
        log_q = -0.5 * self.dim * math.log(2 * math.pi) - torch.sum(
            q0.log_scale + 0.5 * eps ** 2, dim=1)
        if self.n_flow_layers:
            _, W_inv, logs, _, t = self._mixing(eps.dtype)
        for k in range(self.n_flow_layers):
            v1, v2 = z[:, :d1], z[:, d1:]
            shift, scale = self._coupling_params(k, v1)
            z = torch.cat([v1, v2 * torch.exp(scale) + shift], dim=1) @ W_inv[k]
            if t is not None:
                z = z + t[k]
            log_q = log_q - scale.sum(dim=1) + logs[k]
        return z, log_q


class _LogProbFn(torch.autograd.Function):
    """forward: CUDA value (+ input-gradient, + activation tape when a parameter requires grad);
    backward: dx from the saved input-gradient, dtheta from the tape (fab_flow_param_grad_f32)."""

    @staticmethod
    def forward(ctx, flow: B200RealNVP, x, *params):
        need_dx = x.requires_grad
        need_dp = any(ctx.needs_input_grad[2:])
        tape = None
        if need_dp and x.shape[0] > 0:
            log_q, grad, tape = flow.cuda_log_prob_tape(x.detach(), with_grad=need_dx)
        else:
            log_q, grad = flow.cuda_log_prob(x.detach(), with_grad=need_dx)
        ctx.flow = flow
        ctx.n_params = len(params)
        ctx.save_for_backward(grad if grad is not None else x.new_empty(0),
                              tape if tape is not None else x.new_empty(0))
        ctx.has_dx = need_dx
        ctx.has_tape = tape is not None
        return log_q

    @staticmethod
    def backward(ctx, g):
        grad, tape = ctx.saved_tensors
        flow = ctx.flow
        dx = g[:, None] * grad if (ctx.has_dx and ctx.needs_input_grad[1]) else None
        dparams = [None] * ctx.n_params
        if any(ctx.needs_input_grad[2:]) and ctx.has_tape:
            with torch.no_grad():
                gs = flow.cuda_param_grad(tape, g)
            dparams = [gr if need else None for gr, need in zip(gs, ctx.needs_input_grad[2:])]
        return (None, dx, *dparams)


class _SampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flow: B200RealNVP, eps, *params):
        x, log_q = flow.cuda_sample(eps)
        ctx.flow = flow
        ctx.n_params = len(params)
        ctx.save_for_backward(eps)
        return x, log_q

    @staticmethod
    def backward(ctx, gx, glq):
        (eps,) = ctx.saved_tensors
        flow = ctx.flow
        dparams = [None] * ctx.n_params
        if any(ctx.needs_input_grad[2:]):
            with torch.enable_grad():
                params = [p for p in flow.parameters()]
                wanted = [p for p, need in zip(params, ctx.needs_input_grad[2:]) if need]
                x, lq = flow.torch_sample(eps)
                outs, gouts = [], []
                if gx is not None:
                    outs.append(x); gouts.append(gx)
                if glq is not None:
                    outs.append(lq); gouts.append(glq)
                gs = torch.autograd.grad(outs, wanted, grad_outputs=gouts, allow_unused=True)
            it = iter(gs)
            dparams = [next(it) if need else None for need in ctx.needs_input_grad[2:]]
        return (None, None, *dparams)


def make_wrapped_b200_realnvp(dim: int, n_flow_layers: int = 5, layer_nodes_per_dim: int = 10,
                              act_norm: bool = False) -> B200RealNVP:
    """Same signature as `make_wrapped_normflow_realnvp` (make_normflow_model.py:82-96)."""
    return B200RealNVP(dim, n_flow_layers, layer_nodes_per_dim, act_norm)
