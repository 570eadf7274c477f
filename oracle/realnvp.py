"""Oracle restatement of the `normflows` RealNVP used by the reference.  TEST INFRASTRUCTURE.

PARITY UNPINNED: `normflows` (reference `requirements.txt:3`, no version pin) is
not vendored and not installed; this file restates the published algorithm of
normflows 1.x (VincentStimper/normalizing-flows) for exactly the pieces the
reference instantiates:

* `nf.nets.MLP([d1, W, W, 2*d2], init_zeros=True)`  <- experiments/make_flow/make_normflow_model.py:22
* `nf.flows.AffineCouplingBlock(param_map, scale_map="exp")`           <- :24
* `nf.flows.InvertibleAffine(dim)` (LU-parameterised)                   <- :26
* `nf.flows.ActNorm(dim)` (optional, all shipped configs disable it)    <- :29
* `nf.distributions.base.DiagGaussian(dim)`                             <- :88
* `nf.NormalizingFlow(base, flows)`, `.sample(n)`, `.log_prob(x)`       <- :92, fab/wrappers/normflows.py:18,24

Module/attribute names follow normflows so that `state_dict()` keys look like a
reference checkpoint (`_nf_model.q0.loc`, `_nf_model.flows.0.flows.1.param_map.net.0.weight`,
`_nf_model.flows.1.{P,L,U,log_S,sign_S,eye}`), cf. fab/core.py:222-260.
"""
import math
from typing import List, Tuple

import torch
import torch.nn as nn


class DiagGaussianBase(nn.Module):
    """Diagonal Gaussian with trainable `loc`/`log_scale` of shape [1, d]."""

    def __init__(self, dim: int):
        super().__init__()
        self.shape = (dim,)
        self.d = dim
        self.loc = nn.Parameter(torch.zeros(1, dim))
        self.log_scale = nn.Parameter(torch.zeros(1, dim))

    def forward(self, num_samples: int = 1, eps: torch.Tensor = None):
        if eps is None:
            eps = torch.randn((num_samples,) + self.shape, dtype=self.loc.dtype,
                              device=self.loc.device)
        z = self.loc + torch.exp(self.log_scale) * eps
        log_p = -0.5 * self.d * math.log(2 * math.pi) - torch.sum(
            self.log_scale + 0.5 * torch.pow(eps, 2), dim=1)
        return z, log_p

    def log_prob(self, z: torch.Tensor) -> torch.Tensor:
        return -0.5 * self.d * math.log(2 * math.pi) - torch.sum(
            self.log_scale + 0.5 * torch.pow((z - self.loc) / torch.exp(self.log_scale), 2),
            dim=1)


class ParamMLP(nn.Module):
    """Linear -> LeakyReLU(0.0) -> ... -> Linear, stored as `self.net` (Sequential)."""

    def __init__(self, layers: List[int], init_zeros: bool = True):
        super().__init__()
        mods = []
        for k in range(len(layers) - 2):
            mods.append(nn.Linear(layers[k], layers[k + 1]))
            mods.append(nn.LeakyReLU(0.0))
        mods.append(nn.Linear(layers[-2], layers[-1]))
        if init_zeros:
            nn.init.zeros_(mods[-1].weight)
            nn.init.zeros_(mods[-1].bias)
        self.net = nn.Sequential(*mods)

    def forward(self, x):
        return self.net(x)


class _Split(nn.Module):
    """`z.chunk(2, dim=1)`: the first chunk is the wider one when d is odd."""

    def forward(self, z):
        z1, z2 = z.chunk(2, dim=1)
        return [z1, z2], 0

    def inverse(self, zs):
        return torch.cat(zs, 1), 0


class _Merge(_Split):
    def forward(self, zs):
        return super().inverse(zs)

    def inverse(self, z):
        return super().forward(z)


class AffineCoupling(nn.Module):
    """shift = param[:, 0::2], scale = param[:, 1::2]; scale map is exp."""

    def __init__(self, param_map: nn.Module):
        super().__init__()
        self.add_module("param_map", param_map)

    def forward(self, zs):
        z1, z2 = zs
        param = self.param_map(z1)
        shift = param[:, 0::2]
        scale_ = param[:, 1::2]
        z2 = z2 * torch.exp(scale_) + shift
        return [z1, z2], torch.sum(scale_, dim=1)

    def inverse(self, zs):
        z1, z2 = zs
        param = self.param_map(z1)
        shift = param[:, 0::2]
        scale_ = param[:, 1::2]
        z2 = (z2 - shift) * torch.exp(-scale_)
        return [z1, z2], -torch.sum(scale_, dim=1)


class AffineCouplingBlock(nn.Module):
    def __init__(self, param_map: nn.Module):
        super().__init__()
        self.flows = nn.ModuleList([_Split(), AffineCoupling(param_map), _Merge()])

    def forward(self, z):
        log_det_tot = torch.zeros(z.shape[0], dtype=z.dtype, device=z.device)
        for f in self.flows:
            z, ld = f(z)
            log_det_tot = log_det_tot + ld
        return z, log_det_tot

    def inverse(self, z):
        log_det_tot = torch.zeros(z.shape[0], dtype=z.dtype, device=z.device)
        for f in reversed(self.flows):
            z, ld = f.inverse(z)
            log_det_tot = log_det_tot + ld
        return z, log_det_tot


class InvertibleAffine(nn.Module):
    """LU-parameterised d x d mixing; forward multiplies by W^-1, inverse by W."""

    def __init__(self, dim: int):
        super().__init__()
        Q, _ = torch.linalg.qr(torch.randn(dim, dim))
        P, L, U = torch.lu_unpack(*Q.lu())
        self.register_buffer("P", P)
        self.L = nn.Parameter(L)
        S = U.diag()
        self.register_buffer("sign_S", torch.sign(S))
        self.log_S = nn.Parameter(torch.log(torch.abs(S)))
        self.U = nn.Parameter(torch.triu(U, diagonal=1))
        self.register_buffer("eye", torch.diag(torch.ones(dim)))

    def assemble(self, inverse: bool = False) -> torch.Tensor:
        L = torch.tril(self.L, diagonal=-1) + self.eye
        U = torch.triu(self.U, diagonal=1) + torch.diag(self.sign_S * torch.exp(self.log_S))
        if not inverse:
            return self.P @ L @ U
        if self.log_S.dtype == torch.float64:
            L_inv, U_inv = torch.inverse(L), torch.inverse(U)
        else:
            L_inv = torch.inverse(L.double()).type(self.log_S.dtype)
            U_inv = torch.inverse(U.double()).type(self.log_S.dtype)
        return U_inv @ L_inv @ self.P.t()

    def forward(self, z):
        return z @ self.assemble(inverse=True), -torch.sum(self.log_S)

    def inverse(self, z):
        return z @ self.assemble(inverse=False), torch.sum(self.log_S)


class ActNorm(nn.Module):
    """z*exp(s)+t with data-dependent init on the first forward/inverse call."""

    def __init__(self, dim: int):
        super().__init__()
        self.s = nn.Parameter(torch.zeros(1, dim))
        self.t = nn.Parameter(torch.zeros(1, dim))
        self.register_buffer("data_dep_init_done", torch.tensor(0.0))

    def forward(self, z):
        if not self.data_dep_init_done > 0.0:
            s_init = -torch.log(z.std(dim=0, keepdim=True) + 1e-6)
            self.s.data = s_init.data
            self.t.data = (-z.mean(dim=0, keepdim=True) * torch.exp(self.s)).data
            self.data_dep_init_done = torch.tensor(1.0)
        return z * torch.exp(self.s) + self.t, torch.sum(self.s)

    def inverse(self, z):
        if not self.data_dep_init_done > 0.0:
            s_init = torch.log(z.std(dim=0, keepdim=True) + 1e-6)
            self.s.data = s_init.data
            self.t.data = z.mean(dim=0, keepdim=True).data
            self.data_dep_init_done = torch.tensor(1.0)
        return (z - self.t) * torch.exp(-self.s), -torch.sum(self.s)


class FlowModel(nn.Module):
    """`nf.NormalizingFlow`: q0 + ordered list of flows."""

    def __init__(self, q0: nn.Module, flows: List[nn.Module]):
        super().__init__()
        self.q0 = q0
        self.flows = nn.ModuleList(flows)

    def sample(self, num_samples: int = 1, eps: torch.Tensor = None):
        z, log_q = self.q0(num_samples, eps=eps)
        for f in self.flows:
            z, log_det = f(z)
            log_q = log_q - log_det
        return z, log_q

    def log_prob(self, x: torch.Tensor) -> torch.Tensor:
        log_q = torch.zeros(len(x), dtype=x.dtype, device=x.device)
        z = x
        for i in range(len(self.flows) - 1, -1, -1):
            z, log_det = self.flows[i].inverse(z)
            log_q = log_q + log_det
        return log_q + self.q0.log_prob(z)


def build_flow_list(dim: int, n_flow_layers: int, layer_nodes_per_dim: int, act_norm: bool):
    """experiments/make_flow/make_normflow_model.py:11-30."""
    flows = []
    width = dim * layer_nodes_per_dim
    for _ in range(n_flow_layers):
        d1 = int((dim / 2) + 0.5)
        flows.append(AffineCouplingBlock(ParamMLP([d1, width, width, 2 * (dim - d1)],
                                                  init_zeros=True)))
        flows.append(InvertibleAffine(dim))
        if act_norm:
            flows.append(ActNorm(dim))
    return flows


class OracleRealNVP(nn.Module):
    """Surface of `WrappedNormFlowModel` (fab/wrappers/normflows.py:8-31)."""

    def __init__(self, dim: int, n_flow_layers: int = 5, layer_nodes_per_dim: int = 10,
                 act_norm: bool = False):
        super().__init__()
        self.dim = dim
        self._nf_model = FlowModel(DiagGaussianBase(dim),
                                   build_flow_list(dim, n_flow_layers, layer_nodes_per_dim,
                                                   act_norm))
        if act_norm:
            self.sample((500,))  # make_normflow_model.py:94-95

    # optional noise injection: when set, the next `sample_and_log_prob` uses it as eps
    _eps_override = None

    def sample_and_log_prob(self, shape: Tuple[int, ...]):
        assert len(shape) == 1
        eps, self._eps_override = self._eps_override, None
        return self._nf_model.sample(shape[0], eps=eps)

    def sample(self, shape: Tuple[int, ...]):
        return self.sample_and_log_prob(shape)[0]

    def log_prob(self, x: torch.Tensor) -> torch.Tensor:
        return self._nf_model.log_prob(x)

    @property
    def event_shape(self) -> Tuple[int, ...]:
        return self._nf_model.q0.shape


def randomize_last_layers(flow: nn.Module, std: float = 0.01, seed: int = 1) -> None:
    """Benchmark/test init: zero-initialised last MLP layers make every
    coupling the identity and would hide bugs, so draw them ~N(0, std^2)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in flow.named_parameters():
            if ".param_map.net." in name:
                idx = int(name.split(".param_map.net.")[1].split(".")[0])
                # last Linear has the highest index in its Sequential
                block = name.split(".param_map.net.")[0]
                last = max(int(n.split(".param_map.net.")[1].split(".")[0])
                           for n, _ in flow.named_parameters()
                           if n.startswith(block + ".param_map.net."))
                if idx == last:
                    p.copy_(torch.randn(p.shape, generator=g, dtype=torch.float32).to(p.dtype) * std)


def analytic_log_prob_and_grad(flow: OracleRealNVP, x: torch.Tensor):
    """Analytic value and input-gradient of log_prob (last paragraph),
    used to cross-check the CUDA backward against autograd.  Supports act_norm=False."""
    nf_model = flow._nf_model
    blocks = []
    fl = list(nf_model.flows)
    assert len(fl) % 2 == 0
    for k in range(len(fl) // 2):
        blocks.append((fl[2 * k], fl[2 * k + 1]))
    d = x.shape[1]
    d1 = int((d / 2) + 0.5)
    with torch.no_grad():
        z = x
        log_q = torch.zeros(len(x), dtype=x.dtype)
        saved = []
        for coup, mix in reversed(blocks):
            Wm = mix.assemble(False)
            v = z @ Wm
            log_q = log_q + torch.sum(mix.log_S)
            v1, v2 = v[:, :d1], v[:, d1:]
            net = coup.flows[1].param_map.net
            a1 = net[0](v1)
            h1 = torch.relu(a1)
            a2 = net[2](h1)
            h2 = torch.relu(a2)
            param = net[4](h2)
            shift, scale = param[:, 0::2], param[:, 1::2]
            y2 = (v2 - shift) * torch.exp(-scale)
            log_q = log_q - scale.sum(1)
            saved.append((Wm, net, a1 > 0, a2 > 0, scale, y2))
            z = torch.cat([v1, y2], 1)
        q0 = nf_model.q0
        log_q = log_q + q0.log_prob(z)
        g = -(z - q0.loc) / torch.exp(2 * q0.log_scale)
        for Wm, net, m1, m2, scale, y2 in reversed(saved):
            g1, g2 = g[:, :d1], g[:, d1:]
            gv2 = g2 * torch.exp(-scale)
            gshift = -gv2
            gscale = -g2 * y2 - 1.0
            gparam = torch.empty(len(x), 2 * (d - d1), dtype=x.dtype)
            gparam[:, 0::2] = gshift
            gparam[:, 1::2] = gscale
            gh2 = (gparam @ net[4].weight) * m2
            gh1 = (gh2 @ net[2].weight) * m1
            gv1 = g1 + gh1 @ net[0].weight
            g = torch.cat([gv1, gv2], 1) @ Wm.t()
    return log_q, g
