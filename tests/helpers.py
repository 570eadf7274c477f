"""Shared builders for the parity tests: matched (oracle fp64, oracle fp32, product CUDA) objects."""
import copy

import torch

import fab_torch_b200 as fb
from oracle.realnvp import OracleRealNVP, randomize_last_layers
from oracle.targets import OracleManyWell, OracleGMM, to_double


def make_flows(dim, n_layers, nodes_per_dim, seed=0, last_std=0.05, perturb_base=True, device="cuda",
               act_norm=False):
    """Returns (oracle fp64, oracle fp32, product on `device`) sharing the same fp32 weights.  With
    `act_norm` the ActNorm layers are initialised from data by the oracle's constructor (500 samples)
    and then moved off their init so that shifts and scales are not near the identity."""
    torch.manual_seed(seed)
    fo = OracleRealNVP(dim, n_layers, nodes_per_dim, act_norm=act_norm)
    if n_layers:
        randomize_last_layers(fo, last_std, seed=seed + 1)
    if act_norm:
        g = torch.Generator().manual_seed(seed + 3)
        with torch.no_grad():
            for name, p in fo.named_parameters():
                if name.endswith(".s"):
                    p.add_(torch.randn(p.shape, generator=g) * 0.05)
                elif name.endswith(".t"):
                    p.add_(torch.randn(p.shape, generator=g) * 0.15)
    if perturb_base:
        g = torch.Generator().manual_seed(seed + 2)
        with torch.no_grad():
            fo._nf_model.q0.loc.add_(torch.randn(1, dim, generator=g) * 0.2)
            fo._nf_model.q0.log_scale.add_(torch.randn(1, dim, generator=g) * 0.1)
    fo64 = copy.deepcopy(fo).double()
    fp = fb.B200RealNVP(dim, n_layers, nodes_per_dim, act_norm=act_norm)
    fp.load_state_dict(fo.state_dict())
    if device is not None:
        fp = fp.to(device)
    return fo64, fo, fp


def make_manywell(dim, device="cuda"):
    return OracleManyWell(dim), fb.ManyWellEnergy(dim, use_gpu=(device == "cuda"))


def make_gmm(dim, n_mixes, loc_scaling, log_var_scaling=0.1, seed=0, device="cuda"):
    torch.manual_seed(seed)
    to = OracleGMM(dim, n_mixes, loc_scaling, log_var_scaling)
    torch.manual_seed(seed)
    tp = fb.GMM(dim, n_mixes, loc_scaling, log_var_scaling, use_gpu=(device == "cuda"))
    assert torch.equal(to.locs, tp.locs.cpu())
    to64 = to_double(copy.deepcopy(to))
    return to64, to, tp


def make_aldp(dim=60, seed=0, device="cuda"):
    """(oracle fp64, oracle fp32, product) ALDP-surrogate targets sharing the same tables."""
    from oracle.targets import OracleAldpSurrogate, aldp_surrogate_tables
    to = OracleAldpSurrogate(dim, seed)
    to64 = OracleAldpSurrogate(dim, seed).double()
    tp = fb.AldpSurrogateEnergy(dim, seed, tables=aldp_surrogate_tables(dim, seed),
                                use_gpu=(device == "cuda"))
    return to64, to, tp


def rel_err(a, b):
    """max |a-b| / max(1, |b|) elementwise -> scalar (a: test value, b: truth)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return ((a - b).abs() / b.abs().clamp_min(1.0)).max().item()


def row_rel_err(a, b):
    """per-row max |a-b| / max(1, |b|)  (a: test value, b: truth); 1-D inputs count as rows of 1."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    e = (a - b).abs() / b.abs().clamp_min(1.0)
    return e.reshape(e.shape[0], -1).max(dim=1).values if e.ndim else e.reshape(1)


def assert_parity(gpu, truth, cpu32=None, name="", floor=1e-5, factor=4.0, mask=None, outlier_frac=0.0,
                  outlier_cap=100.0):
    """The parity bar (BASELINE.json: 'within 1e-5 relative fp32'): the CUDA result must be within
    `floor` (relative) of the fp64 ground truth, or -- where rounding is amplified by the
    computation itself (chaotic leapfrog, exp of large scales) -- no worse than `factor` x the
    error the reference's own fp32 CPU arithmetic makes on the same inputs.

    `outlier_frac`: the flow's ReLUs make d log q / dx (and everything integrated from it)
    DISCONTINUOUS: a hidden unit whose pre-activation is within rounding of zero takes the other
    branch in any two fp32 implementations (the reference's own CPU run does it too, on other
    particles).  Such rows are allowed for at most this fraction of the particles (min. 1 row when
    > 0) and must still be within `outlier_cap` x the bar (None: no cap -- a flipped ReLU changes the
    gradient by an amount that has nothing to do with rounding); every other row must meet the bar."""
    pick = (lambda t: t.detach().cpu()[mask]) if mask is not None else (lambda t: t.detach().cpu())
    g, t = pick(gpu), pick(truth)
    fin = torch.isfinite(t.double())
    assert torch.equal(torch.isfinite(g.double()), fin), f"{name}: finiteness pattern differs"
    err = rel_err(g[fin], t[fin]) if fin.any() else 0.0
    bar = floor
    err32 = None
    if cpu32 is not None:
        c = pick(cpu32)
        fin32 = fin & torch.isfinite(c.double())
        err32 = rel_err(c[fin32], t[fin32]) if fin32.any() else 0.0
        bar = max(floor, factor * err32)
    if err > bar and outlier_frac > 0 and g.ndim >= 1 and g.shape[0] > 1:
        rows = row_rel_err(torch.where(fin, g.double(), torch.zeros_like(g.double())),
                           torch.where(fin, t.double(), torch.zeros_like(t.double())))
        allowed = max(1, int(outlier_frac * rows.numel()))
        bad = rows > bar
        assert int(bad.sum()) <= allowed and (outlier_cap is None or err <= outlier_cap * bar), (
            f"{name}: {int(bad.sum())} rows above the bar {bar:.3e} (allowed {allowed}), worst {err:.3e}"
            + (f" (cpu fp32 reference err {err32:.3e})" if err32 is not None else ""))
        return float(rows[~bad].max()) if (~bad).any() else 0.0, err32
    assert err <= bar, (f"{name}: cuda rel err {err:.3e} > bar {bar:.3e}"
                        + (f" (cpu fp32 reference err {err32:.3e})" if err32 is not None else ""))
    return err, err32
