"""Test helper: interpret a packed flow blob (layout of include/fab_b200.h) with numpy, operand by
operand, the way the kernels consume it (biases folded in as an extra K row, mixing matrices merged
into the neighbouring MLP GEMMs).  Validates descriptor offsets and the host-side packing without
a GPU."""
import numpy as np


def _r4(v):
    return (v + 3) // 4 * 4


def unpack_operand(blob, off, K, N):
    """packed float4 [K4][NP] -> dense M[K4*4][NP]."""
    K4, NP = (K + 3) // 4, _r4(N)
    a = blob[off:off + K4 * NP * 4].reshape(K4, NP, 4)
    return a.transpose(0, 2, 1).reshape(K4 * 4, NP)


def layer_views(blob, d, k):
    base = d.off_layers + k * d.layer_stride
    WP, dim = d.width_pad, d.dim
    DP, D1P, P2 = _r4(dim), _r4(d.d1), _r4(2 * d.d2)
    g = lambda off, K, N: unpack_operand(blob, base + off, K, N)
    return dict(mw1=g(d.o_mw1, DP + 4, DP + WP), w2=g(d.o_w2, WP + 4, WP), w3=g(d.o_w3, WP + 4, P2),
                w3t=g(d.o_w3t, P2, WP), w2t=g(d.o_w2t, WP, WP), w1mt=g(d.o_w1mt, WP + DP, DP),
                w1=g(d.o_w1, D1P + 4, WP), mix_inv=g(d.o_mix_inv, DP, DP), logs=blob[base + d.o_logs])


def _ext(a, width):
    """[a | 0-pad to `width` | 1 0 0 0]"""
    out = np.zeros((a.shape[0], width + 4))
    out[:, :a.shape[1]] = a
    out[:, width] = 1.0
    return out


def interp_log_prob_and_grad(blob, d, x):
    """Mirrors flow_inverse + flow_backward of fab_torch_b200/csrc/flow_tile.cuh in float64."""
    blob = np.asarray(blob, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    dim, d1, d2, WP = d.dim, d.d1, d.d2, d.width_pad
    DP, P2 = _r4(dim), _r4(2 * d2)
    z = x.copy()
    ld = np.zeros(len(x))
    saved = []
    for k in range(d.n_layers - 1, -1, -1):
        L = layer_views(blob, d, k)
        out = _ext(z, DP) @ L["mw1"]                      # [v | h1pre]
        v, a1 = out[:, :dim], out[:, DP:DP + WP]
        h1 = np.maximum(a1, 0)
        a2 = _ext(h1, WP) @ L["w2"]
        h2 = np.maximum(a2, 0)
        par = _ext(h2, WP) @ L["w3"]
        shift, scale = par[:, :d2], par[:, d2:2 * d2]
        es = np.exp(-scale)
        y2 = (v[:, d1:] - shift) * es
        z = np.concatenate([v[:, :d1], y2], 1)
        ld += L["logs"] - scale.sum(1)
        saved.append((L, a1 > 0, a2 > 0, es, y2))
    loc = blob[d.off_base_loc: d.off_base_loc + dim]
    ls = blob[d.off_base_log_scale: d.off_base_log_scale + dim]
    u = (z - loc) * np.exp(-ls)
    lq = ld - 0.5 * dim * np.log(2 * np.pi) - (ls + 0.5 * u * u).sum(1)
    g = -u * np.exp(-ls)
    for L, m1, m2, es, y2 in reversed(saved):
        g2 = g[:, d1:]
        gv2 = g2 * es
        gpar = np.zeros((len(x), P2))
        gpar[:, :d2] = -gv2
        gpar[:, d2:2 * d2] = -g2 * y2 - 1.0
        gh2 = (gpar @ L["w3t"]) * m2
        gh1 = (gh2 @ L["w2t"]) * m1
        gv = np.zeros((len(x), DP))
        gv[:, :d1] = g[:, :d1]
        gv[:, d1:dim] = gv2
        g = (np.concatenate([gh1, gv], 1) @ L["w1mt"])[:, :dim]
    return lq, g


def interp_sample(blob, d, eps):
    blob = np.asarray(blob, dtype=np.float64)
    eps = np.asarray(eps, dtype=np.float64)
    dim, d1, d2, WP = d.dim, d.d1, d.d2, d.width_pad
    DP, D1P = _r4(dim), _r4(d1)
    loc = blob[d.off_base_loc: d.off_base_loc + dim]
    ls = blob[d.off_base_log_scale: d.off_base_log_scale + dim]
    z = loc + np.exp(ls) * eps
    lq = -0.5 * dim * np.log(2 * np.pi) - (ls + 0.5 * eps * eps).sum(1)
    for k in range(d.n_layers):
        L = layer_views(blob, d, k)
        h1 = np.maximum(_ext(z[:, :d1], D1P) @ L["w1"], 0)
        h2 = np.maximum(_ext(h1, WP) @ L["w2"], 0)
        par = _ext(h2, WP) @ L["w3"]
        shift, scale = par[:, :d2], par[:, d2:2 * d2]
        y = np.zeros((len(eps), DP))
        y[:, :d1] = z[:, :d1]
        y[:, d1:dim] = z[:, d1:] * np.exp(scale) + shift
        z = (y @ L["mix_inv"])[:, :dim]
        lq += L["logs"] - scale.sum(1)
    return z, lq
