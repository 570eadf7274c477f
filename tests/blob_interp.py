"""Test helper: interpret a packed flow blob (layout of include/fab_b200.h) with numpy, operand by
operand, the way the kernels consume it.  Validates descriptor offsets and the host-side packing
without a GPU."""
import numpy as np


def _r4(v):
    return (v + 3) // 4 * 4


def unpack_operand(blob, off, K, N):
    """packed float4 [K4][NP] -> dense M[K][N]."""
    K4, NP = (K + 3) // 4, _r4(N)
    a = blob[off:off + K4 * NP * 4].reshape(K4, NP, 4)
    return a.transpose(0, 2, 1).reshape(K4 * 4, NP)[:K, :N]


def layer_views(blob, d, k):
    base = d.off_layers + k * d.layer_stride
    WP, d1, d2, dim = d.width_pad, d.d1, d.d2, d.dim
    P2 = _r4(2 * d2)
    g = lambda off, K, N: unpack_operand(blob, base + off, K, N)
    return dict(
        mix=g(d.o_mix, dim, dim), mix_t=g(d.o_mix_t, dim, dim), mix_inv=g(d.o_mix_inv, dim, dim),
        w1=g(d.o_w1, d1, WP), w2=g(d.o_w2, WP, WP), w3=g(d.o_w3, WP, 2 * d2),
        w3t=g(d.o_w3t, 2 * d2, WP), w2t=g(d.o_w2t, WP, WP), w1t=g(d.o_w1t, WP, d1),
        b1=blob[base + d.o_b1: base + d.o_b1 + WP], b2=blob[base + d.o_b2: base + d.o_b2 + WP],
        b3=blob[base + d.o_b3: base + d.o_b3 + 2 * d2], logs=blob[base + d.o_logs])


def interp_log_prob_and_grad(blob, d, x):
    """Mirrors flow_inverse + flow_backward of fab_torch_b200/csrc/flow_tile.cuh in float64."""
    blob = np.asarray(blob, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    dim, d1, d2 = d.dim, d.d1, d.d2
    z = x.copy()
    ld = np.zeros(len(x))
    saved = []
    for k in range(d.n_layers - 1, -1, -1):
        L = layer_views(blob, d, k)
        v = z @ L["mix"]
        a1 = v[:, :d1] @ L["w1"] + L["b1"]; h1 = np.maximum(a1, 0)
        a2 = h1 @ L["w2"] + L["b2"]; h2 = np.maximum(a2, 0)
        par = h2 @ L["w3"] + L["b3"]
        shift, scale = par[:, :d2], par[:, d2:]
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
        gpar = np.concatenate([-gv2, -g2 * y2 - 1.0], 1)
        gh2 = (gpar @ L["w3t"]) * m2
        gh1 = (gh2 @ L["w2t"]) * m1
        gv1 = g[:, :d1] + gh1 @ L["w1t"]
        g = np.concatenate([gv1, gv2], 1) @ L["mix_t"]
    return lq, g


def interp_sample(blob, d, eps):
    blob = np.asarray(blob, dtype=np.float64)
    eps = np.asarray(eps, dtype=np.float64)
    dim, d1, d2 = d.dim, d.d1, d.d2
    loc = blob[d.off_base_loc: d.off_base_loc + dim]
    ls = blob[d.off_base_log_scale: d.off_base_log_scale + dim]
    z = loc + np.exp(ls) * eps
    lq = -0.5 * dim * np.log(2 * np.pi) - (ls + 0.5 * eps * eps).sum(1)
    for k in range(d.n_layers):
        L = layer_views(blob, d, k)
        h1 = np.maximum(z[:, :d1] @ L["w1"] + L["b1"], 0)
        h2 = np.maximum(h1 @ L["w2"] + L["b2"], 0)
        par = h2 @ L["w3"] + L["b3"]
        shift, scale = par[:, :d2], par[:, d2:]
        z = np.concatenate([z[:, :d1], z[:, d1:] * np.exp(scale) + shift], 1) @ L["mix_inv"]
        lq += L["logs"] - scale.sum(1)
    return z, lq
