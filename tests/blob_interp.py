"""Test helper: interpret a packed flow blob (layout of include/fab_b200.h) with numpy, operand by
operand, the way the kernels consume it (MMA fragment order, biases as separate vectors, mixing
matrices merged into the neighbouring MLP GEMMs).  Validates descriptor offsets and the host-side
packing without a GPU."""
import numpy as np


def _r(v, m):
    return (v + m - 1) // m * m


def unpack_frag(blob, off, K16, N8):
    """fragment order float4 [K16/16][N8/8][32 lanes] -> dense M[K16][N8] (mma_gemm.cuh)."""
    KT2, NT = K16 // 16, N8 // 8
    a = blob[off:off + KT2 * NT * 128].reshape(KT2, NT, 8, 4, 2, 2)   # [kp][nt][g][t][h][half]
    M = np.zeros((K16, N8), dtype=blob.dtype)
    for kp in range(KT2):
        for h in range(2):
            for half in range(2):
                for t in range(4):
                    k = 16 * kp + 8 * h + 4 * half + t
                    M[k] = a[kp, :, :, t, h, half].reshape(-1)           # n = 8*nt + g
    return M


def layer_views(blob, d, k):
    base = d.off_layers + k * d.layer_stride
    dim = d.dim
    D8, D16 = _r(dim, 8), _r(dim, 16)
    D1K, P8, P16 = _r(d.d1, 16), _r(2 * d.d2, 8), _r(2 * d.d2, 16)
    W8, W16 = d.width_pad, d.width_kpad
    g = lambda off, K16, N8: unpack_frag(blob, base + off, K16, N8)
    v = lambda off, n: blob[base + off: base + off + n]
    return dict(mw1=g(d.o_mw1, D16, D8 + W8), w2=g(d.o_w2, W16, W8), w3=g(d.o_w3, W16, P8),
                w3t=g(d.o_w3t, P16, W8), w2t=g(d.o_w2t, W16, W8), w1mt=g(d.o_w1mt, W16 + D16, D8),
                w1=g(d.o_w1, D1K, W8), mix_inv=g(d.o_mix_inv, D16, D8),
                b1=v(d.o_b1, D8 + W8), b2=v(d.o_b2, W8), b3=v(d.o_b3, P8), logs=blob[base + d.o_logs],
                b1s=v(d.o_b1s, W8), tmix=v(d.o_tmix, D8))


def _pad(a, width):
    out = np.zeros((a.shape[0], width))
    out[:, :a.shape[1]] = a
    return out


def interp_log_prob_and_grad(blob, d, x):
    """Mirrors flow_inverse + flow_backward of fab_torch_b200/csrc/flow_tile.cuh in float64."""
    blob = np.asarray(blob, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    dim, d1, d2 = d.dim, d.d1, d.d2
    D8, D16, P16, W8, W16 = _r(dim, 8), _r(dim, 16), _r(2 * d2, 16), d.width_pad, d.width_kpad
    z = x.copy()
    ld = np.zeros(len(x))
    saved = []
    for k in range(d.n_layers - 1, -1, -1):
        L = layer_views(blob, d, k)
        out = _pad(z, D16) @ L["mw1"] + L["b1"]              # [v | h1pre]
        v, a1 = out[:, :dim], out[:, D8:D8 + W8]
        h1 = np.maximum(a1, 0)
        a2 = _pad(h1, W16) @ L["w2"] + L["b2"]
        h2 = np.maximum(a2, 0)
        par = _pad(h2, W16) @ L["w3"] + L["b3"]
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
        gpar = np.zeros((len(x), P16))
        gpar[:, :d2] = -gv2
        gpar[:, d2:2 * d2] = -g2 * y2 - 1.0
        gh2 = (gpar @ L["w3t"]) * m2
        gh1 = (_pad(gh2, W16) @ L["w2t"]) * m1
        gv = np.zeros((len(x), D16))
        gv[:, :d1] = g[:, :d1]
        gv[:, d1:dim] = gv2
        g = (np.concatenate([_pad(gh1, W16), gv], 1) @ L["w1mt"])[:, :dim]
    return lq, g


def interp_sample(blob, d, eps):
    blob = np.asarray(blob, dtype=np.float64)
    eps = np.asarray(eps, dtype=np.float64)
    dim, d1, d2 = d.dim, d.d1, d.d2
    D8, D16, D1K, W16 = _r(dim, 8), _r(dim, 16), _r(d1, 16), d.width_kpad
    loc = blob[d.off_base_loc: d.off_base_loc + dim]
    ls = blob[d.off_base_log_scale: d.off_base_log_scale + dim]
    z = loc + np.exp(ls) * eps
    lq = -0.5 * dim * np.log(2 * np.pi) - (ls + 0.5 * eps * eps).sum(1)
    for k in range(d.n_layers):
        L = layer_views(blob, d, k)
        # the kernel feeds rows [0, D1K) of the z operand: rows >= d1 meet zero weight rows
        h1 = np.maximum(_pad(z, max(D16, D1K))[:, :D1K] @ L["w1"] + L["b1s"], 0)
        h2 = np.maximum(_pad(h1, W16) @ L["w2"] + L["b2"], 0)
        par = _pad(h2, W16) @ L["w3"] + L["b3"]
        shift, scale = par[:, :d2], par[:, d2:2 * d2]
        y = np.zeros((len(eps), D16))
        y[:, :d1] = z[:, :d1]
        y[:, d1:dim] = z[:, d1:] * np.exp(scale) + shift
        z = (y @ L["mix_inv"] + L["tmix"])[:, :dim]
        lq += L["logs"] - scale.sum(1)
    return z, lq
