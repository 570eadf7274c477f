"""CPU: libfab_b200.so loads, exports every symbol include/fab_b200.h declares, and its host-only
entry points behave (no kernel launches here)."""
import ctypes as C
import os
import re

import pytest

from fab_torch_b200 import _lib
from fab_torch_b200.csrc import build as _build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    _build.build()          # no-op when the in-tree .so is up to date
    return _lib.lib()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "fab_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fab_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(L):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert n in _lib.SIGNATURES, f"{n} declared in the header but not bound in _lib.py"
        assert getattr(L, n) is not None
    assert set(_lib.SIGNATURES) == set(names)


def test_struct_sizes_match_header(L):
    # natural C alignment of the header structs
    assert C.sizeof(_lib.FlowDesc) == 8 * 4 + 19 * 8      # 7 int32 + pad, 19 int64
    assert C.sizeof(_lib.Gamma) == 16
    assert C.sizeof(_lib.PointPtrs) == 40
    assert C.sizeof(_lib.TargetDesc) == 8 * 4 + 3 * 8
    assert C.sizeof(_lib.HmcArgs) == 6 * 4 + 16 + 4 + 16 + 16 + 4
    assert C.sizeof(_lib.MetropolisArgs) == 4 * 4 + 16 + 4 + 16 + 16 + 4


def test_flow_desc_init(L):
    d = _lib.FlowDesc()
    n = L.fab_flow_desc_init(d, 32, 320, 10)
    assert n == d.total_floats > 0
    assert (d.dim, d.d1, d.d2, d.width_pad, d.width_kpad, d.n_layers) == (32, 16, 16, 320, 320, 10)
    # fragment-ordered operands (K16 x N8): o_mw1 (32x352) o_w2 (320x320) o_w3 (320x32)
    # o_w3t (32x320) o_w2t (320x320) o_w1mt (352x32) o_w1 (16x320) o_mix_inv (32x32);
    # vectors o_b1 (352) o_b2 (320) o_b3 (32) o_logs (4) o_b1s (320) o_tmix (32)
    per_layer = (32 * 352 + 320 * 320 + 320 * 32 + 32 * 320 + 320 * 320 + 352 * 32 + 16 * 320 + 32 * 32
                 + 352 + 320 + 32 + 4 + 320 + 32)
    assert d.layer_stride == per_layer
    assert d.total_floats == 64 + 10 * per_layer + 512
    assert L.fab_flow_desc_init(d, 5, 15, 2) > 0
    assert (d.d1, d.d2, d.width_pad, d.width_kpad) == (3, 2, 16, 16)
    assert L.fab_flow_desc_init(d, 1, 10, 1) < 0
    assert b"dim>=2" in L.fab_last_error()


def test_bad_arguments_are_rejected_without_touching_the_gpu(L):
    d = _lib.FlowDesc()
    L.fab_flow_desc_init(d, 4, 8, 1)
    assert L.fab_flow_sample_f32(d, None, None, None, None, 10, None) == -1
    assert L.fab_flow_logprob_grad_f32(d, None, None, None, None, 10, None) == -1
    assert L.fab_resample_systematic_u64(None, 0, 0, None, None, None) == -1
    assert L.fab_hmc_workspace_bytes(d, 2048) >= 2 * 4 * 512
    assert L.fab_filter_workspace_bytes(100, 32) >= 100 * (3 * 32 + 3) * 4
    assert L.fab_version() >= 100
