"""Builds libfab_b200.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build().

Two translation units (the warp-level engine + misc kernels, and the row-tile tcgen05 engine) are
compiled to objects under build/obj and linked; only a unit whose sources changed is recompiled.
FAB_NVCC_FLAGS adds experiment flags (e.g. -DFAB_UMMA_WATCHDOG); a change of flags rebuilds."""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
INC = os.path.join(ROOT, "include", "fab_b200.h")
UNITS = {
    "fab_b200.cu": ["common.cuh", "mma_gemm.cuh", "flow_tile.cuh", "param_grad.cuh", "target_tile.cuh", "tile_kernels.cuh",
                    "misc_kernels.cuh", "buffer_kernels.cuh", "reduce_finish.cuh", "host_util.h", INC],
    "fab_umma.cu": ["common.cuh", "target_tile.cuh", "reduce_finish.cuh", "umma.cuh", "umma_engine.cuh",
                    "host_util.h", INC],
}
SOURCES = list(UNITS)
OUT = os.path.join(HERE, "libfab_b200.so")
OBJ_DIR = os.path.join(ROOT, "build", "obj")


def _abs(p):
    return p if os.path.isabs(p) else os.path.join(HERE, p)


def _flags():
    return os.environ.get("FAB_NVCC_FLAGS", "").split()


def _obj(src):
    tag = hashlib.sha1(" ".join(_flags()).encode()).hexdigest()[:8]
    return os.path.join(OBJ_DIR, f"{os.path.splitext(src)[0]}.{tag}.o")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(_abs(d)) > t for d in deps)


def needs_build() -> bool:
    return any(_stale(_obj(s), [s] + d) for s, d in UNITS.items()) or \
        _stale(OUT, [_obj(s) for s in UNITS])


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    base = [nvcc] + _flags() + ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                                "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include")]
    if verbose:
        base.insert(1, "-Xptxas=-v")
    procs = []
    for src, deps in UNITS.items():
        if force or _stale(_obj(src), [src] + deps):
            cmd = base + ["-c", "-o", _obj(src), os.path.join(HERE, src)]
            print("[fab_torch_b200] " + " ".join(cmd), flush=True)
            procs.append((cmd, subprocess.Popen(cmd, cwd=HERE)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", OUT] + [_obj(s) for s in UNITS]
    print("[fab_torch_b200] " + " ".join(link), flush=True)
    subprocess.run(link, check=True, cwd=HERE)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
