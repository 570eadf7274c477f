"""Builds libfab_b200.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build()."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SOURCES = ["fab_b200.cu"]
HEADERS = ["common.cuh", "mma_gemm.cuh", "flow_tile.cuh", "target_tile.cuh", "tile_kernels.cuh",
           "misc_kernels.cuh", "buffer_kernels.cuh", os.path.join(ROOT, "include", "fab_b200.h")]
OUT = os.path.join(HERE, "libfab_b200.so")


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, s) for s in SOURCES] + \
           [h if os.path.isabs(h) else os.path.join(HERE, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"),
           "-o", OUT] + [os.path.join(HERE, s) for s in SOURCES]
    extra = os.environ.get("FAB_NVCC_FLAGS", "").split()
    if extra:                      # experiment knobs, e.g. -DFAB_PROF -DFAB_MIN_CTAS=2
        cmd[1:1] = extra
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    print("[fab_torch_b200] " + " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True, cwd=HERE)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
