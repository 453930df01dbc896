"""In-tree build of libtdet_b200.so with nvcc for sm_100a (`python -m torch_detection_b200.build`).

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels with the working tree.
"""
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OUT = os.path.join(CSRC, "libtdet_b200.so")
SOURCES = ["tdet_api.cu"]
HEADERS = ["conv_gemm.cuh", "conv_swap.cuh", "bottleneck_fused.cuh", "bottleneck_tail2.cuh", "wgrad_gemm.cuh", "aux_kernels.cuh", "ptx_sm100.cuh",
           os.path.join("..", "..", "include", "tdet_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",  # NOT -arch=sm_100a: ptxas rejects tcgen05 there
    "-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wno-format-truncation",
]


def needs_build():
    if not os.path.isfile(OUT):
        return True
    t = os.path.getmtime(OUT)
    for f in SOURCES + HEADERS:
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return False


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (res.stdout, res.stderr))
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
