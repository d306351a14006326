"""Build libnbg_b200.so (the C-ABI library: hand-written sm_100a kernels) in-tree with nvcc.

    python -m numbagg_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU.  The .so is git-ignored but travels to the GPU
box with the gpurun snapshot.  The CUDA runtime is linked statically, so the library depends
only on the driver (libcuda) at run time and shares the primary context (and therefore
device pointers and streams) with PyTorch in the same process.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(CSRC, "_build")
LIB = os.path.join(HERE, "libnbg_b200.so")
SOURCES = ["nbg_abi.cu", "nbg_move.cu", "nbg_move_exp.cu", "nbg_fill.cu", "nbg_group.cu", "nbg_reduce.cu", "nbg_quantile.cu", "nbg_matrix.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [os.path.join("..", "..", "include", "nbg_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",  # the reference never contracts a*b+c (numbagg/decorators.py:34-49)
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-O2",
]


def nvcc() -> str:
    cand = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    if os.path.exists(cand):
        return cand
    found = shutil.which("nvcc")
    if not found:
        raise RuntimeError("nvcc not found: numbagg_b200 needs the CUDA toolkit to build its kernels")
    return found


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    env = dict(os.environ)
    # the image exports CC/CXX pointing at a wrapper; nvcc should use the system g++
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else None

    def compile_one(src: str) -> str:
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        path = os.path.join(CSRC, src)
        if force or _stale(obj, [path] + hdrs):
            cmd = [nvcc(), *NVCC_FLAGS, "-c", path, "-o", obj]
            if ccbin:
                cmd += ["-ccbin", ccbin]
            if verbose:
                cmd += ["-Xptxas", "-v"]
            r = subprocess.run(cmd, capture_output=True, text=True, env=env)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        objs = list(ex.map(compile_one, srcs))
    if force or _stale(LIB, objs):
        cmd = [nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        if ccbin:
            cmd += ["-ccbin", ccbin]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
