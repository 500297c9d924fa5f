"""In-tree build of libdmxq.so (CUDA, sm_100a only).

    python dmx_compressor_b200/build.py [--force] [--verbose]

Plain nvcc command lines, one per translation unit, run in parallel; the shared library
lands in dmx_compressor_b200/lib/libdmxq.so (git-ignored, shipped to the GPU box by gpurun).
Numerics flags are the nvcc defaults the reference is built with (no fast-math, -ftz=false,
IEEE div/sqrt); -fmad=false so no fp32 add/mul pair is ever contracted behind our back.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shlex
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libdmxq.so")
OBJDIR = os.path.join(HERE, "build")
SOURCES = ["dmxq_api.cu", "dmxq_rows_a.cu", "dmxq_rows_b.cu", "dmxq_rows_c.cu", "dmxq_rows_d.cu", "dmxq_rows_e.cu", "dmxq_rows_f.cu", "dmxq_rows_g.cu", "dmxq_cols.cu", "dmxq_misc.cu", "dmxq_softmax.cu", "dmxq_tma.cu"]
HEADERS = ["dmxq_numerics.cuh", "dmxq_kernels.cuh", "dmxq_stages.cuh", "dmxq_rows.cuh", os.path.join("..", "..", "include", "dmxq.h")]

NVCC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-fmad=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, verbose):
    obj = os.path.join(OBJDIR, src + ".o")
    deps = [os.path.join(CSRC, src)] + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    if not _stale(obj, deps):
        return obj, ""
    cmd = [NVCC] + ARCH + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{' '.join(shlex.quote(c) for c in cmd)}\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    if force:
        for s in SOURCES:
            try:
                os.remove(os.path.join(OBJDIR, s + ".o"))
            except FileNotFoundError:
                pass
    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        res = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    objs = [o for o, _ in res]
    if verbose:
        for _, log in res:
            sys.stderr.write(log)
    if _stale(LIB, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
