"""Build recipe for the *real* reference kernels, compiled where they lie.

TEST INFRASTRUCTURE ONLY -- nothing under ``oracle/`` is part of the product path.

The reference ships its native numerics as two torch extensions that it JIT-builds at
import time (reference ``src/dmx/compressor/quant/quant_function.py:6-28``):

* ``quant_cpu``  <- ``quant/quant_cpu/{quant_cpu,bit_helper,sim_helper}.cpp``
* ``quant_cuda`` <- ``quant/quant_cuda/{quant_cuda.cpp,block_kernel.cu,float_kernel.cu,
  fixed_point_kernel.cu,quant.cu}`` (``bit_helper.cu`` / ``sim_helper.cu`` are textually
  ``#include``d by the kernel TUs, reference ``block_kernel.cu:1-3``)

This script compiles those files *in place from /root/reference* with plain ``g++`` /
``nvcc`` command lines (no reference build system, no source copied into this repo) and
writes only binaries into ``oracle/_ref/``:

* ``oracle/_ref/ref_quant_cpu.so``   -- CPU oracle + ``cpu_baseline.kind == "reference"``
* ``oracle/_ref/ref_quant_cuda.so``  -- the reference's own CUDA kernels for sm_100, used
  on the GPU box as a secondary oracle (FixedPoint tie mode, stochastic parity) and as a
  "reference CUDA path" timing comparison.

``oracle/_ref/`` is git-ignored but not gpurun-ignored, so the binaries travel to the GPU
box; ``/root/reference`` does not exist there and is never read at run time.
"""
from __future__ import annotations

import os
import shlex
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_ROOT = os.environ.get("DMX_REFERENCE_ROOT", "/root/reference")
QDIR = os.path.join(REF_ROOT, "src", "dmx", "compressor", "quant")


def _torch_flags():
    import torch
    from torch.utils import cpp_extension as ce

    inc = ce.include_paths() + [sysconfig.get_paths()["include"]]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    return inc, libdir, abi


def _run(cmd):
    print("+", " ".join(shlex.quote(c) for c in cmd), flush=True)
    subprocess.check_call(cmd)


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources + [os.path.abspath(__file__)])


def build_cpu(force=False):
    srcs = [os.path.join(QDIR, "quant_cpu", f) for f in ("quant_cpu.cpp", "bit_helper.cpp", "sim_helper.cpp")]
    out = os.path.join(OUT, "ref_quant_cpu.so")
    if not all(os.path.exists(s) for s in srcs):
        return out if os.path.exists(out) else None
    if not force and _newer(out, srcs):
        return out
    os.makedirs(OUT, exist_ok=True)
    inc, libdir, abi = _torch_flags()
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w",
           "-DTORCH_EXTENSION_NAME=ref_quant_cpu", "-DTORCH_API_INCLUDE_EXTENSION_H",
           f"-D_GLIBCXX_USE_CXX11_ABI={abi}"]
    for i in inc:
        cmd += ["-isystem", i]
    cmd += srcs + ["-o", out, f"-L{libdir}", "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_python",
                   f"-Wl,-rpath,{libdir}"]
    _run(cmd)
    return out


def build_cuda(force=False):
    cdir = os.path.join(QDIR, "quant_cuda")
    cu = [os.path.join(cdir, f) for f in ("block_kernel.cu", "float_kernel.cu", "fixed_point_kernel.cu", "quant.cu")]
    cpp = os.path.join(cdir, "quant_cuda.cpp")
    out = os.path.join(OUT, "ref_quant_cuda.so")
    if not all(os.path.exists(s) for s in cu + [cpp]):
        return out if os.path.exists(out) else None
    if not force and _newer(out, cu + [cpp]):
        return out
    os.makedirs(OUT, exist_ok=True)
    inc, libdir, abi = _torch_flags()
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    objs = []
    common = ["-DTORCH_EXTENSION_NAME=ref_quant_cuda", "-DTORCH_API_INCLUDE_EXTENSION_H",
              f"-D_GLIBCXX_USE_CXX11_ABI={abi}"]
    for s in cu:
        o = os.path.join(OUT, os.path.basename(s) + ".o")
        # the reference passes no arch flags (quant_function.py:14-26); sm_100 is what the
        # B200 box needs.  Default nvcc numerics (no fast-math), exactly as the reference.
        cmd = [os.path.join(cuda_home, "bin", "nvcc"), "-O2", "-std=c++17", "-w", "-c",
               "-gencode", "arch=compute_100,code=sm_100", "-Xcompiler", "-fPIC"] + common
        for i in inc + [os.path.join(cuda_home, "include")]:
            cmd += ["-isystem", i]
        cmd += [s, "-o", o]
        _run(cmd)
        objs.append(o)
    o = os.path.join(OUT, "quant_cuda.cpp.o")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-w", "-c"] + common
    for i in inc + [os.path.join(cuda_home, "include")]:
        cmd += ["-isystem", i]
    cmd += [cpp, "-o", o]
    _run(cmd)
    objs.append(o)
    _run(["g++", "-shared", "-o", out] + objs +
         [f"-L{libdir}", "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10", "-lc10_cuda", "-ltorch_python",
          f"-L{os.path.join(cuda_home, 'lib64')}", "-lcudart", f"-Wl,-rpath,{libdir}"])
    for o in objs:
        os.remove(o)
    return out


def stage_python(force=False):
    """Stage the reference's python package verbatim into ``oracle/_ref/pysrc/dmx`` (git-ignored, like the binaries
    above: made by this committed recipe from /root/reference, never part of the history, but it travels to the GPU
    box, where /root/reference does not exist).  It is what lets the plugin parity tests run the reference's OWN
    ``CastTo`` / ``Sparsify`` / ``DmxModule`` code on a B200, unpatched (its CUDA path) and patched (libdmxq), and what
    bench.py's reference arms call.  Only ``*.py`` files are staged; the native sources stay where they are (the
    prebuilt ``ref_quant_{cpu,cuda}.so`` answer the import-time JIT, see oracle/refshim/load_reference.py)."""
    import shutil

    src = os.path.join(REF_ROOT, "src", "dmx")
    dst = os.path.join(OUT, "pysrc", "dmx")
    if not os.path.isdir(src):
        return dst if os.path.isdir(dst) else None
    stamp = os.path.join(OUT, "pysrc", ".staged")
    newest = max(os.path.getmtime(os.path.join(d, f)) for d, _, fs in os.walk(src) for f in fs if f.endswith(".py"))
    if not force and os.path.exists(stamp) and os.path.getmtime(stamp) >= newest:
        return dst
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    n = 0
    for d, _, fs in os.walk(src):
        for f in fs:
            if not f.endswith(".py"):
                continue
            rel = os.path.relpath(os.path.join(d, f), src)
            os.makedirs(os.path.dirname(os.path.join(dst, rel)), exist_ok=True)
            shutil.copy2(os.path.join(d, f), os.path.join(dst, rel))
            n += 1
    with open(stamp, "w") as fh:
        fh.write(f"{n} python files staged verbatim from {src}\n")
    return dst


def load(name):
    """Import oracle/_ref/<name>.so as a python module (None if it was never built)."""
    import importlib.util

    path = os.path.join(OUT, name + ".so")
    if not os.path.exists(path):
        return None
    import torch  # noqa: F401  (libtorch must be loaded first)

    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    force = "--force" in sys.argv
    if not os.path.isdir(QDIR):
        print(f"reference not present at {QDIR}; keeping prebuilt oracle/_ref as is")
        return 0
    print("cpu :", build_cpu(force))
    if "--no-cuda" not in sys.argv:
        print("cuda:", build_cuda(force))
    print("python:", stage_python(force))
    return 0


if __name__ == "__main__":
    sys.exit(main())
