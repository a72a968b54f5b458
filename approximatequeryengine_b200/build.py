"""Builds the native parts IN-TREE (the .so files travel to the GPU box with the snapshot):

  approximatequeryengine_b200/_lib/libaqe_b200.so                 C-ABI engine: CUDA kernels for sm_100a + host code
  approximatequeryengine_b200/_lib/aqe_backend<EXT_SUFFIX>        pybind11 module with the reference's surface

    python -m approximatequeryengine_b200.build [--force] [--verbose]

nvcc cross-compiles sm_100a without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "_lib")
INCLUDE = os.path.join(ROOT, "include")
LIB = os.path.join(LIBDIR, "libaqe_b200.so")
EXT = os.path.join(LIBDIR, "aqe_backend" + sysconfig.get_config_var("EXT_SUFFIX"))

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _host_cxx() -> str:
    # the image's /opt/gcc wrapper links libstdc++ statically; the system g++ is the one that matches
    # the libstdc++.so.6 every other extension in the process uses
    for c in (os.environ.get("AQE_CXX"), "/usr/bin/g++", shutil.which("g++")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("g++ not found")


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed:\n" + " ".join(cmd) + "\n" + r.stdout)
    if verbose and r.stdout.strip():
        print(r.stdout)
    return r.stdout


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> None:
    os.makedirs(LIBDIR, exist_ok=True)
    headers = [os.path.join(INCLUDE, "aqe_b200.h")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp", ".h"))]
    lib_src = [os.path.join(CSRC, "aqe_engine.cu"), os.path.join(CSRC, "aqe_plan.cpp"), os.path.join(CSRC, "aqe_sql.cpp"), os.path.join(CSRC, "aqe_order.cpp")]
    cxx = _host_cxx()
    if force or not _newer(LIB, lib_src + headers):
        cmd = [_nvcc(), "-std=c++17", "-O3", *ARCH, "-lineinfo", "-ccbin", cxx, "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden",
               "-shared", "-I", INCLUDE, "-I", CSRC, "-DAQE_BUILDING", "-x", "cu", *lib_src, "-o", LIB]
        if ptxas_info:
            cmd[1:1] = ["-Xptxas", "-v"]
        out = _run(cmd, verbose)
        if ptxas_info:
            print(out)
    ext_src = [os.path.join(CSRC, "aqe_pybind.cpp")]
    if os.path.exists(ext_src[0]) and (force or not _newer(EXT, ext_src + headers + [LIB])):
        import pybind11
        cmd = [cxx, "-std=c++17", "-O2", "-fPIC", "-shared", "-fvisibility=hidden", "-I", INCLUDE, "-I", pybind11.get_include(),
               "-I", sysconfig.get_paths()["include"], *ext_src, "-o", EXT, "-L", LIBDIR, "-laqe_b200", "-Wl,-rpath,$ORIGIN"]
        _run(cmd, verbose)


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv, ptxas_info="--ptxas" in sys.argv)
    print("built", LIB, EXT if os.path.exists(EXT) else "")
