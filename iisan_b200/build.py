"""Build recipe for libiisan_b200.so (hand-written CUDA for sm_100a behind a C ABI).

    python -m iisan_b200.build            # incremental
    python -m iisan_b200.build --force

nvcc cross-compiles without a GPU.  The shared object is written in-tree (iisan_b200/lib/) so that it
travels to the GPU box with the repo snapshot; it is git-ignored.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
# Build variants (debug aids; the product is the default variant).  IISAN_B200_BUILD_VARIANT=trace compiles the chain kernels
# with wait-time accounting (-DIISAN_CHAIN_TRACE, san_chain.cu) into lib/libiisan_b200_trace.so; load it with
# IISAN_B200_LIB=<path> (iisan_b200/_lib.py), see scripts/chain_trace.py.
VARIANT = os.environ.get("IISAN_B200_BUILD_VARIANT", "")
# any other variant name: the flags come from IISAN_B200_EXTRA_FLAGS (A/B builds of kernel parameters, e.g. "-DC3_NT=5 -DC3_NW=4")
VARIANT_FLAGS = {"": [], "trace": ["-DIISAN_CHAIN_TRACE"]}.get(VARIANT, os.environ.get("IISAN_B200_EXTRA_FLAGS", "").split())
OBJDIR = os.path.join(HERE, "build" + ("_" + VARIANT if VARIANT else ""))
LIB = os.path.join(LIBDIR, "libiisan_b200" + ("_" + VARIANT if VARIANT else "") + ".so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr", "-I", INCLUDE, *VARIANT_FLAGS]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers_mtime():
    m = os.path.getmtime(os.path.join(INCLUDE, "iisan_b200.h"))
    for f in os.listdir(CSRC):
        if f.endswith(".cuh"):
            m = max(m, os.path.getmtime(os.path.join(CSRC, f)))
    return m


def compile_one(src, force, hm, log):
    obj = os.path.join(OBJDIR, src[:-3] + ".o")
    spath = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(spath), hm):
        return obj, False
    cmd = [NVCC, *FLAGS, "-c", spath, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(os.path.join(OBJDIR, src[:-3] + ".ptxas.log"), "w") as f:
        f.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if log:
        print(f"[iisan_b200.build] compiled {src}")
    return obj, True


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    hm = headers_mtime()
    with ThreadPoolExecutor(max_workers=8) as ex:
        res = list(ex.map(lambda s: compile_one(s, force, hm, verbose), sources()))
    objs = [o for o, _ in res]
    if force or any(c for _, c in res) or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[iisan_b200.build] linked {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
