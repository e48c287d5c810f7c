"""Build recipe for libvml_b200.so (hand-written CUDA for sm_100a, C ABI).

    python -m vml_b200.build          # or: python video-moment-localization_b200/build.py

nvcc cross-compiles without a GPU.  Objects go to ``build/`` (git-ignored), the shared
library next to this file (git-ignored, but it travels to the GPU box with the repo
snapshot).  Only sm_100a code is generated -- there is no other backend.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libvml_b200.so")
OBJ_DIR = os.path.join(ROOT, "build", "vml_b200")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--expt-relaxed-constexpr", "-Xptxas", "-v"] + os.environ.get("VML_EXTRA_CFLAGS", "").split()


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(paths, cflags=None):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode() + b"\0" + f.read())
    h.update(" ".join(ARCH + (CFLAGS if cflags is None else cflags)).encode())
    return h.hexdigest()


LIB_DBG = os.path.join(HERE, "libvml_b200_dbg.so")


LIB_TIMING = os.path.join(HERE, "libvml_b200_timing.so")


def build_timing() -> str:
    """Development aid: the library with %globaltimer stamps in the content-unit kernels (-DVML_CU_TIMING; tools/cu_timing.py,
    tools/pp_timing.py), as ``libvml_b200_timing.so``; select it with ``VML_LIB=<path>``."""
    return _build(LIB_TIMING, os.path.join(ROOT, "build", "vml_b200_timing"), CFLAGS + ["-DVML_CU_TIMING"], False, False)


def build(force: bool = False, verbose: bool = False, debug: bool = False) -> str:
    """``debug=True``: the bounds-checking build (-DVML_DEBUG_BOUNDS, see csrc/common.cuh) as ``libvml_b200_dbg.so``; select it
    at run time with ``VML_LIB=<path>``."""
    if debug:
        return _build(LIB_DBG, os.path.join(ROOT, "build", "vml_b200_dbg"), CFLAGS + ["-DVML_DEBUG_BOUNDS"], force, verbose)
    return _build(LIB, OBJ_DIR, CFLAGS, force, verbose)


def _build(LIB, OBJ_DIR, CFLAGS, force, verbose) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + \
        [os.path.join(ROOT, "include", "vml_b200.h")]
    stamp = os.path.join(OBJ_DIR, "stamp")
    dig = _digest(deps, CFLAGS)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        cmd = [NVCC, *ARCH, *CFLAGS, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(obj + ".log", "w") as f:
            f.write(log)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{log}")
        if verbose:
            print(log)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources()))) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    if "--timing" in sys.argv:
        print(build_timing())
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, debug="--debug" in sys.argv))
