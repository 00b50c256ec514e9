"""Builds libneedle_b200.so (the C-ABI library of include/needle_b200.h) with
plain nvcc for sm_100a, in-tree, next to this file.

    python -m needle_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "libneedle_b200.so")
CAPI_LIB = os.path.join(HERE, "libneedle.so")   # the needle-capi ABI (include/needle.h) over libneedle_b200.so
OBJ_DIR = os.path.join(HERE, "_obj")

SOURCES = ["api.cu", "match.cu", "fingerprint.cu", "vote_device.cu", "multi.cu", "vote.cpp", "persist.cpp"]
HEADERS = [os.path.join(CSRC, "common.h"), os.path.join(CSRC, "fp_tables.h"), os.path.join(CSRC, "tma.cuh"),
           os.path.join(CSRC, "fp_chroma_fold.inc"),
           os.path.join(INCLUDE, "needle_b200.h"), os.path.join(INCLUDE, "needle.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall,-Wextra,-fno-fast-math,-ffp-contract=off,-mpopcnt",
    "--fmad=true", "-Xptxas", "-v",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libneedle_b200.so cannot be built")
    return exe


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    objs = []
    log = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        if force or _stale(obj, [path, __file__] + HEADERS):
            cmd = [nvcc()] + NVCC_FLAGS + ["-I", INCLUDE, "-c", path, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            log.append(r.stderr)
            if verbose:
                sys.stderr.write(r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed on %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    if force or _stale(LIB, objs):
        cmd = [nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + \
              ["-cudart", "static", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    # libneedle.so: host C++ only, resolves nb200_* from the library next to it
    capi_src = os.path.join(CSRC, "capi.cpp")
    if force or _stale(CAPI_LIB, [capi_src, LIB, __file__] + HEADERS):
        cxx = shutil.which("g++") or "g++"
        cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wextra", "-o", CAPI_LIB, capi_src,
               "-L" + HERE, "-lneedle_b200", "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose:
            sys.stderr.write(r.stderr)
        if r.returncode != 0:
            raise RuntimeError("libneedle.so failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(os.path.join(OBJ_DIR, "ptxas.log"), "a") as f:
        f.write("".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
