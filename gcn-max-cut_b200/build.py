"""Builds gcn-max-cut_b200/lib/libgcnmaxcut.so from csrc/*.cu with nvcc for sm_100a ONLY.

    python gcn-max-cut_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container; the
resulting .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build", "obj")
LIB = os.path.join(LIBDIR, "libgcnmaxcut.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
          "--expt-relaxed-constexpr", "-Xptxas", "-warn-spills"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp(path: str) -> str:
    h = hashlib.sha1()
    deps = [path] + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(HERE, "..", "include", "gcnmaxcut.h"))
    for d in deps:
        with open(d, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(ARCH_FLAGS + COMMON).encode())
    return h.hexdigest()


def _compile(src: str, force: bool, verbose: bool) -> str:
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJDIR, src[:-3] + ".o")
    stamp_file = obj + ".stamp"
    stamp = _stamp(path)
    if not force and os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return obj
    cmd = [NVCC] + ARCH_FLAGS + COMMON + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
    if verbose or "warning" in res.stderr.lower():
        sys.stderr.write(f"--- {src} ---\n{res.stderr}\n")
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return obj


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [NVCC] + ARCH_FLAGS + ["-shared", "-o", LIB] + objs
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
