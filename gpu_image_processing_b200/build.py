"""Builds libgip_b200.so in-tree with nvcc for sm_100a (and nothing else).

    python -m gpu_image_processing_b200.build [--force]

The shared library lands next to this file so that it travels to the GPU box with the repo
snapshot.  cudart is linked statically: the library shares the primary context with any other
runtime in the process (torch), and has no libcudart.so version to match.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libgip_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=default", "-Xptxas", "-v",
          "--expt-relaxed-constexpr"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


_INC = None


def _includes(path, seen):
    """Files reachable from `path` through #include "..." (relative to the including file)."""
    import re
    global _INC
    _INC = _INC or re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)
    if path in seen or not os.path.exists(path):
        return
    seen.add(path)
    with open(path) as f:
        text = f.read()
    for inc in _INC.findall(text):
        _includes(os.path.normpath(os.path.join(os.path.dirname(path), inc)), seen)


def _deps_mtime(src_path):
    """Newest mtime among a source file and everything it includes (headers, .inc tables)."""
    seen = set()
    _includes(src_path, seen)
    return max(os.path.getmtime(p) for p in seen)


def _compile(src: str, force: bool) -> str:
    obj = os.path.join(OBJ, src[:-3] + ".o")
    path = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > _deps_mtime(path):
        return obj
    cmd = [NVCC, *ARCH, *CFLAGS, *os.environ.get("GIP_EXTRA_NVCC_FLAGS", "").split(), "-c", path, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    with open(obj + ".log", "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"nvcc failed on {src}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force), srcs))
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, *ARCH, "-shared", "-cudart", "static", "-o", LIB, *objs]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed")
    if verbose:
        print(LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
