"""In-tree build of libkge_b200.so (sm_100a only) with nvcc.

The shared library is the whole native product: hand-written CUDA kernels behind the C ABI
declared in include/kge_b200.h.  It is built next to this file so that it travels with the
source tree (the GPU box receives the tree, not a JIT cache).
"""

from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
BUILD_DIR = os.path.join(ROOT, "build", "kge_b200")
LIB_PATH = os.path.join(HERE, "libkge_b200.so")

NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-O3",
    "-std=c++17",
    "-Xcompiler",
    "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libkge_b200.so")


def have_nvcc() -> bool:
    try:
        _nvcc()
        return True
    except RuntimeError:
        return False


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime() -> float:
    files = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    files += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    return max(os.path.getmtime(f) for f in files)


def needs_build() -> bool:
    return not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < _deps_mtime()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every csrc/*.cu for sm_100a and link libkge_b200.so.  Returns the library path."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(BUILD_DIR, exist_ok=True)
    # one builder at a time: the ranks of a multi-GPU job all import the package at once and share the object files
    import fcntl

    with open(os.path.join(BUILD_DIR, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not needs_build():   # another process built it while this one waited
            return LIB_PATH
        return _build_locked(nvcc, force, verbose)


def _build_locked(nvcc: str, force: bool, verbose: bool) -> str:
    header_mtime = max(
        os.path.getmtime(os.path.join(d, f))
        for d in (CSRC, INCLUDE)
        for f in os.listdir(d)
        if f.endswith((".cuh", ".h"))
    )

    def compile_one(src: str) -> str:
        obj = os.path.join(BUILD_DIR, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), header_mtime):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-I", CSRC, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            print(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, sources()))
    tmp = LIB_PATH + ".tmp"
    res = subprocess.run([nvcc, "-shared", "-o", tmp, *objs, "-lcudart_static", "-ldl", "-lrt", "-lpthread"],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
