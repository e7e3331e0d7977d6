"""Builds libvsr_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the tree).

    python -m video_super_resolution_b200.build [--force] [--verbose] [--knockout]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libvsr_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libvsr_b200.so cannot be built")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime() -> float:
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    files.append(os.path.join(HERE, "..", "include", "vsr_b200.h"))
    return max(os.path.getmtime(f) for f in files)


def build(force: bool = False, verbose: bool = False, knockout: bool = False) -> str:
    """knockout=True compiles the timing-experiment switches in (-DVSR_KNOCKOUT: VSR_DECONV_DEBUG / VSR_FUSED_DEBUG
    make results wrong on purpose); the shipped library never has them."""
    # the knock-out build has its own objects and its own file name, so it can never be mistaken for the product
    # (load it with VSR_B200_LIB=.../libvsr_b200_knockout.so)
    LIB = globals()["LIB"] if not knockout else os.path.join(HERE, "libvsr_b200_knockout.so")
    OBJ = globals()["OBJ"] if not knockout else os.path.join(HERE, "build_knockout")
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    hdr_mtime = max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC)
                    if f.endswith((".cuh", ".h")))
    hdr_mtime = max(hdr_mtime, os.path.getmtime(os.path.join(HERE, "..", "include", "vsr_b200.h")))

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_mtime):
            return obj
        cmd = [nvcc] + NVCC_FLAGS + (["-DVSR_KNOCKOUT"] if knockout else []) + \
            (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link of libvsr_b200.so failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, knockout="--knockout" in sys.argv))
