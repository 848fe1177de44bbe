"""Build the sm_100a shared library IN-TREE (racon_gpu_b200/libb200poa.so).

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box with
the gpurun snapshot.  `python -m racon_gpu_b200.build` or __graft_entry__.build().
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200poa.so")
SOURCES = ["b200poa.cu", "b200aln.cu", "host/cuda_batch.cpp", "host/cuda_polisher.cpp"]
HEADERS = ["poa_core.cuh", "poa_fill.cuh", "poa_simt.cuh", "host/b200_window.hpp", "host/cuda_batch.hpp",
           "host/cuda_polisher.hpp", "host/b200poa_batch.hpp", "host/window_arena.hpp", "../../include/b200poa.h",
           "aln_core.cuh", "host/aln_levels.hpp", "host/b200aln_aligner.hpp", "../../include/b200aln.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wextra,-pthread", "-shared", "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the B200 POA engine has no CPU fallback")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS:
        path = os.path.join(CSRC, f)
        if os.path.exists(path) and os.path.getmtime(path) > t:
            return True
    return False


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    srcs = [os.path.join(CSRC, f) for f in SOURCES if os.path.exists(os.path.join(CSRC, f))]
    cmd = [_nvcc(), *NVCC_FLAGS, "-x", "cu", "-I", os.path.join(HERE, "..", "include"), "-I", CSRC,
           "-o", LIB, *srcs]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log"), "w") as fh:
        fh.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libb200poa.so")
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
