"""Build libstmask_b200.so in-tree with nvcc for sm_100a (no torch, no JIT cache).

    python -m stmask_b200.build [--force] [--verbose]

The shared library lands in stmask_b200/lib/ (git-ignored, travels to the GPU box with the
snapshot).  Cross-compiles without a GPU.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libstmask_b200.so")
SOURCES = ["abi.cu", "dcn_simt.cu", "dcn_tc.cu", "conv_tma.cu", "corr_simt.cu", "corr_tc.cu", "layout.cu", "roi_align.cu", "temporal_net.cu", "detect_nms.cu", "mask_assembly.cu", "track.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
              "-Xptxas", "-v"] + (["-DSTM_DCN_EXPERIMENTS"] if os.environ.get("STM_DCN_EXPERIMENTS") else []) + \
             (["-DSTM_CONV_TMA_TRACE"] if os.environ.get("STM_CONV_TMA_TRACE") else []) + \
             (["-DSTM_DCN_TRACE"] if os.environ.get("STM_DCN_TRACE") else [])


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; libstmask_b200.so cannot be built")
    return cand


def _digest() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ["../../include/stmask_b200.h"]
    for f in files:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    stamp = LIB + ".stamp"
    return os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == _digest()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and is_current():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    logs = {}

    def compile_one(src: str) -> str:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        logs[src] = r.stderr + r.stdout
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stderr}\n{r.stdout}")
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    # hidden visibility + explicit default visibility on the extern "C" entry points keeps the
    # exported surface equal to include/stmask_b200.h
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stderr}")
    with open(os.path.join(LIBDIR, "ptxas.log"), "w") as f:
        for s in SOURCES:
            f.write(f"==== {s} ====\n{logs[s]}\n")
    with open(LIB + ".stamp", "w") as f:
        f.write(_digest())
    if verbose:
        print(open(os.path.join(LIBDIR, "ptxas.log")).read())
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
