"""Builds rmem_b200/lib/librmem_b200.so (hand-written sm_100a CUDA + the C ABI) with nvcc, in-tree.

    python -m rmem_b200.build [--force] [--verbose]

nvcc cross-compiles sm_100a without a GPU; the .so travels to the GPU box with the repo snapshot.
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
LIB = os.path.join(LIBDIR, "librmem_b200.so")
SOURCES = ["capi.cu", "gemm.cu", "gemm_tc.cu", "tma.cu", "ops.cu", "attn_dense.cu", "attn_tc2.cu", "attn_tc3.cu", "mha_tc.cu", "local_attn_tc.cu", "engine.cu", "train_loss.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--use_fast_math", "-Xptxas", "-v"]
# --use_fast_math (approximate exp / log / division, FTZ) is for the tensor-core kernels, whose exponentials are MUFU ex2
# by design.  ops.cu holds the mask head, the soft aggregation, evict_rel and the TTA head, whose softmax / logit / argmax
# arithmetic must be the reference's IEEE fp32 (expf, logf, exact division next to the explicit _rn operations): it is
# compiled WITHOUT the flag; so is train_loss.cu (the training loss and its gradient: expf / logf / division as torch's).
EXACT_MATH = {"ops.cu", "train_loss.cu"}


def _digest() -> str:
    h = hashlib.sha256()
    for root, _, files in os.walk(CSRC):
        for f in sorted(files):
            with open(os.path.join(root, f), "rb") as fh:
                h.update(f.encode()); h.update(fh.read())
    with open(os.path.join(HERE, "..", "include", "rmem_b200.h"), "rb") as fh:
        h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    h.update(",".join(sorted(EXACT_MATH)).encode())
    return h.hexdigest()


def up_to_date() -> bool:
    """True when the .so in the tree was built from exactly the sources / header / flags of this tree."""
    stamp = os.path.join(LIBDIR, "build.sha256")
    return os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == _digest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "build.sha256")
    dig = _digest()
    if not force and up_to_date():
        return LIB
    if not os.path.exists(NVCC):
        raise RuntimeError(f"nvcc not found at {NVCC} and no up-to-date {LIB}; cannot build the CUDA extension")

    def compile_one(src):
        obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        flags = [f for f in FLAGS if not (src in EXACT_MATH and f == "--use_fast_math")]
        cmd = [NVCC, *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    with ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(compile_one, SOURCES))
    objs = []
    log = []
    for src, obj, r in results:
        log.append(f"==== {src}\n{r.stdout}\n{r.stderr}")
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        objs.append(obj)
    with open(os.path.join(LIBDIR, "build.log"), "w") as fh:
        fh.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
