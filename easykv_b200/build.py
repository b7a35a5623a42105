"""Builds `easykv_b200/libeasykv_b200.so` in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m easykv_b200.build [--force] [--verbose]

The shared library is the product's only compute path; nothing falls back to PyTorch when it is
missing (`easykv_b200._lib` raises).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libeasykv_b200.so")
SOURCES = ["ekv_api.cu", "ekv_decode.cu", "ekv_decode_cluster.cu", "ekv_chunk.cu", "ekv_chunk_tc.cu", "ekv_aux.cu", "ekv_sample.cu", "ekv_umma_probe.cu", "ekv_chunk_umma.cu", "ekv_decode_umma.cu"]
HEADERS = ["ekv_common.cuh", "ekv_select.cuh", "ekv_bucket.cuh", "ekv_decode_common.cuh", "ekv_mma.cuh", "ekv_umma.cuh", "ekv_chunk_plan.h", "ekv_kernels.h", os.path.join("..", "..", "include", "easykv_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--expt-relaxed-constexpr", "--expt-extended-lambda", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-fvisibility=hidden", "-Xcompiler", "-Wall"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    procs = []
    for s in SOURCES:
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, s), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for s, obj, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {s}")
        objs.append(obj)
    cmd = [nvcc, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static",
           "-Xcompiler", "-fPIC"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
