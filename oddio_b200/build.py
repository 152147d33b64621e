"""Builds liboddio_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m oddio_b200.build [--force] [--verbose]

The library links the CUDA runtime statically, so at run time it needs only the driver
(libcuda.so.1). There is no CPU build: without nvcc this raises.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.environ.get("ODB_SO") or os.path.join(HERE, "liboddio_b200.so")  # ODB_SO / ODB_NVCC_EXTRA: developer experiments
SOURCES = ["odb_host.cu", "odb_scene.cu", "odb_mixer.cu", "odb_spatial.cu", "odb_mix_fast.cu", "odb_scene_mix.cu", "odb_mixer_kernels.cu", "odb_ring.cu", "odb_exchange.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # the reference never contracts a*b+c; FMA is opt-in via intrinsics
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-Wall",
    "-cudart", "static",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: oddio_b200 has no CPU build")
    return exe


def _deps():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    out.append(os.path.join(os.path.dirname(HERE), "include", "oddio_b200.h"))
    out.append(os.path.abspath(__file__))
    return out


def stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(p) > t for p in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return SO
    objs = []
    bdir = os.path.join(HERE, "build", os.path.basename(SO).replace(".so", ""))
    extra = os.environ.get("ODB_NVCC_EXTRA", "").split()
    os.makedirs(bdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(bdir, src.replace(".cu", ".o"))
        cmd = [nvcc(), *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out, file=sys.stderr)
    link = [nvcc(), "-shared", "-o", SO, *objs, "-cudart", "static", "-Xcompiler", "-fPIC",
            "-Xlinker", "--exclude-libs,ALL"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
