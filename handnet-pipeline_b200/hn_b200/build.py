"""Build libhandnet_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m hn_b200.build [--force] [--debug] [-v]

--debug compiles the convolution kernel's bring-up instrumentation in (-DHN_CONV_DEBUG: the timing-experiment flags of
hn_conv_desc.debug and the clock64 trace); production builds carry none of it.

The shared library is the product's only native artefact: a C-ABI (include/handnet_b200.h) over the
hand-written CUDA kernels in csrc/.  It links the shared CUDA runtime so that it shares the runtime
instance PyTorch has already loaded in the process.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhandnet_b200.so")
SOURCES = ["hn_lib.cu", "hn_conv_igemm.cu", "hn_pre.cu", "hn_post.cu", "hn_pose.cu", "hn_io.cu", "hn_mesh.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-cudart", "shared"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "..", "include", "handnet_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, debug: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if debug:
            cmd.insert(1, "-DHN_CONV_DEBUG")
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out)
    cmd = [_nvcc(), "-shared", "-cudart", "shared", "-o", LIB, *objs]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv or "--debug" in sys.argv, verbose="-v" in sys.argv, debug="--debug" in sys.argv))
