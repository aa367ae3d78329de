"""Build libmsf_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python multimodal-sensor-fusion-with-attention-rajeevatla_b200/build.py [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# MSF_BUILD_VARIANT=timeline: a debug library with the per-kernel step timeline compiled in (msf_common.cuh),
# built beside the product library as libmsf_b200_timeline.so (select it with MSF_B200_LIB)
VARIANT = os.environ.get("MSF_BUILD_VARIANT", "")
OBJ = os.path.join(HERE, "build" + ("_" + VARIANT if VARIANT else ""))
LIB = os.path.join(HERE, "libmsf_b200" + ("_" + VARIANT if VARIANT else "") + ".so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr",
    "-Xptxas", "-v" if os.environ.get("MSF_PTXAS_V") else "-O3",
]
if VARIANT == "timeline":
    FLAGS.append("-DMSF_TIMELINE")
elif VARIANT:
    raise RuntimeError(f"unknown MSF_BUILD_VARIANT {VARIANT!r}")


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _newest_header():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "msf_b200.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, force, verbose):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    path = os.path.join(CSRC, src)
    stale = (force or not os.path.exists(obj)
             or os.path.getmtime(obj) < max(os.path.getmtime(path), _newest_header()))
    if stale:
        cmd = [NVCC, *FLAGS, "-c", path, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        if verbose or os.environ.get("MSF_PTXAS_V"):
            sys.stderr.write(res.stderr)
    return obj, stale


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    with concurrent.futures.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        results = list(ex.map(lambda s: _compile(s, force, verbose), _sources()))
    objs = [o for o, _ in results]
    if force or any(stale for _, stale in results) or not os.path.exists(LIB):
        # static cudart (self-contained next to torch's own runtime); the driver API
        # (cuTensorMapEncodeTiled) is fetched with cudaGetDriverEntryPoint, so there is
        # no link-time libcuda dependency and the .so loads on a box without a driver.
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-cudart", "static"]
        if verbose:
            print(" ".join(cmd), flush=True)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
