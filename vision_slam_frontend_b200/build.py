"""Build recipe for libvsf_cuda.so (sm_100a only, in-tree).

`python -m vision_slam_frontend_b200.build` or `build_all()` from
__graft_entry__.build().  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB = os.path.join(PKG, "libvsf_cuda.so")
FRONTEND_LIB = os.path.join(PKG, "libvsf_frontend.so")
NCCL_LIB = os.path.join(PKG, "libvsf_nccl.so")
DRIVER_BIN = os.path.join(PKG, "vsf_sequence_driver")
PROBE_BIN = os.path.join(PKG, "vsf_latency_probe")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")

CUDA_SOURCES = ["knn2_kernel.cu", "knn2_tc_kernel.cu", "knn2_tc64_kernel.cu", "stereo_kernels.cu", "sort_kernel.cu",
                "aux_kernels.cu", "vsf_api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "-std=c++17", "-Xcompiler", "-fPIC,-fopenmp", "--use_fast_math=false"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(INCLUDE, "vsf.h"))
    return hdrs


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    if os.environ.get("VSF_TC_TRACE"):
        # per-CTA timeline stamps inside knn2_tc_kernel (tools/tc_timeline.py); costs ~6 % of the
        # kernel even when switched off at run time, so never part of a normal build
        flags.append("-DVSF_TC_TRACE")
    if os.environ.get("VSF_TC_BUCKET"):
        flags.append("-DVSF_TC_BUCKET=" + os.environ["VSF_TC_BUCKET"])   # selection bucket of the tensor engine (8, 16, 32; default 16)
    if os.environ.get("VSF_TC_BRINGUP"):
        flags.append("-DVSF_TC_BRINGUP")   # engine flags 2 / 4 (tools/tc_scaling.py, tools/tc_probe.py)
    objs = []
    rebuilt = False
    for src in CUDA_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + _deps()):
            cmd = [_nvcc()] + flags + ["-I", INCLUDE, "-I", CSRC, "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), file=sys.stderr)
            subprocess.check_call(cmd)
            rebuilt = True
        objs.append(o)
    if rebuilt or not os.path.exists(LIB):
        cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB + ".tmp"] + objs + ["-cudart", "static", "-Xcompiler", "-fopenmp"]
        subprocess.check_call(cmd)
        os.replace(LIB + ".tmp", LIB)
    return LIB


def build_nccl(force: bool = False) -> str:
    """libvsf_nccl.so: the result gathers of the sharded path (include/vsf_nccl.h)."""
    src = os.path.join(CSRC, "nccl", "vsf_nccl.cc")
    if force or _stale(NCCL_LIB, [src, os.path.join(INCLUDE, "vsf_nccl.h"), os.path.join(INCLUDE, "vsf.h"), LIB]):
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-I", INCLUDE,
               "-I", os.path.join(CUDA_HOME, "include"), "-o", NCCL_LIB + ".tmp", src,
               "-L", PKG, "-lvsf_cuda", "-L", os.path.join(CUDA_HOME, "lib64"), "-lcudart", "-lnccl",
               "-Wl,-rpath,$ORIGIN"]
        subprocess.check_call(cmd)
        os.replace(NCCL_LIB + ".tmp", NCCL_LIB)
    return NCCL_LIB


def build_driver(force: bool = False) -> str:
    """vsf_sequence_driver: synthetic stereo source -> sharded slam::Frontend -> one SLAMProblem."""
    src = os.path.join(CSRC, "frontend", "sequence_driver.cc")
    if force or _stale(DRIVER_BIN, [src, FRONTEND_LIB, NCCL_LIB]):
        cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-I", INCLUDE, "-I", os.path.join(CSRC, "frontend"),
               "-o", DRIVER_BIN + ".tmp", src, "-L", PKG, "-lvsf_frontend", "-lvsf_nccl", "-lvsf_cuda",
               "-Wl,-rpath,$ORIGIN", "-lpthread"]
        subprocess.check_call(cmd)
        os.replace(DRIVER_BIN + ".tmp", DRIVER_BIN)
    psrc = os.path.join(CSRC, "frontend", "latency_probe.cc")
    if force or _stale(PROBE_BIN, [psrc, FRONTEND_LIB]):
        cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-I", INCLUDE, "-I", os.path.join(CSRC, "frontend"),
               "-o", PROBE_BIN + ".tmp", psrc, "-L", PKG, "-lvsf_frontend", "-lvsf_cuda", "-Wl,-rpath,$ORIGIN"]
        subprocess.check_call(cmd)
        os.replace(PROBE_BIN + ".tmp", PROBE_BIN)
    return DRIVER_BIN


def build_frontend(force: bool = False) -> str:
    """C++ host mirror of slam::Frontend on top of the C ABI."""
    fdir = os.path.join(CSRC, "frontend")
    if not os.path.isdir(fdir):
        return ""
    srcs = [os.path.join(fdir, f) for f in sorted(os.listdir(fdir))
            if f.endswith(".cc") and f not in ("sequence_driver.cc", "latency_probe.cc")]
    hdrs = [os.path.join(fdir, f) for f in os.listdir(fdir) if f.endswith(".h")]
    if not srcs:
        return ""
    if force or _stale(FRONTEND_LIB, srcs + hdrs + [os.path.join(INCLUDE, "vsf.h"), LIB]):
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-I", INCLUDE, "-I", fdir,
               "-o", FRONTEND_LIB + ".tmp"] + srcs + [
                   "-L", PKG, "-lvsf_cuda", "-Wl,-rpath,$ORIGIN"]
        subprocess.check_call(cmd)
        os.replace(FRONTEND_LIB + ".tmp", FRONTEND_LIB)
    return FRONTEND_LIB


def build_all(force: bool = False, verbose: bool = False) -> dict:
    out = {"cuda": build_cuda(force, verbose), "frontend": build_frontend(force)}
    out["nccl"] = build_nccl(force)
    out["driver"] = build_driver(force)
    return out


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
