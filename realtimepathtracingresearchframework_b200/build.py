"""In-tree build of librptr_cuda.so (hand-written CUDA for sm_100a + the C ABI of include/rptr_cuda.h).

`python -m realtimepathtracingresearchframework_b200.build` or __graft_entry__.build().  nvcc cross-compiles without a
GPU; the .so is git-ignored but travels to the GPU box with the snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librptr_cuda.so")
SOURCES = ["rptr_cuda.cu", "rptr_bvh_build.cu", "rptr_host.cpp"]
HEADERS = ["rptr_math.cuh", "rptr_shading.cuh", "rptr_bvh.cuh", "rptr_trace_kernels.cuh", "rptr_reorder.cuh", "rptr_trace_tail.cuh", "rptr_post.cuh", "rptr_pointsets.cuh", "rptr_host.hpp", "rptr_bvh_build.hpp",
           "../../include/rptr_cuda.h", "../../include/rptr_types.h"]

# RPTR-FP contract (csrc/rptr_math.cuh): no FMA contraction on either side, IEEE division and square root.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false", "-prec-div=true",
              "-prec-sqrt=true", "-ftz=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-mfma,-mavx2,-O2", "-shared"]


def nvcc_path():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False, extra=()):
    if not force and not needs_build():
        return LIB
    host_cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [nvcc_path(), "-ccbin", host_cxx] + NVCC_FLAGS + list(extra) + ["-o", LIB] + [os.path.join(CSRC, f) for f in SOURCES]
    if verbose:
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose and (r.stdout or r.stderr):
        print(r.stdout + r.stderr)
    return LIB


CLI = os.path.join(HERE, "host", "rptr_cuda_cli")


def build_cli(force=False):
    """host/rptr_cuda_cli: the headless validation driver (C++ over the C ABI), linked against the in-tree librptr_cuda.so."""
    src = os.path.join(HERE, "host", "rptr_cuda_cli.cpp")
    loader = os.path.join(HERE, "host", "vks_loader.cpp")
    deps = [src, loader, os.path.join(HERE, "host", "vks_loader.hpp"), os.path.join(HERE, "host", "sky_fits.inc"),
            os.path.join(HERE, "..", "include", "rptr_cuda.h"), os.path.join(HERE, "..", "include", "rptr_types.h")]
    if not force and os.path.exists(CLI) and all(os.path.getmtime(d) <= os.path.getmtime(CLI) for d in deps):
        return CLI
    build()
    host_cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [host_cxx, "-O2", "-std=c++17", "-ffp-contract=off", "-o", CLI, src, loader, "-L" + HERE, "-l:librptr_cuda.so", "-Wl,-rpath,$ORIGIN/.."]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building rptr_cuda_cli failed:\n" + r.stdout + r.stderr)
    return CLI


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, extra=["-Xptxas", "-v"] if "--ptxas" in sys.argv else []))
