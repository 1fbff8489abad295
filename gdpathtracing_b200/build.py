"""In-tree builds of the two shared libraries (no JIT cache: the .so files travel with the repo).

  libgdpt_cuda.so   nvcc, sm_100a only, -fmad=false (arithmetic contract), -lineinfo for ncu
  libgdpt_host.so   g++ host layer, links against libgdpt_cuda.so via $ORIGIN rpath

``python -m gdpathtracing_b200.build`` rebuilds both.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.normpath(os.path.join(HERE, ".."))
INCLUDE = os.path.join(REPO, "include")
CUDA_SRC = [os.path.join(HERE, "csrc", "cuda", f) for f in ("pt_kernels.cu", "gdpt_capi.cu")]
CUDA_HDR = sorted(os.path.join(HERE, "csrc", "cuda", f) for f in os.listdir(os.path.join(HERE, "csrc", "cuda"))
                  if f.endswith((".cuh", ".h")))
HOST_SRC = [os.path.join(HERE, "csrc", "host", f) for f in ("accel_build.cpp", "geometry_group3d.cpp", "compute_shader.cpp",
                                                            "path_tracing_camera.cpp", "host_capi.cpp")]
HOST_HDR = [os.path.join(HERE, "csrc", "host", f) for f in ("accel_build.h", "geometry_group3d.h", "compute_shader.h",
                                                            "path_tracing_camera.h", "xform_math.h")]
API_HDR = [os.path.join(INCLUDE, f) for f in ("gdpt.h", "gdpt_wire.h", "gdpt_host.h")]
CUDA_LIB = os.path.join(HERE, "libgdpt_cuda.so")
HOST_LIB = os.path.join(HERE, "libgdpt_host.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC", "-shared"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    print("+", " ".join(cmd), file=sys.stderr, flush=True)
    subprocess.run(cmd, check=True)


def nvcc_path():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build_cuda(force=False, extra=()):
    if force or _stale(CUDA_LIB, CUDA_SRC + CUDA_HDR + API_HDR):
        _run([nvcc_path()] + NVCC_FLAGS + list(extra) + ["-I", INCLUDE, "-o", CUDA_LIB] + CUDA_SRC)
    return CUDA_LIB


def build_host(force=False):
    if force or _stale(HOST_LIB, HOST_SRC + HOST_HDR + API_HDR + [CUDA_LIB]):
        _run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-Wall", "-shared", "-I", INCLUDE,
              "-I", os.path.join(HERE, "csrc", "host"), "-o", HOST_LIB] + HOST_SRC
             + ["-L", HERE, "-lgdpt_cuda", "-Wl,-rpath,$ORIGIN"])
    return HOST_LIB


def build_all(force=False):
    build_cuda(force)
    build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
