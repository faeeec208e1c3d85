"""Builds the in-tree native libraries of the package with explicit nvcc / g++ commands.

  csrc/libsolr_b200.so       the engine: CUDA kernels + C ABI (include/solr_b200.h), sm_100a only
  csrc/libsolr_b200_host.so  host-side scene container (setters + box compaction), plain C++

nvcc cross-compiles without a GPU.  --use_fast_math matches the reference engine's own build flags
(/root/reference/solr/CMakeLists.txt:41-44) so that / sqrtf rsqrtf sinf cosf powf lower to the same
approximate instructions the reference's pixels are made of.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ENGINE_LIB = os.path.join(CSRC, "libsolr_b200.so")
HOST_LIB = os.path.join(CSRC, "libsolr_b200_host.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CXX = "/usr/bin/g++"

NVCC_FLAGS = ["-std=c++17", "-O3", "--use_fast_math", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-ccbin", CXX, "-cudart", "static", "-shared",
              # keep the statically linked CUDA runtime private: another libcudart (PyTorch's) may share the process
              "-Xlinker", "-Bsymbolic", "-Xlinker", "--exclude-libs=ALL"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def engine_sources():
    inc = os.path.join(os.path.dirname(HERE), "include")
    return [os.path.join(CSRC, f) for f in ("engine.cu", "shade.cuh", "trace.cuh", "tracegroup.cuh", "treebuild.cuh", "animate.cuh", "vec.cuh")] + \
           [os.path.join(inc, f) for f in ("solr_b200.h", "solr_b200_types.h")]


def build_engine(force=False, verbose=False, extra=()):
    src = engine_sources()
    if force or _stale(ENGINE_LIB, src):
        # SOLR_B200_NVCC_FLAGS: extra flags for experimental builds, e.g. "-DUW_GROUP=0" (one ray per lane instead of the group walk)
        cmd = [NVCC] + NVCC_FLAGS + list(extra) + os.environ.get("SOLR_B200_NVCC_FLAGS", "").split() + \
              (["-Xptxas", "-v"] if verbose else []) + ["-o", ENGINE_LIB, src[0]]
        subprocess.check_call(cmd)
    return ENGINE_LIB


def build_host(force=False):
    src = [os.path.join(CSRC, "scene_host.cpp"), os.path.join(CSRC, "scene_host.h"), os.path.join(CSRC, "loaders.cpp")]
    if not os.path.exists(src[0]):
        return None
    if force or _stale(HOST_LIB, src):
        build_engine()
        subprocess.check_call([CXX, "-std=c++17", "-O2", "-fPIC", "-shared", "-pthread", "-Wall", "-Wno-sign-compare", "-ffp-contract=off", "-o", HOST_LIB, src[0], src[2],
                               "-L" + CSRC, "-lsolr_b200", "-Wl,-rpath,$ORIGIN"])
    return HOST_LIB


def build_all(force=False, verbose=False):
    return build_engine(force, verbose), build_host(force)


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
