"""Build recipe of libmbgc_b200.so (CUDA kernels + C ABI), sm_100a only, in-tree.

    python -m mbgc_b200.build            # build if sources are newer than the library
    python -m mbgc_b200.build --force

nvcc cross-compiles without a GPU.  The shared object is git-ignored but travels with the tree.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmbgc_b200.so")
SOURCES = ["kernels.cu", "pack.cu", "ingest.cu", "coders.cu", "coders.cpp", "rcmatch.cu", "rcmatch.cpp", "decode.cu", "decode.cpp", "matcher.cpp", "packer.cpp", "ingest.cpp", "collection.cpp", "capi.cpp"]
HOST_SOURCES = ["hostpack.cpp"]  # plain g++ (AVX2 via function targets), linked into the same library
HEADERS = ["kernels.h", "pack.h", "ingest.h", "coders.h", "rcmatch.h", "decode.h", "decode.hpp", "matcher.hpp", "collection.hpp", "hostpack.hpp", os.path.join("..", "..", "include", "mbgc_b200.h")]
CXX = os.environ.get("CXX_HOST", "/usr/bin/g++")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "-shared", "-cudart", "static"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HOST_SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    for f in HOST_SOURCES:
        obj = os.path.join(CSRC, f.replace(".cpp", ".o"))
        res = subprocess.run([CXX, "-O3", "-std=c++17", "-fPIC", "-Wall", "-pthread", "-c", os.path.join(CSRC, f), "-o", obj],
                             capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError(f"g++ failed building {f}")
        objs.append(obj)
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(CSRC, f) for f in SOURCES] + objs + ["-Xlinker", "-lpthread", "-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libmbgc_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
