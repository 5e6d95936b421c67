"""In-tree build of libmaexp_b200.so (hand-written sm_100a CUDA behind a C ABI).

    python -m maexp_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU; the .so is git-ignored but travels to the
GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.environ.get("MAEXP_LIB") or os.path.join(HERE, "libmaexp_b200.so")  # MAEXP_LIB: experiment builds (tools/)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = os.environ.get("MAEXP_EXTRA_NVCC", "").split() + ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--fmad=false",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-I", INCLUDE]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    return deps


def build(force=False, verbose=False):
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in _deps()):
        return LIB
    from concurrent.futures import ThreadPoolExecutor

    def compile_one(src):
        obj = os.path.join(os.environ.get("MAEXP_OBJ_DIR") or CSRC, os.path.basename(src)[:-3] + ".o")
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
        objs = list(pool.map(compile_one, sources()))
    subprocess.check_call([NVCC, "-shared", "-o", LIB] + objs + ["-lcudart"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
