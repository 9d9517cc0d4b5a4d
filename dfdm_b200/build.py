"""
Build ``dfdm_b200/libdfdm_b200.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m dfdm_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdfdm_b200.so")
SOURCES = ["plan.cpp", "api.cu", "direct.cu", "pcg.cu", "peer.cu"]
HEADERS = [os.path.join(ROOT, "include", "dfdm_b200.h"), os.path.join(CSRC, "plan.hpp"), os.path.join(CSRC, "common.cuh")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-cudart", "static"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """``defines`` / ``out`` build a tuning variant of the library elsewhere (tools/ only); the product is LIB."""
    target = out or LIB
    if out is None and not force and not stale():
        return LIB
    objdir = CSRC if out is None else os.path.dirname(os.path.abspath(out))
    os.makedirs(objdir, exist_ok=True)
    tag = "" if out is None else "." + os.path.splitext(os.path.basename(out))[0]
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, os.path.splitext(src)[0] + tag + ".o")
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-D" + d for d in defines] + \
              ["-I", os.path.join(ROOT, "include"), "-I", CSRC, "-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj]
        subprocess.run(cmd, check=True)
        objs.append(obj)
    cmd = [_nvcc()] + ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o", target] + objs
    subprocess.run(cmd, check=True)
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
