"""In-tree build of the C-ABI library (libedf_b200.so) with nvcc for sm_100a.

    python -m elasticdeform_b200.build [--force] [--verbose]

The library is built next to this file so that it travels with the source tree
(the GPU box receives built .so files; a JIT cache under ~/.cache would not).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_NAME = "libedf_b200.so"
LIB_PATH = os.path.join(HERE, LIB_NAME)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "--split-compile", "0",          # ptxas of the (many) kernel instantiations in parallel: 4 min -> 1 min
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def _sources():
    out = []
    for root, _, files in os.walk(CSRC):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h")):
                out.append(os.path.join(root, f))
    out.append(os.path.join(os.path.dirname(HERE), "include", "edf_b200.h"))
    return out


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in _sources() if os.path.exists(s))


def build_library(force=False, verbose=False, extra_flags=(), out=None):
    """Compile csrc/edf_api.cu -> libedf_b200.so (or ``out``). Returns the library path."""
    if out is not None:
        cmd = [_nvcc(), *NVCC_FLAGS, *extra_flags, "-o", out, os.path.join(CSRC, "edf_api.cu")]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        return out
    if not force and not needs_build():
        return LIB_PATH
    nvcc = _nvcc()
    if nvcc is None:
        if os.path.exists(LIB_PATH):
            return LIB_PATH          # prebuilt library, no compiler on this box
        raise RuntimeError("nvcc not found and no prebuilt %s" % LIB_PATH)
    cmd = [nvcc, *NVCC_FLAGS, *extra_flags, "-o", LIB_PATH, os.path.join(CSRC, "edf_api.cu")]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    # python -m elasticdeform_b200.build [--force] [--ptxas] [--out PATH] [-DNAME=VALUE ...]
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    defs = [a for a in sys.argv[1:] if a.startswith("-D")]
    p = build_library(force="--force" in sys.argv, verbose=True, out=out,
                      extra_flags=(["-Xptxas", "-v"] if "--ptxas" in sys.argv else []) + defs)
    print("built", p)
