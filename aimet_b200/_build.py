"""Builds the CUDA library of the hot path in-tree: aimet_b200/lib/libaimet_b200.so (sm_100a only).

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container; the built .so travels to the GPU box
with the repo snapshot. No JIT at import time: the Python package only dlopens the file this script produces.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.environ.get("AIMET_B200_LIB") or os.path.join(LIB_DIR, "libaimet_b200.so")   # override: kernel-variant experiments
SOURCES = ["qdq.cu", "stats.cu", "tfe_search.cu", "learned_grid.cu", "broadcast.cu", "packed.cu", "qc_op.cu", "entropy.cu"]
HEADERS = ["common.cuh", "encoding_math.h", "tfe_math.h", "percentile_math.h", "mse_math.h", "entropy_math.h", os.path.join("..", "..", "include", "aimet_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--fmad=false",                       # no FMA contraction: results are defined operation by operation
    "-Xcompiler", "-fPIC,-ffp-contract=off",
    "-Xptxas", "-v",
    "--shared", "-cudart", "shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the aimet_b200 CUDA library cannot be built")


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > built for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into one shared library. Returns its path."""
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    extra = os.environ.get("AB_NVCC_EXTRA", "").split()   # kernel-variant experiments, together with AIMET_B200_LIB
    cmd = [_nvcc()] + NVCC_FLAGS + extra + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB_PATH]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libaimet_b200.so")
    with open(os.path.join(LIB_DIR, "ptxas_info.txt"), "w") as f:
        f.write(res.stdout)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
