"""In-tree build of libba_b200.so (sm_100a only) with plain nvcc.

``python -m pysfm_b200.build`` or ``__graft_entry__.build()``.  The shared library lands in
``pysfm_b200/csrc/`` (git-ignored, but shipped to the GPU box with the working tree).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_NAME = "libba_b200.so"
LIB_PATH = os.path.join(CSRC, LIB_NAME)
SOURCES = ["ba_api.cu", "ba_kernels.cu", "ba_solve.cu", "ba_comm.cu", "ba_pack.cu"]
HEADERS = ["ba_math.cuh", "ba_context.h", "ba_peer.cuh", "ba_solve_tc.cuh", os.path.join("..", "..", "include", "ba_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=true",           # FP64 contraction on, as on the host BLAS path of the reference
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found; libba_b200.so cannot be built")


def _stale():
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    """Compile every .cu of the package for sm_100a and link them into one shared object."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = _nvcc()
    objs = []
    log = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log.append(res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s\n%s" % (src, res.stdout, res.stderr))
        objs.append(obj)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (res.stdout, res.stderr))
    with open(os.path.join(CSRC, "ptxas_report.txt"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        sys.stderr.write("\n".join(log))
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
