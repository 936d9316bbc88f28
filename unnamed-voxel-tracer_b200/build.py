"""Build the in-tree shared library (CUDA kernels + C ABI + host formats) for sm_100a.

nvcc cross-compiles without a GPU.  The .so stays in-tree (git-ignored) so it travels to
the GPU box with the snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libuvt.so")

SOURCES = [
    os.path.join(CSRC, "uvt.cu"),
    os.path.join(CSRC, "group.cpp"),
    os.path.join(CSRC, "host", "noise.cpp"),
    os.path.join(CSRC, "host", "world.cpp"),
    os.path.join(CSRC, "host", "vox.cpp"),
    os.path.join(CSRC, "host", "atlas.cpp"),
    os.path.join(CSRC, "host", "camera.cpp"),
]
import glob

# everything a source may include: an edit to any of them makes the library stale
HEADERS = sorted(glob.glob(os.path.join(CSRC, "**", "*.cuh"), recursive=True) + glob.glob(os.path.join(CSRC, "**", "*.h"), recursive=True) +
                 glob.glob(os.path.join(ROOT, "include", "*.h")))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",            # every fp32 op individually rounded: bit-exact parity with the oracle
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-fno-fast-math,-Wall",
    "-shared", "-cudart", "static",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS + [os.path.abspath(__file__)])


def build(force=False, verbose=False, out=None, defines=()):
    """out/defines: build an experimental variant (e.g. -DUVT_WARP_W=4) next to the product library."""
    if out is None and not force and not is_stale():
        return LIB
    out = out or LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + SOURCES
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc failed building libuvt.so")
    if verbose:
        print(r.stdout)
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
