"""Build libgraphem_b200.so in-tree with nvcc for sm_100a (no torch C++ dependency)."""
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
SRC = os.path.join(PKG_DIR, "csrc", "graphem_b200.cu")
HDR = os.path.join(ROOT, "include", "graphem_b200.h")
LIB = os.path.join(PKG_DIR, "libgraphem_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",      # REQUIRED for bit parity: only explicit fmaf() may fuse
    "-shared", "-Xcompiler", "-fPIC",
]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return t < os.path.getmtime(SRC) or t < os.path.getmtime(HDR)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libgraphem_b200.so")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-I", os.path.join(ROOT, "include"), SRC, "-o", LIB]
    subprocess.check_call(cmd)
    return LIB


DIAG_LIB = os.path.join(PKG_DIR, "libgraphem_b200_diag.so")


def build_diag(level: int = 1) -> str:
    """Diagnostic build (-DGEM_SCAN_DIAG=level): the scan / preparation / select kernels record %globaltimer points
    of their phases (scripts/scan_diag.py).  A separate file: the product library is never replaced by it."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build the diagnostic library")
    subprocess.check_call([nvcc] + NVCC_FLAGS + [f"-DGEM_SCAN_DIAG={int(level)}", "-I", os.path.join(ROOT, "include"), SRC,
                                                 "-o", DIAG_LIB])
    return DIAG_LIB


if __name__ == "__main__":
    import sys
    if "--diag" in sys.argv:
        print(build_diag())
    else:
        print(build(force=True, verbose=True))
