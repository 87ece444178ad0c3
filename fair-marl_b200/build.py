"""In-tree build of ``libfairmarl.so`` with nvcc for sm_100a (no JIT cache, no torch linkage)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
SOURCES = ["fm_kernels.cu", "fm_aw.cu", "fm_formation.cu", "fm_abi.cu"]
HEADERS = ["fm_device.cuh", "fm_launch.h", "fm_small.cuh", os.path.join("..", "..", "include", "fairmarl.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def library_path() -> str:
    return os.path.join(_HERE, "libfairmarl.so")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found: fair_marl_b200 needs the CUDA toolkit to build libfairmarl.so")


def _stale() -> bool:
    lib = library_path()
    if not os.path.isfile(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(os.path.join(_CSRC, f)) > t for f in SOURCES + HEADERS)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources if the library is missing or older than them.  Returns its path."""
    if not force and not _stale():
        return library_path()
    extra = os.environ.get("FM_NVCC_EXTRA", "").split()      # diagnostics only, e.g. FM_NVCC_EXTRA="-G" for compute-sanitizer runs
    cmd = [_nvcc()] + NVCC_FLAGS + extra + [os.path.join(_CSRC, s) for s in SOURCES] + ["-o", library_path()]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return library_path()
