"""In-tree build of the sm_100a shared library (``pctrans_b200/_build/libpctrans_msda.so``).

Plain ``nvcc`` (no torch cpp_extension, no libtorch link): the library is a C-ABI plugin that any
host language can load.  ``nvcc`` cross-compiles without a GPU, so this runs in the authoring
container; the built ``.so`` is git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OUT_DIR = os.path.join(PKG_DIR, "_build")
LIB_PATH = os.path.join(OUT_DIR, "libpctrans_msda.so")
SOURCES = ["msda_capi.cu", "msda_host.cu", "linear_tc.cu"]
DEPS = ["msda_capi.cu", "msda_host.cu", "linear_tc.cu", "msda_kernels.cuh",
        os.path.join("..", "..", "include", "pct_msda.h"), os.path.join("..", "..", "include", "pct_linear.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
    "--cudart", "shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libpctrans_msda.so")


def _source_hash(extra: str = "") -> str:
    import hashlib
    h = hashlib.sha256(extra.encode())
    for d in DEPS:
        with open(os.path.join(CSRC, d), "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _is_current(lib_path: str, extra: str = "") -> bool:
    """A library is current when the hash of the sources it was built from (stored next to it)
    matches the tree -- file mtimes do not survive the snapshot copy to the GPU box."""
    try:
        with open(lib_path + ".srchash") as f:
            return os.path.exists(lib_path) and f.read().strip() == _source_hash(extra)
    except OSError:
        return False


def _stamp(lib_path: str, extra: str = "") -> None:
    with open(lib_path + ".srchash", "w") as f:
        f.write(_source_hash(extra))


def needs_build() -> bool:
    return not _is_current(LIB_PATH)


DEBUG_LIB_PATH = os.path.join(OUT_DIR, "libpctrans_msda_dbg.so")


def build_library(force: bool = False, verbose: bool = False, debug: bool = False) -> str:
    """Compile the CUDA sources for sm_100a.  Returns the path of the shared library.
    ``debug=True`` builds libpctrans_msda_dbg.so with -DPCT_MSDA_BOUNDS_CHECK (tap addresses
    checked in-kernel; used by tests/test_gpu_bounds.py through PCTRANS_MSDA_LIB)."""
    if debug:
        os.makedirs(OUT_DIR, exist_ok=True)
        if not force and _is_current(DEBUG_LIB_PATH, "debug"):
            return DEBUG_LIB_PATH
        cmd = [_nvcc()] + NVCC_FLAGS + ["-DPCT_MSDA_BOUNDS_CHECK", "-o", DEBUG_LIB_PATH] + \
            [os.path.join(CSRC, s) for s in SOURCES] + ["-ccbin", "/usr/bin/g++"]
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if proc.returncode != 0:
            sys.stderr.write(proc.stdout)
            raise RuntimeError("nvcc failed building the bounds-check library")
        _stamp(DEBUG_LIB_PATH, "debug")
        return DEBUG_LIB_PATH
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    # the image exports CC/CXX wrappers that nvcc must not pick up as host compiler
    proc = subprocess.run(cmd + ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else cmd,
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    log_path = os.path.join(OUT_DIR, "build.log")
    with open(log_path, "w") as f:
        f.write(" ".join(cmd) + "\n" + proc.stdout)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed (exit {proc.returncode}); see {log_path}")
    _stamp(LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
