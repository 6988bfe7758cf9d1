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
SOURCES = ["msda_capi.cu", "msda_inst_fwd.cu", "msda_inst_bwd.cu", "msda_inst_binned.cu",
           "msda_host.cu", "linear_tc.cu", "layernorm_bwd.cu"]
HEADERS = ["msda_kernels.cuh", "msda_bwd_binned.cuh", "msda_launch.h",
           os.path.join("..", "..", "include", "pct_msda.h"),
           os.path.join("..", "..", "include", "pct_linear.h")]
DEPS = SOURCES + HEADERS
# headers a translation unit does NOT include (so editing them does not rebuild it)
NOT_INCLUDED = {
    "msda_inst_fwd.cu": ["msda_bwd_binned.cuh", "pct_linear.h"],
    "msda_inst_bwd.cu": ["msda_bwd_binned.cuh", "pct_linear.h"],
    "msda_inst_binned.cu": ["pct_linear.h"],
    "msda_host.cu": ["msda_kernels.cuh", "msda_bwd_binned.cuh", "msda_launch.h", "pct_linear.h"],
    "linear_tc.cu": ["msda_kernels.cuh", "msda_bwd_binned.cuh", "msda_launch.h"],
    "layernorm_bwd.cu": ["msda_kernels.cuh", "msda_bwd_binned.cuh", "msda_launch.h"],
}
# every translation unit is compiled on its own (in parallel, objects cached by source hash) and
# linked into one shared library
COMPILE_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "--cudart", "shared"]
NVCC_FLAGS = COMPILE_FLAGS + LINK_FLAGS


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


def _ccbin() -> list:
    # the image exports CC/CXX wrappers that nvcc must not pick up as host compiler
    return ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []


def _compile_one(src: str, obj_dir: str, defines: list, force: bool):
    """Compile one translation unit to an object file; skipped when the object was built from the
    same bytes (sources + headers + flags).  Returns (object path, compiler output)."""
    import hashlib
    h = hashlib.sha256((" ".join(COMPILE_FLAGS + defines)).encode())
    for d in [src] + [h_ for h_ in HEADERS if os.path.basename(h_) not in NOT_INCLUDED.get(src, [])]:
        with open(os.path.join(CSRC, d), "rb") as f:
            h.update(f.read())
    obj = os.path.join(obj_dir, os.path.splitext(src)[0] + ".o")
    stamp = obj + ".srchash"
    log = obj + ".log"
    try:
        with open(stamp) as f:
            if not force and os.path.exists(obj) and f.read().strip() == h.hexdigest():
                with open(log) as lf:
                    return obj, lf.read()
    except OSError:
        pass
    cmd = [_nvcc()] + COMPILE_FLAGS + defines + ["-c", "-o", obj, os.path.join(CSRC, src)] + _ccbin()
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    out = " ".join(cmd) + "\n" + proc.stdout
    if proc.returncode != 0:
        sys.stderr.write(out)
        raise RuntimeError(f"nvcc failed on {src} (exit {proc.returncode})")
    with open(log, "w") as f:
        f.write(out)
    with open(stamp, "w") as f:
        f.write(h.hexdigest())
    return obj, out


def _build(lib_path: str, obj_dir: str, defines: list, force: bool, verbose: bool, log_name: str) -> str:
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(obj_dir, exist_ok=True)
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        results = list(pool.map(lambda s: _compile_one(s, obj_dir, defines, force), SOURCES))
    cmd = [_nvcc()] + LINK_FLAGS + ["-o", lib_path] + [o for o, _ in results] + _ccbin()
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    text = "".join(out for _, out in results) + " ".join(cmd) + "\n" + proc.stdout
    with open(os.path.join(OUT_DIR, log_name), "w") as f:
        f.write(text)
    if verbose or proc.returncode != 0:
        sys.stderr.write(text)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc link failed (exit {proc.returncode}); see {log_name}")
    return lib_path


def build_library(force: bool = False, verbose: bool = False, debug: bool = False) -> str:
    """Compile the CUDA sources for sm_100a.  Returns the path of the shared library.
    ``debug=True`` builds libpctrans_msda_dbg.so with -DPCT_MSDA_BOUNDS_CHECK (tap addresses
    checked in-kernel; used by tests/test_gpu_bounds.py through PCTRANS_MSDA_LIB)."""
    os.makedirs(OUT_DIR, exist_ok=True)
    if debug:
        if not force and _is_current(DEBUG_LIB_PATH, "debug"):
            return DEBUG_LIB_PATH
        _build(DEBUG_LIB_PATH, os.path.join(OUT_DIR, "obj_dbg"), ["-DPCT_MSDA_BOUNDS_CHECK"], force,
               verbose, "build_dbg.log")
        _stamp(DEBUG_LIB_PATH, "debug")
        return DEBUG_LIB_PATH
    if not force and not needs_build():
        return LIB_PATH
    _build(LIB_PATH, os.path.join(OUT_DIR, "obj"), [], force, verbose, "build.log")
    _stamp(LIB_PATH)
    return LIB_PATH


def build_variant(name: str, defines: list, sources=("msda_inst_binned.cu",)) -> str:
    """Tuning builds: the listed translation units recompiled with extra -D flags and linked with the
    release objects into _build/variants/libpctrans_msda_<name>.so (select with PCTRANS_MSDA_LIB)."""
    build_library()
    vdir = os.path.join(OUT_DIR, "variants")
    odir = os.path.join(vdir, "obj_" + name)
    os.makedirs(odir, exist_ok=True)
    objs = []
    for src in SOURCES:
        if src in sources:
            objs.append(_compile_one(src, odir, list(defines), False)[0])
        else:
            objs.append(os.path.join(OUT_DIR, "obj", os.path.splitext(src)[0] + ".o"))
    lib = os.path.join(vdir, f"libpctrans_msda_{name}.so")
    cmd = [_nvcc()] + LINK_FLAGS + ["-o", lib] + objs + _ccbin()
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout)
        raise RuntimeError("variant link failed")
    return lib


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
