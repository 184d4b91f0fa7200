"""Build libgrafp_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Every ``csrc/*.cu`` is compiled to its own object (in parallel, rebuilt only when the source or a
header is newer) and the objects are linked into one shared library.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB_PATH = os.path.join(LIB_DIR, "libgrafp_b200.so")
HEADER = os.path.join(os.path.dirname(PKG_DIR), "include", "grafp_b200.h")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + [HEADER]


def _obj_path(src: str) -> str:
    return os.path.join(OBJ_DIR, os.path.splitext(os.path.basename(src))[0] + ".o")


def _newer(path: str, deps) -> bool:
    """True when `path` is missing or older than any of `deps`."""
    if not os.path.isfile(path):
        return True
    built = os.path.getmtime(path)
    return any(os.path.getmtime(d) > built for d in deps)


def _stale() -> bool:
    return _newer(LIB_PATH, sources() + _headers())


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source into the shared library; returns its path."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libgrafp_b200.so cannot be built")
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdrs = _headers()
    srcs = sources()
    todo = [s for s in srcs if force or _newer(_obj_path(s), [s] + hdrs)]

    def compile_one(src):
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", _obj_path(src)]
        return src, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 1) or 1) as pool:
        results = list(pool.map(compile_one, todo))
    failed = False
    for src, res in results:
        if verbose or res.returncode != 0:
            sys.stderr.write(f"--- {os.path.basename(src)}\n" + res.stdout + res.stderr)
        failed = failed or res.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libgrafp_b200.so")
    # drop objects whose source is gone
    keep = {_obj_path(s) for s in srcs}
    for o in glob.glob(os.path.join(OBJ_DIR, "*.o")):
        if o not in keep:
            os.remove(o)
    tmp = LIB_PATH + ".tmp"
    # libcuda is NOT linked: cuTensorMapEncodeTiled is resolved at run time through
    # cudaGetDriverEntryPoint, so the library links and loads on a box without the driver.
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + sorted(keep) + ["-o", tmp]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libgrafp_b200.so")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
