"""Builds the CUDA library in-tree (hestonexotics_b200/lib/libhexo_gpu.so).

nvcc cross-compiles for sm_100a without a GPU.  The .so is git-ignored but
travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libhexo_gpu.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "-Xfatbin", "-compress-all",   # 96 kernel instantiations with line info: 45 MB -> 15 MB
]
OBJ_DIR = os.path.join(LIB_DIR, "obj")


def _sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def _deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "hexo_gpu.h"))
    return deps


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    """Compile every .cu under csrc/ (one nvcc per file, in parallel) and link them into one
    shared library.  Returns its path.

    `defines` / `out` build a VARIANT next to the product library (development only): e.g.
    defines=("HEXO_DEV_PROBES",), out="libhexo_gpu_dev.so" compiles the probes the shipped
    library does not contain (HEXO_NO_REFILL, HEXO_BLOCK); tools select it with HEXO_GPU_LIB.
    The product library is always built without defines."""
    if out is not None or defines:
        return _build_variant(tuple(defines), out or "libhexo_gpu_variant.so", verbose)
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libhexo_gpu.so")
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = _sources()
    objs = [os.path.join(OBJ_DIR, os.path.basename(s)[:-3] + ".o") for s in srcs]

    def compile_one(pair):
        src, obj = pair
        return subprocess.run([nvcc, *NVCC_FLAGS, "-c", "-o", obj, src], capture_output=True,
                              text=True)

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, zip(srcs, objs)))
    log = "".join(r.stdout + r.stderr for r in results)
    failed = [s for s, r in zip(srcs, results) if r.returncode != 0]
    if not failed:
        tmp = LIB_PATH + ".tmp"
        link = subprocess.run([nvcc, "-shared", "-o", tmp, *objs], capture_output=True, text=True)
        log += link.stdout + link.stderr
        if link.returncode != 0:
            failed = ["link"]
    if verbose or failed:
        sys.stderr.write(log)
    if failed:
        raise RuntimeError(f"nvcc failed building libhexo_gpu.so ({failed}; see stderr)")
    os.replace(tmp, LIB_PATH)
    with open(os.path.join(LIB_DIR, "ptxas.log"), "w") as f:
        f.write(log)
    return LIB_PATH


def _build_variant(defines, out, verbose):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    tag = os.path.splitext(out)[0]
    obj_dir = os.path.join(LIB_DIR, "obj_" + tag)
    os.makedirs(obj_dir, exist_ok=True)
    srcs = _sources()
    objs = [os.path.join(obj_dir, os.path.basename(s)[:-3] + ".o") for s in srcs]
    flags = NVCC_FLAGS + ["-D" + d for d in defines]

    def compile_one(pair):
        return subprocess.run([nvcc, *flags, "-c", "-o", pair[1], pair[0]], capture_output=True,
                              text=True)

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, zip(srcs, objs)))
    log = "".join(r.stdout + r.stderr for r in results)
    path = os.path.join(LIB_DIR, out)
    bad = [s for s, r in zip(srcs, results) if r.returncode != 0]
    if not bad:
        link = subprocess.run([nvcc, "-shared", "-o", path, *objs], capture_output=True, text=True)
        log += link.stdout + link.stderr
        if link.returncode != 0:
            bad = ["link"]
    if verbose or bad:
        sys.stderr.write(log)
    if bad:
        raise RuntimeError(f"nvcc failed building {out} ({bad})")
    with open(os.path.join(LIB_DIR, tag + ".ptxas.log"), "w") as f:
        f.write(log)
    shutil.rmtree(obj_dir, ignore_errors=True)
    return path


if __name__ == "__main__":
    if "--dev" in sys.argv:   # the library with the development probes compiled in
        print(build(defines=("HEXO_DEV_PROBES",), out="libhexo_gpu_dev.so", verbose=True))
    else:
        print(build(force="--force" in sys.argv, verbose=True))
