"""Builds the engine's CUDA sources into ``anatomix_b200/lib/libanatomix_b200.so``.

Explicit nvcc for sm_100a (``-gencode arch=compute_100a,code=sm_100a``; the
``-arch=sm_100a`` spelling drops the ``a`` features in this toolchain), in-tree so
the library travels with the repository snapshot.  No torch headers are needed:
the boundary is the plain C ABI of ``include/anatomix_b200.h``.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(PKG, "csrc")
# experiments: ANX_LIB_VARIANT=<tag> builds lib/libanatomix_b200_<tag>.so with -DANX_EXPERIMENTS (the ANX_ABLATE /
# ANX_NO_* / ANX_ROWS / ANX_X_LEAD timing switches, which skip work and produce wrong results, exist only in
# such builds) plus the extra nvcc flags in ANX_BUILD_DEFS (e.g. -DANX_EPI_WARPS=16); _lib.py loads the same
# variant when the variable is set.  The product library (no variant) ignores all of those variables.
VARIANT = os.environ.get("ANX_LIB_VARIANT", "")
OBJ = os.path.join(PKG, "lib", "obj" + ("_" + VARIANT if VARIANT else ""))
LIB = os.path.join(PKG, "lib", "libanatomix_b200" + ("_" + VARIANT if VARIANT else "") + ".so")
SOURCES = ["engine.cu", "selftest.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc():
    for cand in (os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile if anything under csrc/ or include/ is newer than the library."""
    headers = [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(PKG, "..", "include", "anatomix_b200.h"))
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    jobs = []
    for name in SOURCES:
        src = os.path.join(SRC, name)
        obj = os.path.join(OBJ, name.replace(".cu", ".o"))
        if force or _stale(obj, [src] + headers):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        defs = (["-DANX_EXPERIMENTS"] if VARIANT else []) + os.environ.get("ANX_BUILD_DEFS", "").split()
        r = subprocess.run([nvcc, *NVCC_FLAGS, *defs, "-c", src, "-o", obj],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return src, r.stderr

    with ThreadPoolExecutor(max_workers=4) as ex:
        for src, log in ex.map(compile_one, jobs):
            if verbose:
                print(f"--- {os.path.basename(src)}\n{log}")
    objs = [os.path.join(OBJ, n.replace(".cu", ".o")) for n in SOURCES]
    if force or jobs or _stale(LIB, objs):
        r = subprocess.run([nvcc, "-shared", "-o", LIB, *objs, "-lcudart"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
