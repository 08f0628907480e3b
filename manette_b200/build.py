"""Builds manette_b200/libmanette_b200.so (the CUDA kernels + C ABI) in-tree with nvcc for sm_100a only."""
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libmanette_b200.so")
SOURCES = [os.path.join(_PKG, "csrc", f) for f in ("pool.cu",)]
HEADERS = [os.path.join(_PKG, "csrc", f) for f in ("emu_core.cuh", "cpu_defs.h", "atari_env.cuh", "decode_tables.h", "game_db.h")] + \
          [os.path.join(os.path.dirname(_PKG), "include", "manette_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile if the library is missing or older than its sources.  Returns the library path."""
    if force or stale():
        # MN_BUILD_DEFS="-DMN_CHECK -DMN_FILL_NOINLINE": diagnostic builds (tools/gpu_check_build.sh)
        extra = os.environ.get("MN_BUILD_DEFS", "").split()
        cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
        subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))


def source_hash():
    """sha256 (first 16 hex digits) over the CUDA sources and headers the library is built from: stamps ncu-derived
    counters under profiles/ so that bench.py can tell whether they describe the kernels it is running."""
    import hashlib
    h = hashlib.sha256()
    for path in sorted(SOURCES + HEADERS):
        with open(path, "rb") as f:
            h.update(os.path.basename(path).encode() + b"\0" + f.read() + b"\0")
    return h.hexdigest()[:16]
