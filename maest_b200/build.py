"""In-tree build of libmaest_b200.so (sm_100a only) with nvcc; no JIT cache, no torch extension machinery."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libmaest_b200.so")
SOURCES = ["api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "maest_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


SAFE_LIB_PATH = os.path.join(LIB_DIR, "libmaest_b200_safe.so")


def build_variant(name: str, defines: list[str]) -> str:
    """Experiment build: libmaest_b200_<name>.so with extra -D flags (A/B timing on the GPU box; select with MAEST_B200_LIB)."""
    out = os.path.join(LIB_DIR, f"libmaest_b200_{name}.so")
    os.makedirs(LIB_DIR, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")] + [f"-D{d}" for d in defines]
    cmd = [_nvcc(), *flags, "-o", out, *[os.path.join(CSRC, s) for s in SOURCES], "-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return out


def build(force: bool = False, verbose: bool = False, safe: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into maest_b200/lib/libmaest_b200.so; returns the path.
    safe=True builds libmaest_b200_safe.so with -DMB_SAFE_WAIT (mbarrier waits trap after 2 s instead of hanging):
    the bring-up build for new kernels, selected at run time with MAEST_B200_LIB=<path>."""
    out = SAFE_LIB_PATH if safe else LIB_PATH
    if not force and not safe and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")] + (["-DMB_SAFE_WAIT"] if safe else [])
    cmd = [_nvcc(), *flags, "-o", out, *[os.path.join(CSRC, s) for s in SOURCES], "-lcudart"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose=True))
