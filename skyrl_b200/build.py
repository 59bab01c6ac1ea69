"""Build recipe for libskyb200.so (sm_100a only, in-tree so the .so travels with the repo snapshot).

    python -m skyrl_b200.build          # incremental
    python -m skyrl_b200.build --force
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
INCLUDE = PKG_DIR.parent / "include"
OBJ_DIR = PKG_DIR / "build"
LIB_PATH = PKG_DIR / "libskyb200.so"

SOURCES = ["core.cu", "loss.cu", "advantage.cu", "offpolicy.cu", "covsel.cu", "value_head.cu", "head.cu"]
HEADERS = ["common.cuh", "ptx.cuh", "gemm_sm100.cuh", "internal.h", "loss_token.cuh", "tile_order.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    # "variable used before its value is set" is an error: it once let the compiler drop whole loss-type branches
    # of a kernel without a word at run time
    "--diag-error", "549",
]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; libskyb200.so cannot be built")
    return cand


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a and link libskyb200.so.  Returns the library path."""
    nvcc = _nvcc()
    OBJ_DIR.mkdir(exist_ok=True)
    hdr_deps = [CSRC / h for h in HEADERS] + [INCLUDE / "skyb200.h", Path(__file__)]
    jobs = []
    for src in SOURCES:
        obj = OBJ_DIR / (src[:-3] + ".o")
        if force or _stale(obj, [CSRC / src] + hdr_deps):
            jobs.append([nvcc, *NVCC_FLAGS, "-I", str(INCLUDE), "-c", str(CSRC / src), "-o", str(obj)])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")

    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(run, jobs))
    objs = [OBJ_DIR / (s[:-3] + ".o") for s in SOURCES]
    if force or jobs or _stale(LIB_PATH, objs):
        run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH), *map(str, objs)])
    return LIB_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print(p)
