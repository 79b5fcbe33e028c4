"""Build recipe for libdomainrag_b200.so (in-tree, sm_100a only).

`python -m domain_rag_b200.build` or `__graft_entry__.build()` compiles every .cu under csrc/ with
nvcc (cross-compiles without a GPU) and links them into one C-ABI shared library next to this
file. Objects are cached by source mtime so incremental rebuilds take seconds.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
OBJ_DIR = PKG_DIR / "_build"
LIB_PATH = PKG_DIR / "libdomainrag_b200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo", "-DDRAG_NO_FAST_MATH",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "--expt-relaxed-constexpr",
]


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _headers_mtime() -> float:
    hdrs = list(CSRC.glob("*.cuh")) + list((PKG_DIR.parent / "include").glob("*.h"))
    return max((h.stat().st_mtime for h in hdrs), default=0.0)


def _compile_one(src: Path, verbose: bool, ptxas_v: bool) -> Path:
    obj = OBJ_DIR / (src.stem + ".o")
    newest = max(src.stat().st_mtime, _headers_mtime())
    if obj.exists() and obj.stat().st_mtime > newest:
        return obj
    cmd = [NVCC, *ARCH_FLAGS, *COMMON_FLAGS, "-c", str(src), "-o", str(obj)]
    if ptxas_v:
        cmd[1:1] = ["-Xptxas", "-v"]
    if verbose:
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")
    if verbose or ptxas_v:
        sys.stdout.write(res.stdout)
        sys.stderr.write(res.stderr)
    return obj


def build(verbose: bool = False, ptxas_v: bool = False, force: bool = False) -> Path:
    """Compile and link libdomainrag_b200.so; returns its path."""
    OBJ_DIR.mkdir(exist_ok=True)
    if force:
        for o in OBJ_DIR.glob("*.o"):
            o.unlink()
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile_one(s, verbose, ptxas_v), srcs))
    if (not LIB_PATH.exists()) or any(o.stat().st_mtime > LIB_PATH.stat().st_mtime for o in objs):
        cmd = [NVCC, *ARCH_FLAGS, "-shared", "-cudart", "static", "-o", str(LIB_PATH), *map(str, objs)]
        if verbose:
            print(" ".join(cmd), flush=True)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    p = build(verbose=True, ptxas_v="--ptxas" in sys.argv, force="--force" in sys.argv)
    print("built", p)
