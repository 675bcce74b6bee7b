"""Build the sm_100a shared library `hicdiff_b200/lib/libhicdiff_b200.so` in-tree with nvcc.

The library has a plain C ABI (`include/hicdiff_b200.h`) and no torch dependency; Python binds it with ctypes.
nvcc cross-compiles without a GPU, so this also runs in CPU-only containers.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIBDIR = ROOT / "lib"
OBJDIR = ROOT / "build"
LIB = LIBDIR / "libhicdiff_b200.so"

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; hicdiff_b200 has no prebuilt or CPU fallback path")


def _stamp(src: Path) -> str:
    h = hashlib.sha256()
    h.update(src.read_bytes())
    for hdr in sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh")) + [ROOT.parent / "include" / "hicdiff_b200.h"]:
        h.update(hdr.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    nvcc = _nvcc()
    OBJDIR.mkdir(exist_ok=True)
    LIBDIR.mkdir(exist_ok=True)
    sources = sorted(CSRC.glob("*.cu"))
    todo = []
    for src in sources:
        obj = OBJDIR / (src.stem + ".o")
        stamp = OBJDIR / (src.stem + ".stamp")
        want = _stamp(src)
        if force or not obj.exists() or not stamp.exists() or stamp.read_text() != want:
            todo.append((src, obj, stamp, want))

    def compile_one(item):
        src, obj, stamp, want = item
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        stamp.write_text(want)
        return src.name, r.stderr

    if todo:
        with ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            for name, log in ex.map(compile_one, todo):
                if verbose and log:
                    print(f"--- {name}\n{log}", file=sys.stderr)
    if todo or not LIB.exists():
        objs = [str(OBJDIR / (s.stem + ".o")) for s in sources]
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
