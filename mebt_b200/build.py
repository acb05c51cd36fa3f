"""In-tree nvcc build of the mebt_b200 C-ABI shared library (sm_100a only).

The library is a plain `extern "C"` .so with no torch types in its interface; Python binds it with
ctypes (`mebt_b200/_lib.py`).  `python -m mebt_b200.build` rebuilds when a source is newer than the
.so.  nvcc cross-compiles without a GPU, so this also runs in the CPU-only build container.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIB = ROOT / "libmebt_b200.so"
INCLUDE = ROOT.parent / "include"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
]


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _newest_src_mtime() -> float:
    files = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h"))
    return max(f.stat().st_mtime for f in files)


def needs_build() -> bool:
    return (not LIB.exists()) or LIB.stat().st_mtime < _newest_src_mtime()


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    objdir = ROOT / "build"
    objdir.mkdir(exist_ok=True)
    procs = []
    objs = []
    for src in sources():
        obj = objdir / (src.stem + ".o")
        objs.append(obj)
        deps = [src] + list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h"))
        if not force and obj.exists() and obj.stat().st_mtime >= max(d.stat().st_mtime for d in deps):
            continue
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(INCLUDE), "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed for {src.name}:\n{out}\n")
        elif verbose and out:
            print(out)
    if failed:
        raise RuntimeError("mebt_b200: nvcc build failed")
    cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"mebt_b200: link failed:\n{r.stdout}")
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
