"""Builds biblbm_b200/libbiblbm_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

-fmad=false: no multiply-add contraction, so the kernels perform exactly the IEEE double operations the reference's
x86-64 build performs and results can be compared bit for bit (the LBM kernel is HBM-bound).
"""
from __future__ import annotations

import os
import pathlib
import shutil
import subprocess
import sys

HERE = pathlib.Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libbiblbm_b200.so"
SOURCES = ["api.cu", "lbm_kernels.cu", "ib_kernels.cu", "ib_multislab.cu", "comm.cu", "group.cu"]
HEADERS = [CSRC / "domain.cuh", HERE.parent / "include" / "biblbm_b200.h"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def nccl_include_flags() -> list[str]:
    if os.path.exists("/usr/include/nccl.h"):
        return []
    try:  # the torch-bundled NCCL ships its header too
        import nvidia.nccl
        inc = pathlib.Path(list(nvidia.nccl.__path__)[0]) / "include"
        if (inc / "nccl.h").exists():
            return [f"-I{inc}"]
    except Exception:
        pass
    return []


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in [CSRC / s for s in SOURCES] + HEADERS)


def build(force: bool = False, verbose: bool = False, fmad: bool = False, out: pathlib.Path | None = None, defines: tuple = ()) -> pathlib.Path:
    if out is None and not force and not needs_build():
        return LIB
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=true" if fmad else "-fmad=false",
           "-Xcompiler", "-fPIC", "-shared", *[f"-D{x}" for x in defines], *nccl_include_flags(), "-o", str(out or LIB)] + [str(CSRC / s) for s in SOURCES] + ["-ldl", "-lpthread", "-Xlinker", "--no-undefined"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout, r.stderr, file=sys.stderr)
    return out or LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
