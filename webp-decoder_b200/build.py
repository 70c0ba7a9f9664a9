"""Builds libvp8gpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB = HERE / "libvp8gpu.so"
SOURCES = [HERE / "csrc" / "vp8_kernels.cu", HERE / "csrc" / "vp8_pairs.cu", HERE / "csrc" / "vp8_gpu.cu", HERE / "csrc" / "vp8_parse.cpp"]
HEADERS = [HERE / "csrc" / "vp8_dev.h", HERE / "csrc" / "vp8_common.cuh", HERE / "csrc" / "vp8_lf2.cuh", HERE / "csrc" / "vp8_tables.h",
           HERE / "csrc" / "vp8_pairs_image.inc", HERE / "csrc" / "vp8_pairs_row.inc", HERE / "csrc" / "vp8_pairs_step_a.inc", HERE / "csrc" / "vp8_pairs_step_b.inc", HERE.parent / "include" / "vp8_gpu.h", HERE.parent / "include" / "vp8_abi.h",
           HERE.parent / "include" / "vp8_parse.h"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libvp8gpu.so cannot be built (there is no CPU fallback)")


def stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.exists() and p.stat().st_mtime > t for p in SOURCES + HEADERS)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    srcs = [s for s in SOURCES if s.exists()]
    if not force and not stale():
        return LIB
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC,-O3,-pthread", "-shared", "-o", str(LIB)] + [str(s) for s in srcs]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True, cwd=str(HERE))
    # batch CLI front end, linked against the library it sits next to
    cli = HERE / "csrc" / "vp8gpu_batch.cpp"
    if cli.exists():
        subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", str(HERE / "vp8gpu_batch"), str(cli), "-L" + str(HERE), "-lvp8gpu",
                        "-Wl,-rpath,$ORIGIN"], check=True, cwd=str(HERE))
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
