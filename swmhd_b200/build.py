"""Build libswmhd_cuda.so in-tree with nvcc for sm_100a (no JIT, no torch extension).

    python -m swmhd_b200.build [--force]

substage_kernel.cu is compiled twice: STRICT (-fmad=false, bit-identical to the oracle's
arithmetic) and FAST (FMA contraction); both instantiations live in the one shared library.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = CSRC / "_obj"
LIB = HERE / "libswmhd_cuda.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]

UNITS = [
    # (source, object, extra flags)
    ("substage_kernel.cu", "substage_strict.o", ["-DSWMHD_STRICT=1", "-fmad=false", "-prec-div=true", "-prec-sqrt=true"]),
    ("substage_kernel.cu", "substage_fast.o", ["-DSWMHD_STRICT=0"]),
    ("substage_rb.cu", "substage_rb.o", []),
    ("aux_kernels.cu", "aux_kernels.o", []),
    ("swmhd_api.cu", "swmhd_api.o", []),
    ("nccl_dyn.cpp", "nccl_dyn.o", []),
]


def _deps():
    return list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cpp")) + list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "swmhd.h"]


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in _deps())


def _compile(unit):
    src, obj, extra = unit
    cmd = [NVCC, *ARCH, *COMMON, *extra, "-c", str(CSRC / src), "-o", str(OBJ / obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    (OBJ / (obj + ".log")).write_text(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src} -> {obj}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    OBJ.mkdir(exist_ok=True)
    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(_compile, UNITS))
    cmd = [NVCC, *ARCH, "-shared", "-o", str(LIB), *[str(OBJ / o) for o in objs], "-cudart", "static", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        for o in objs:
            print((OBJ / (o + ".log")).read_text())
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
