"""Build libmcl_b200.so (sm_100a only) in-tree with nvcc.

    python slam-robot_simu_b200/build.py [--force] [--verbose]

The library is the C ABI of include/mcl.h: hand-written CUDA kernels plus their host plumbing, no
torch or Python types.  nvcc cross-compiles without a GPU, so this runs in the authoring container
and the built .so travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmcl_b200.so")
SOURCES = [os.path.join(CSRC, "mcl_api.cu")]
DEPS = SOURCES + [os.path.join(CSRC, "mcl_kernels.cuh"), os.path.join(CSRC, "mcl_device.cuh"),
                  os.path.join(ROOT, "include", "mcl.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC,-ffp-contract=off",
    # fp64 rounding of the comb (i * (1/N) + ofs) must not be contracted on the device either;
    # the kernels use explicit __dmul_rn/__dadd_rn there, everything else may use FMA.
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmcl_b200.so cannot be built (there is no CPU fallback)")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    extra = ["-DMCL_TIMELINE"] if os.environ.get("MCL_TIMELINE") == "1" else []  # phase stamps for tools/timeline.py
    extra += os.environ.get("MCL_NVCC_EXTRA", "").split()  # experiments: e.g. -DMCL_SCAN_CTAS_PER_SM=5
    out = os.environ.get("MCL_LIB_OUT", LIB)               # ... built beside the product library, never over it
    cmd = [find_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + SOURCES
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libmcl_b200.so:\n" + proc.stderr[-4000:])
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
