"""Builds libfqsk.so (hand-written sm_100a kernels + the C-ABI of include/fqsk.h) in-tree with nvcc.

    python -m fqsqueezer_b200.build [--force]

nvcc cross-compiles without a GPU; the .so sits next to this file so that it travels with the repo snapshot."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfqsk.so")
SOURCES = [os.path.join(CSRC, "fqsk.cu")]
HEADERS = [os.path.join(CSRC, "fqsk_dev.cuh"), os.path.join(CSRC, "fqsk_kernels.cuh"), os.path.join(CSRC, "fqsk_pipeline.cuh"), os.path.join(CSRC, "fqsk_pe.cuh"),
           os.path.join(CSRC, "fqsk_sort.cuh"), os.path.join(CSRC, "fqsk_front.cuh"), os.path.join(CSRC, "fqsk_mtjump.h"),
           os.path.join(os.path.dirname(HERE), "include", "fqsk.h"), os.path.join(os.path.dirname(HERE), "include", "fqsk_ctx.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--expt-relaxed-constexpr", "-diag-suppress", "177,550"]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS + [os.path.abspath(__file__)])


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-o", LIB, *SOURCES, "-lcudart"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
