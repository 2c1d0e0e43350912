"""In-tree build of libcfgpu.so (nvcc, sm_100a only).  Used by __graft_entry__.build() and runnable
directly: `python -m mcp_context_forge_b200.build`.  nvcc cross-compiles without a GPU."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libcfgpu.so")
SOURCES = ["cfgpu.cu", "cf_host.cpp", "re_backend.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libcfgpu.so cannot be built")


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    so_m = os.path.getmtime(SO)
    for root, _, files in os.walk(CSRC):
        for f in files:
            if os.path.getmtime(os.path.join(root, f)) > so_m:
                return True
    inc = os.path.join(os.path.dirname(HERE), "include", "cfgpu.h")
    return os.path.exists(inc) and os.path.getmtime(inc) > so_m


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return SO
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO] + srcs
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        sys.stderr.write(proc.stderr)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
