"""Builds solaris_b200/libsolaris_b200.so (hand-written CUDA for sm_100a + the C-ABI) in-tree with nvcc.

    python -m solaris_b200.build [--force]

Three translation units with different floating-point contracts (see csrc/common.cuh):
gravity.cu keeps FMA contraction (explicit fma anyway), elementwise.cu is compiled with -fmad=false
so stage combinations / error norms reproduce the reference's non-fused x86-64 arithmetic bit for bit.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libsolaris_b200.so")
OBJ = os.path.join(HERE, "csrc", "_obj")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall"] + ARCH
UNITS = {
    "gravity.cu": [],
    "elementwise.cu": ["-fmad=false"],
    "api.cu": ["-fmad=false"],
}


def _nccl_include() -> list[str]:
    for cand in ("/usr/include/nccl.h",):
        if os.path.exists(cand):
            return []
    try:
        import nvidia.nccl  # type: ignore
        inc = os.path.join(os.path.dirname(nvidia.nccl.__file__), "include")
        if os.path.exists(os.path.join(inc, "nccl.h")):
            return ["-I", inc]
    except Exception:
        pass
    return []


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(HERE, "..", "include", "solaris_b200.h"))
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    procs = []
    objs = []
    for src, extra in UNITS.items():
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc] + COMMON + extra + _nccl_include() + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    link = [nvcc, "-shared"] + ARCH + ["-o", OUT] + objs + ["-lcudart", "-ldl"]
    subprocess.check_call(link)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
