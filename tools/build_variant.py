"""A/B helper: builds solaris_b200/libsolaris_b200_<tag>.so with extra -D flags for ONE translation unit.

    python tools/build_variant.py <tag> <unit.cu> -DX=1 [-DY=2 ...]
Use with SOLARIS_B200_LIB=<path> (tools/probe_variant.py, tools/bench_configs.py)."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from solaris_b200 import build as B

tag, unit, defs = sys.argv[1], sys.argv[2], sys.argv[3:]
B.build()
obj = os.path.join(B.OBJ, unit.replace(".cu", f"_{tag}.o"))
subprocess.check_call(["nvcc"] + B.COMMON + B.UNITS[unit] + B._nccl_include() + defs + ["-c", os.path.join(B.CSRC, unit), "-o", obj])
objs = [os.path.join(B.OBJ, u.replace(".cu", ".o")) for u in B.UNITS if u != unit] + [obj]
out = os.path.join(os.path.dirname(B.OUT), f"libsolaris_b200_{tag}.so")
subprocess.check_call(["nvcc", "-shared"] + B.ARCH + ["-o", out] + objs + ["-lcudart", "-ldl"])
print(out)
