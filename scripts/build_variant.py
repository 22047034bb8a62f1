"""A/B builds of one kernel file: python scripts/build_variant.py NAME FILE.cu -DX=1 ...  compiles socialways_b200/csrc/FILE.cu with the
extra flags and links it with the objects of the regular build into socialways_b200/build/ab/libsw_NAME.so (select it with
SOCIALWAYS_B200_LIB=<path>; the directory travels to the GPU box with the snapshot, it is git-ignored)."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from socialways_b200 import build as B

name, src, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
B.build()
ab = os.path.join(B.HERE, "build", "ab")
os.makedirs(ab, exist_ok=True)
obj = os.path.join(ab, f"{name}_{src[:-3]}.o")
subprocess.check_call(["nvcc"] + [f for f in B.NVCC_FLAGS if f not in ("-Xptxas", "-v")] + flags +
                      ["-I", B.CSRC, "-I", B.INCLUDE, "-c", os.path.join(B.CSRC, src), "-o", obj])
objs = [o for o in glob.glob(os.path.join(B.HERE, "build", "*.o")) if os.path.basename(o) != src[:-3] + ".o"] + [obj]
lib = os.path.join(ab, f"libsw_{name}.so")
subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs)
print(lib)
