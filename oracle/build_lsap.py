"""gcc recipe for oracle/lsap_ref.c -> oracle/liblsap_ref.so (test infrastructure; see the header of the .c file)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC, LIB = os.path.join(HERE, "lsap_ref.c"), os.path.join(HERE, "liblsap_ref.so")


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-ffp-contract=off", SRC, "-o", LIB, "-lm"], check=True)
    return LIB


def solve(cost):
    """cost [n, n] float64 numpy -> col4row int32 [n] (raises ValueError if infeasible)."""
    import ctypes
    import numpy as np
    lib = ctypes.CDLL(build())
    lib.lsap_ref_solve.restype = ctypes.c_int
    lib.lsap_ref_solve.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    n = cost.shape[0]
    col = np.empty(n, dtype=np.int32)
    if lib.lsap_ref_solve(cost.ctypes.data, n, col.ctypes.data):
        raise ValueError("cost matrix is infeasible")
    return col


if __name__ == "__main__":
    print(build(force=True))
