"""Timeline of one tile pair of the pair decode kernel (build: python scripts/build_variant.py trace decode_fwd_pair.cu -DSW_PAIR_TRACE;
run with SOCIALWAYS_B200_LIB=socialways_b200/build/ab/libsw_trace.so): clock64 of thread 0 of CTA 0 at the points SW_TR marks, third tile pair."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv = [sys.argv[0], "pair"]
exec(open(os.path.join(ROOT, "scripts", "ncu_tcx.py")).read())
from socialways_b200 import _lib
L = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_longlong * 64)()
assert L.sw_pair_trace_read(buf) == 0
v = list(buf)
names = {0: "tile pair starts", 1: "c0 / S cp.async / x_last issued", 18: "prologue staging of both slots done, prefetches issued", 48: "last step done"}
for sl in (0, 1):
    b = 2 + 8 * sl
    names.update({b: f"slot {sl}: h0 loads issued / staging starts", b + 1: f"slot {sl}: cp.async landed", b + 2: f"slot {sl}: noise block landed",
                  b + 3: f"slot {sl}: barrier 1", b + 4: f"slot {sl}: [S;z] -> TMEM", b + 5: f"slot {sl}: barrier 2", b + 6: f"slot {sl}: h0 split, arrive (hoist may start)",
                  19 + 2 * sl: f"slot {sl}: hoist MMAs done", 20 + 2 * sl: f"slot {sl}: c1 -> scratch, arrive"})
for sl in (0, 1):
    names.update({49 + 4 * sl: f"  slot {sl}: row index arithmetic", 50 + 4 * sl: f"  slot {sl}: c0 loads issued", 51 + 4 * sl: f"  slot {sl}: S cp.async issued"})
for base, tag in ((24, "t=1"), (32, "t=n-2"), (40, "t=n-1")):
    for i, nm in enumerate(("step top", "L1(0)", "cell(1)+pf", "L2(0)", "L1(1)", "cell(0)+pf", "L2(1)")):
        names[base + i] = f"{tag}: {nm}"
t0 = v[0]
prev = t0
for i in sorted(names, key=lambda j: v[j]):
    if v[i]:
        print(f"{i:3d} {v[i] - t0:8d} (+{v[i] - prev:6d})  {names[i]}")
        prev = v[i]
