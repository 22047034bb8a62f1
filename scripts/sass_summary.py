"""Per-kernel SASS instruction summary of the built library (what proves a Blackwell-native kernel, B200_PROFILING.md):
UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UBLKCP = TMA (tensor-map / bulk copy), SYNCS = mbarrier,
FFMA / MUFU = CUDA-core arithmetic.   python scripts/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "socialways_b200", "libsocialways_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "FFMA", "MUFU",
        "LDGSTS", "LDG", "STG", "LDS", "STS", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            cur = kernels.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            cur["total"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + ".") or (k.startswith("UTC") and k != "UTCBAR" and op.startswith("UTC") and op.endswith("MMA") and k == "UTCMMA"):
                    cur[k] += 1
                    break
            if op.startswith("UTC") and "MMA" in op:
                cur["tcgen05.mma (UTC*MMA)"] += 1
    print("# cuobjdump -sass socialways_b200/libsocialways_b200.so : instruction counts per kernel (static)")
    cols = ["total", "tcgen05.mma (UTC*MMA)", "LDTM", "STTM", "UTMALDG", "UBLKCP", "SYNCS", "FFMA", "MUFU", "LDG", "STG", "LDS", "STS", "BAR"]
    print(f"{'kernel':58s} " + " ".join(f"{c.split(' ')[0][:9]:>9s}" for c in cols))
    for name, c in kernels.items():
        print(f"{name[:58]:58s} " + " ".join(f"{c[k]:9d}" for k in cols))


if __name__ == "__main__":
    main()
