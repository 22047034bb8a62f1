"""Work / wait split of a kernel from an ncu report (--set full --import-source on): walks the SASS in address order and prints,
for every mbarrier wait loop (SYNCS.PHASECHK.TRYWAIT ... BRA) and every BAR.SYNC, the warp-stall samples spent in the wait
and the samples of the code since the previous wait -- i.e. how long each phase of a warp-specialised kernel computes and how
long it waits for the MMAs it depends on.   python scripts/ncu_wait_sites.py report.ncu-rep [start_marker]"""
import csv
import subprocess
import sys


def main(path, marker=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source=sass"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    for i, r in enumerate(rows):
        if r and r[0] == "Address":
            H, start = r, i
            break
    si, src = H.index("# Samples"), H.index("Source")
    ins = [(r[0], r[src], int(r[si])) for r in rows[start + 1:] if len(r) == len(H)]
    i = 0
    if marker:
        i = [k for k, (_, s, _) in enumerate(ins) if marker in s][0]
    total = sum(n for _, _, n in ins)
    print(f"# {path}: {total} samples")
    work = 0
    while i < len(ins):
        a, s, n = ins[i]
        if "TRYWAIT" in s:
            w, j = 0, i
            while j < len(ins) and "BRA" not in ins[j][1] and j - i < 12:
                w += ins[j][2]
                j += 1
            w += ins[j][2] if j < len(ins) else 0
            print(f"work {work:6d} | wait {w:6d} at {a[-5:]} {s.strip()[:64]}")
            work, i = 0, j + 1
            continue
        if "BAR.SYNC" in s:
            print(f"work {work:6d} | BAR  {n:6d} at {a[-5:]}")
            work = 0
        work += n
        i += 1
    print("tail work", work)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
