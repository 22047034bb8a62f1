"""Aggregate the warp-stall samples of an ncu report (--set full --import-source on, built with -lineinfo) by CUDA source
line:  python scripts/ncu_source_lines.py report.ncu-rep [top_n]  -- the per-line share of samples and the dominant stall
reasons, files interleaved as ncu prints them."""
import csv
import subprocess
import sys


def main(path, top=45):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source=cuda,sass"], stdout=subprocess.PIPE,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, H, lines = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Line No":
            H = r
        elif H and r[0] != "" and len(r) == len(H):
            lines.append((cur_file, r))
    si = H.index("# Samples")
    stall_cols = [i for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[si]) for _, r in lines)
    print(f"# {path}: {tot} samples")
    lines.sort(key=lambda fr: -int(fr[1][si]))
    for f, r in lines[:top]:
        st = {H[i][6:]: int(r[i]) for i in stall_cols if int(r[i]) > 0}
        tops = " ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:4])
        print(f"{f}:{r[0]:>4} {100 * int(r[si]) / tot:5.1f}%  {r[1].strip()[:80]:80s} {tops}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
