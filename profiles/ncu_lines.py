"""Per-CUDA-source-line instruction and stall-sample shares of an ncu report (needs -lineinfo and
--import-source on):  python profiles/ncu_lines.py x.ncu-rep [min_pct]"""
import collections
import csv
import subprocess
import sys


def main(path, min_pct=0.5):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname = ""
    acc = collections.OrderedDict()
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Name":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            ia, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or len(r) <= ia or not r[ia].isdigit():
            continue
        key = (fname, r[0], r[1].strip())
        a = acc.setdefault(key, [0, 0])
        a[0] += int(r[ia])
        a[1] += int(r[isamp]) if r[isamp].isdigit() else 0
    tot = sum(a[0] for a in acc.values()) or 1
    stot = sum(a[1] for a in acc.values()) or 1
    print(f"total warp-instr {tot/1e6:.1f}M  samples {stot}")
    for (f, ln, src), (i, s) in acc.items():
        if i / tot * 100 >= min_pct or s / stot * 100 >= min_pct:
            print(f"{f}:{ln:>4}  instr {i/tot*100:5.1f}%  samples {s/stot*100:5.1f}%  {src[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.5)
