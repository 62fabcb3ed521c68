"""Hot spots of an ncu report's source page (run here, no GPU):
python profiles/ncu_source_hot.py x.ncu-rep [top]  — per-source-line instruction and stall-sample shares."""
import csv
import re
import subprocess
import sys


def main(path, top=25, view="sass"):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass,cuda" if view == "cuda" else "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] in ("Address", "#")]
    for h in hi:
        hdr = rows[h]
        if "Instructions Executed" not in hdr:
            continue
        ia, isamp, it = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
        data = [r for r in rows[h + 1:] if len(r) > ia and r[ia].isdigit()]
        tot = sum(int(r[ia]) for r in data)
        stot = sum(int(r[isamp]) for r in data)
        print(f"kernel: {rows[h-1][1][:100] if h else ''}  warp-instr {tot/1e6:.1f}M  samples {stot}")
        ranked = sorted(data, key=lambda r: -int(r[isamp]))[:top]
        for r in ranked:
            print(f"  {r[0][-5:]}  instr {int(r[ia])/tot*100:5.2f}%  samples {int(r[isamp])/max(stot,1)*100:5.2f}%  "
                  f"thr/warp {int(r[it])/max(int(r[ia]),1):4.1f}  {r[1].strip()[:80]}")
        break


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
