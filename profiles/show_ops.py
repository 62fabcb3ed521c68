import json, sys
for l in sys.stdin:
    l = l.strip()
    if not l:
        continue
    d = json.loads(l)
    if "error" in d:
        print("ERR", d)
        continue
    print(f"{d['ms']:10.3f} ms  {d['achieved_GBps']:8.1f} GB/s  {d['frac_of_peak']:.3f}  {d['op']}")
