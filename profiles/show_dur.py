import csv, sys
for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    for i, r in enumerate(rows):
        if 'Kernel Name' in r:
            h = r; s = i + 1; break
    kn, mn, mv, idc = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('ID')
    print(f)
    cur = None
    for r in rows[s:]:
        if len(r) > mv and int(r[idc]) < 2:
            if r[idc] != cur:
                print('  ', r[kn][:62]); cur = r[idc]
            print('      ', r[mn], r[mv])
