import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); ii = h.index('ID')
seq = [(int(r[ii]), r[ki], float(r[vi].replace(',', ''))) for r in rows[hi + 2:] if len(r) > vi]
if len(sys.argv) > 2 and sys.argv[2] == 'seq':
    lo, hi2 = int(sys.argv[3]), int(sys.argv[4])
    for s in seq[lo:hi2]: print(s[0], s[1][:70], s[2] / 1e3)
else:
    agg = collections.OrderedDict()
    for _, k, v in seq:
        k = k[:70]
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{100*a[1]/tot:5.1f}% {a[0]:4d} {a[1]/a[0]/1e3:9.1f} us  {k}")
    print(len(seq), "launches", tot / 1e6, "ms")
