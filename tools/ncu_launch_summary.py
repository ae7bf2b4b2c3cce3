"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
d = collections.defaultdict(list)
for r in rows[1:]:
    try:
        d[r[ki][:70]].append(float(r[vi].replace(',', '')))
    except ValueError:
        pass
for k, v in d.items():
    v.sort()
    print(f"{k:70s} n={len(v):3d} median={v[len(v)//2]/1000:8.2f} us  min={v[0]/1000:8.2f}  max={v[-1]/1000:8.2f}")
