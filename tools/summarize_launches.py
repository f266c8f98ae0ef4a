"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share."""
import csv, sys, re
rows = list(csv.reader(open(sys.argv[1], errors='ignore')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
H = rows[hdr]; ki, vi, ui = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
agg = {}
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    n = r[ki]
    m = re.search(r'(?:anonymous namespace)::(\w+)(<[^>]*>)?', n)
    name = (m.group(1) + (m.group(2) or '')) if m else re.sub(r'^void ', '', n).split('(')[0][:70]
    v = float(r[vi].replace(',', '')); u = r[ui]
    us = v / 1000 if u.startswith('ns') else v
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
print(f'total {tot/1000:.3f} ms over {sum(a[0] for a in agg.values())} launches')
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f'{t/1000:9.3f} ms {100*t/tot:5.1f}% {n:5d}x  {k}')
