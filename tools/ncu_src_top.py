"""Top stall instructions of a kernel from `ncu -i rep --page source --csv` output: python tools/ncu_src_top.py file.csv [n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ia, isrc, iall, inot, iex = hdr.index('Address'), hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Warp Stall Sampling (Not-issued Samples)'), hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'not_issued' not in h.lower()]
data = [r for r in rows[2:] if len(r) > iall and r[iall].isdigit()]
tot = sum(int(r[iall]) for r in data)
print('total samples', tot, 'instructions', len(data))
ops = collections.Counter()
for r in data:
    ops[r[isrc].split()[0] if not r[isrc].strip().startswith('@') else r[isrc].split()[1]] += int(r[iex])
print('executed by opcode:', ', '.join(f'{k}:{v}' for k, v in ops.most_common(25)))
for idx, r in sorted(enumerate(data), key=lambda z: -int(z[1][iall]))[:n]:
    top = sorted(((int(r[i]), hdr[i]) for i in stall_cols if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:3]
    print(f'{idx:5d} {int(r[iall]):6d} ({100*int(r[iall])/tot:4.1f}%) ex={r[iex]:>8s}  {r[isrc].strip()[:70]:70s} ' + ' '.join(f'{h[6:]}={v}' for v, h in top))
