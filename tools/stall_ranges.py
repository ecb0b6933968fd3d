"""Scratch: samples per SASS index range (role) with stall-reason totals.  usage: stall_ranges.py src.csv bin"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 100
isrc, isamp = hdr.index('Source'), hdr.index('# Samples')
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
bins = collections.defaultdict(lambda: [0, collections.Counter(), collections.Counter()])
for k, r in enumerate(rows[2:]):
    try: n = int(r[isamp])
    except (ValueError, IndexError): continue
    b = bins[k // B]; b[0] += n
    for i, h in stall: b[1][h] += int(r[i] or 0)
    op = r[isrc].strip().split()[0] if r[isrc].strip() else ''
    if op.startswith('@'): op = r[isrc].strip().split()[1]
    b[2][op.split('.')[0]] += 1
tot = sum(b[0] for b in bins.values())
for k in sorted(bins):
    n, c, ops = bins[k]
    print(f'{k*B:5d}-{k*B+B-1:5d} {n:7d} {100*n/tot:5.1f}%  {[(h[6:], v) for h, v in c.most_common(3)]}  ops {[o for o,_ in ops.most_common(5)]}')
