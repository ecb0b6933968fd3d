"""Scratch: top stall sites of an `ncu --page source --csv` export.  usage: top_stalls.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
ia, isrc, isamp = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples')
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for k, r in enumerate(rows[2:]):
    try: n = int(r[isamp])
    except (ValueError, IndexError): continue
    data.append((n, k, r))
tot = sum(d[0] for d in data)
print('total samples', tot)
for n, k, r in sorted(data, key=lambda d: -d[0])[:N]:
    top = sorted(((int(r[i] or 0), h) for i, h in stall), reverse=True)[:2]
    print(f'{n:7d} {100*n/tot:5.1f}%  #{k:5d} {r[isrc].strip()[:70]:70s} {top}')
