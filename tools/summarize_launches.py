"""Condense an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list of one bench step into a
per-kernel table (markdown on stdout) and profiles/r1_conv_traffic.json (average DRAM bytes per convolution launch, read by bench.py)."""
import collections, csv, json, re, sys
src = sys.argv[1]
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
h = rows[0]; ik = h.index('Kernel Name'); im = h.index('Metric Name'); iv = h.index('Metric Value'); iu = h.index('Metric Unit'); iid = h.index('ID')
L = collections.OrderedDict()
for r in rows[1:]:
    d = L.setdefault(r[iid], {'name': r[ik]})
    try:
        v = float(r[iv].replace(',', ''))
    except ValueError:
        continue
    u = r[iu]
    if r[im] == 'gpu__time_duration.sum':
        d['us'] = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    elif r[im].startswith('dram__bytes_read'):
        d['rd'] = v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
    elif r[im].startswith('dram__bytes_write'):
        d['wr'] = v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]


def short(n):
    n = re.sub(r'^void ', '', n).replace('<unnamed>::', '').replace('(int)', '')
    n = re.sub(r'\(.*$', '', n)
    return re.sub(r'<.*', '', n) if n.startswith('at::') else n


agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
tot = 0.0
for d in L.values():
    if 'us' not in d:
        continue
    a = agg[short(d['name'])]; a[0] += 1; a[1] += d['us']; a[2] += d.get('rd', 0); a[3] += d.get('wr', 0); tot += d['us']
print(f'{len(L)} launches, {tot / 1e3:.2f} ms of kernel time in one 64-frame step (cold-cache, serialised under ncu: compare shares)\n')
print('| kernel | launches | ms / step | share | DRAM read GB | DRAM write GB | DRAM GB/s |')
print('|---|---|---|---|---|---|---|')
conv_b = conv_n = 0
for k, (n, t, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'| `{k}` | {n} | {t / 1e3:.2f} | {100 * t / tot:.1f}% | {rd / 1e9:.2f} | {wr / 1e9:.2f} | {(rd + wr) / t / 1e3:.0f} |')
    if k.startswith('conv_'):
        conv_b += rd + wr; conv_n += n
if len(sys.argv) > 2:
    json.dump({'dram_bytes_per_conv_launch': conv_b / conv_n, 'conv_launches': conv_n,
               'source': f'{src} (ncu dram__bytes_read.sum + dram__bytes_write.sum, one 64-frame step)'}, open(sys.argv[2], 'w'), indent=1)
