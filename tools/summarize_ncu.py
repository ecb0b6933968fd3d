"""Condense an `ncu --page raw --csv` export into a per-launch table (markdown) for profiles/."""
import csv, re, sys
src, title = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src)))
hdr = rows[0]
def col(name):
    m = [i for i, h in enumerate(hdr) if h == name] or [i for i, h in enumerate(hdr) if h.endswith('.' + name)]
    return m[0] if m else None
cols = [('kernel', 'Kernel Name'), ('grid', 'launch__grid_size'), ('block', 'launch__block_size'), ('regs', 'launch__registers_per_thread'),
        ('dur_us', 'gpu__time_duration.sum'), ('dram_rd_MB', 'dram__bytes_read.sum'), ('dram_wr_MB', 'dram__bytes_write.sum'),
        ('dram_%', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
        ('tensor_%', 'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed'),
        ('tensor_mem_%', 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'),
        ('smem_tc_wavefronts_%', 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'),
        ('sm_%', 'sm__throughput.avg.pct_of_peak_sustained_elapsed'), ('issue_%', 'sm__inst_issued.avg.pct_of_peak_sustained_active'),
        ('l2_%', 'lts__throughput.avg.pct_of_peak_sustained_elapsed')]
idx = [(n, col(c)) for n, c in cols]
units = rows[1]
print(f'### {title}\n')
print('| ' + ' | '.join(n + (f' [{units[i]}]' if i is not None and units[i] and n not in ("kernel",) else '') for n, i in idx) + ' |')
print('|' + '---|' * len(idx))
for r in rows[2:]:
    vals = []
    for n, i in idx:
        v = r[i] if i is not None else ''
        if n == 'kernel':
            v = re.sub(r'\(.*', '', v).replace('<unnamed>::', '').replace('void ', '')
        else:
            try: v = f'{float(v):.3g}' if '.' in v else v
            except ValueError: pass
        vals.append(v)
    print('| ' + ' | '.join(vals) + ' |')
