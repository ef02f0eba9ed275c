"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list per kernel.
usage: launch_list.py file.csv"""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors='replace')) if len(r) > 8]
h = rows[0]
ik, im, iv, iu = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('Metric Unit')
t = collections.defaultdict(list); by = collections.defaultdict(float)
for r in rows[1:]:
    v = float(r[iv].replace(',', ''))
    if r[im] == 'gpu__time_duration.sum':
        t[r[ik]].append(v * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(r[iu], 1e-6))
    elif r[im].startswith('dram__bytes'):
        by[r[ik]] += v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(r[iu], 1)
tot = sum(sum(v) for v in t.values())
print('total %.2f ms over %d launches' % (tot, sum(len(v) for v in t.values())))
for k, v in sorted(t.items(), key=lambda kv: -sum(kv[1])):
    extra = '  dram %.2f GB/launch' % (by[k] / len(v) / 1e9) if k in by else ''
    print('%-70s n=%4d  avg %9.3f ms  sum %9.2f ms  %5.1f%%%s' % (k[:70], len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot, extra))
