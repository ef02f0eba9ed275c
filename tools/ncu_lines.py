"""Per-CUDA-source-line share of warp-stall samples and executed instructions.  usage: ncu_lines.py rep kernel-substr [min-pct]"""
import csv, collections, subprocess, sys
rep, sel = sys.argv[1], sys.argv[2]
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
def num(s):
    try: return int(s)
    except ValueError: return 0
cur_file = cur_fn = hdr = None
agg = collections.defaultdict(lambda: [0, 0, ''])
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': cur_fn = r[1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if r[0] and r[0].isdigit():
        k = (cur_fn, cur_file, int(r[0]))
        agg[k][0] += num(r[hdr.index('# Samples')]); agg[k][1] += num(r[hdr.index('Instructions Executed')]); agg[k][2] = r[1]
items = sorted(((k, v) for k, v in agg.items() if sel in k[0]), key=lambda kv: (kv[0][1], kv[0][2]))
ts = sum(v[0] for k, v in items) or 1; ti = sum(v[1] for k, v in items) or 1
print(sel, 'samples', ts, 'warp instructions', ti)
for k, v in items:
    if 100 * v[0] / ts >= minp or 100 * v[1] / ti >= minp:
        print('%-16s %4d samp %5.1f%% inst %5.1f%%  %s' % (k[1], k[2], 100 * v[0] / ts, 100 * v[1] / ti, v[2][:110]))
