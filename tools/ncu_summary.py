"""Summarise an .ncu-rep: per-kernel headline metrics, opcode histogram, top stall lines.  usage: ncu_summary.py rep [kernel-substr]"""
import csv, collections, re, subprocess, sys
rep = sys.argv[1]; sel = sys.argv[2] if len(sys.argv) > 2 else ''
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'inst_executed',
        'sass__thread_inst_executed_true_per_opcode', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']
kn = hdr.index('Kernel Name')
for r in rows[2:]:
    if sel not in r[kn]: continue
    print('==', r[kn][:90])
    for w in want:
        if w in hdr: print('   %-95s %s %s' % (w, r[hdr.index(w)], rows[1][hdr.index(w)]))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
kern = None; data = {}; h2 = None
for r in rows:
    if r and r[0] == 'Kernel Name': kern = r[1]; data[kern] = []; continue
    if r and r[0] == 'Address': h2 = r; continue
    if kern and len(r) > 10: data[kern].append(r)
for kname, rr in data.items():
    if sel not in kname: continue
    ie, isamp, isrc = h2.index('Instructions Executed'), h2.index('# Samples'), h2.index('Source')
    ops = collections.Counter(); samp = collections.Counter(); tot = 0
    for r in rr:
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[isrc]); op = m.group(2).split('.')[0] if m else '?'
        n = int(r[ie]); ops[op] += n; samp[op] += int(r[isamp]); tot += n
    print('== opcodes', kname[:80], 'total', tot)
    for op, n in ops.most_common(18): print('   %-10s %12d %5.1f%%  samples %d' % (op, n, 100 * n / tot, samp[op]))
