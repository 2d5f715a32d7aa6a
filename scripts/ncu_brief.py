"""Brief digest of one .ncu-rep: time, occupancy, pipe utilisation, stalls, SASS opcode mix.  usage: ncu_brief.py rep [rep..]"""
import collections, csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__maximum_warps_per_active_cycle_pct',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct']
for rep in sys.argv[1:]:
    o = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(o.splitlines()))
    h, u = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(h, r)); uu = dict(zip(h, u))
        print('==', rep, d.get('Kernel Name', '?')[:120])
        for k in KEYS:
            if k in d: print('  %-70s %s %s' % (k, d[k], uu[k]))
        st = {k.split('issue_stalled_')[1].split('_per_issue')[0]: float(v) for k, v in d.items()
              if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio') and v}
        print('  stalls (warps per issue):', ', '.join('%s %.2f' % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
    o = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(o.splitlines()))
    hdr = rows[1]; ix = {hh: i for i, hh in enumerate(hdr)}
    cnt = collections.Counter(); smp = collections.Counter()
    for r in rows[2:]:
        if len(r) < len(hdr): continue
        t = r[ix['Source']].split()
        if not t: continue
        op = t[1] if t[0].startswith('@') and len(t) > 1 else t[0]
        op = op.split('.')[0] + ('.' + op.split('.')[1] if op.startswith('MUFU') or op.startswith('F2F') else '')
        cnt[op] += int(r[ix['Instructions Executed']] or 0); smp[op] += int(r[ix['# Samples']] or 0)
    tot = sum(cnt.values()) or 1; ts = sum(smp.values()) or 1
    print('  SASS mix (%d warp-inst): ' % tot + ', '.join('%s %.1f%% (smp %.1f%%)' % (op, 100.0 * n / tot, 100.0 * smp[op] / ts) for op, n in cnt.most_common(18)))
