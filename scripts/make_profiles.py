"""Summarise gpurun_out/*.ncu-rep + launches.csv into tracked files under profiles/ (round tag as argv[1])."""
import collections, csv, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(ROOT, 'profiles'); os.makedirs(out_dir, exist_ok=True)
go = os.path.join(ROOT, 'gpurun_out')

def raw(rep):
    o = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(o.splitlines()))
    return rows[0], rows[1], rows[2:]

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']
STALLS = ['long_scoreboard', 'short_scoreboard', 'wait', 'barrier', 'math_pipe_throttle', 'mio_throttle', 'not_selected',
          'branch_resolving', 'dispatch_stall', 'no_instruction', 'lg_throttle']

def opcodes(rep, top=14):
    o = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(o.splitlines()))
    hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
    cnt = collections.Counter(); smp = collections.Counter()
    for r in rows[2:]:
        if len(r) < len(hdr): continue
        t = r[ix['Source']].split()
        if not t: continue
        op = t[1] if t[0].startswith('@') and len(t) > 1 else t[0]
        op = op.split('.')[0]
        cnt[op] += int(r[ix['Instructions Executed']] or 0); smp[op] += int(r[ix['# Samples']] or 0)
    tot = sum(cnt.values()) or 1; ts = sum(smp.values()) or 1
    return tot, [(op, n, 100.0 * n / tot, 100.0 * smp[op] / ts) for op, n in cnt.most_common(top)]

lines = ['# ncu summaries, round %s' % tag, '',
         'One `ncu --set full --clock-control none --import-source on` capture per hot kernel (one launch each, taken inside',
         '`python bench.py --steps 3 --warmup 3`), read with `ncu -i ... --page raw/source --csv` (scripts/make_profiles.py).',
         'Times under ncu are cold-cache and serialised: they document the SHARE and the pipe mix, not the bench value.', '']
for rep in sorted(f for f in os.listdir(go) if f.endswith('.ncu-rep')):
    hdr, units, rows = raw(os.path.join(go, rep))
    for r in rows:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        lines += ['## %s' % rep, '', '`%s`' % d.get('Kernel Name', '?'), '', '| metric | value | unit |', '|---|---|---|']
        for k in KEYS:
            if k in d: lines.append('| %s | %s | %s |' % (k, d[k], u[k]))
        for st in STALLS:
            k = 'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % st
            if k in d: lines.append('| stall %s (warps per issue) | %s | |' % (st, d[k]))
        tot, ops = opcodes(os.path.join(go, rep))
        lines += ['', 'SASS mix (%d warp-instructions): ' % tot + ', '.join('%s %.1f%% (stall samples %.1f%%)' % (op, p, s) for op, n, p, s in ops), '']
open(os.path.join(out_dir, '%s_ncu_summary.md' % tag), 'w').write('\n'.join(lines) + '\n')

# per-kernel DRAM traffic of the captured launch (bench.py reports it as roofline.traffic)
import json, re
traffic = {}
def _bytes(v, u):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
for rep in sorted(f for f in os.listdir(go) if f.endswith('.ncu-rep')):
    hdr, units, rows = raw(os.path.join(go, rep))
    for r in rows:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        name = re.sub(r'^void |unnamed>::|<unnamed>::|gabo::', '', d.get('Kernel Name', '?')).split('(')[0]
        if name in traffic:      # first capture of a kernel wins (prof_spd_gram = the headline launch sorts before _n8192)
            continue
        traffic[name] = {'dram_bytes_read': _bytes(d['dram__bytes_read.sum'], u['dram__bytes_read.sum']),
                         'dram_bytes_write': _bytes(d['dram__bytes_write.sum'], u['dram__bytes_write.sum']),
                         'gpu_time_us_under_ncu': float(d['gpu__time_duration.sum'].replace(',', '')) * {'us': 1, 'ms': 1e3, 'ns': 1e-3}[u['gpu__time_duration.sum']],
                         'report': rep}
json.dump(traffic, open(os.path.join(out_dir, '%s_traffic.json' % tag), 'w'), indent=1)

# launch list: aggregate per kernel + keep the raw csv (small)
lc = os.path.join(go, 'launches.csv')
if os.path.exists(lc):
    rows = [r for r in csv.reader(open(lc)) if len(r) > 5]
    hdr = None; agg = collections.OrderedDict()
    for r in rows:
        if r[0] == 'ID': hdr = r; continue
        if hdr is None: continue
        d = dict(zip(hdr, r))
        try: v = float(d['Metric Value'].replace(',', ''))
        except ValueError: continue
        v = v / 1e3 if d['Metric Unit'] == 'ns' else (v * 1e3 if d['Metric Unit'] == 'ms' else v)
        agg.setdefault((d['Kernel Name'], d['Grid Size'], d['Block Size']), []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(out_dir, '%s_launches_summary.md' % tag), 'w') as f:
        f.write('# launch list, round %s\n\n`ncu --metrics gpu__time_duration.sum --clock-control none -c 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline`\n'
                '(cold-cache, serialised; shares only).  `FillFunctor<unsigned char>` is the 256 MB L2 flush between timed steps.\n\n'
                '| kernel | grid | block | launches | mean us | total us | share |\n|---|---|---|---|---|---|---|\n' % tag)
        for (k, g, b), v in agg.items():
            f.write('| `%s` | %s | %s | %d | %.1f | %.0f | %.1f%% |\n' % (k[:110], g, b, len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot))
    import shutil; shutil.copy(lc, os.path.join(out_dir, '%s_launches.csv' % tag))
print('wrote', os.listdir(out_dir))
