"""Aggregate an `ncu --page source --csv` dump by SASS opcode: executed warp-instructions and stall samples."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); samples = collections.Counter()
tot = 0; tots = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix['Source']].strip()
    toks = src.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
    op = op.split('.')[0] if '--full' not in sys.argv else op
    n = int(r[ix['Instructions Executed']] or 0); s = int(r[ix['# Samples']] or 0)
    ops[op] += n; samples[op] += s; tot += n; tots += s
print('total warp-instr', tot, 'samples', tots)
for op, n in ops.most_common(30):
    print('%-12s %12d %5.1f%%   samples %5.1f%%' % (op, n, 100.0 * n / tot, 100.0 * samples[op] / max(tots, 1)))
