"""usage: sass_stats.py <object/.so> <function-substring>: opcode histogram of one kernel's SASS (static count)."""
import collections, re, subprocess, sys
out = subprocess.run(['cuobjdump', '-sass', sys.argv[1]], capture_output=True, text=True).stdout
on, cnt = False, collections.Counter()
for line in out.splitlines():
    if 'Function :' in line:
        on = sys.argv[2] in line
        continue
    if not on:
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(.*?);', line)
    if not m:
        continue
    ins = re.sub(r'^@!?U?P\w+\s+', '', m.group(1).strip())
    cnt[ins.split()[0].split('.')[0]] += 1
tot = sum(cnt.values())
print(tot, 'TOTAL')
for k, v in cnt.most_common(25):
    print('%6d %s' % (v, k))
