"""Instruction count of every out-of-line function inside a kernel's SASS (cuobjdump -sass)."""
import re, subprocess, sys
out = subprocess.run(['cuobjdump', '-sass', sys.argv[1]], capture_output=True, text=True).stdout
cur = 'kernel-body'; counts = {}
for line in out.splitlines():
    m = re.match(r'\s+(\$\S+):', line)
    if m:
        cur = m.group(1).split('$')[-1]
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', line):
        counts[cur] = counts.get(cur, 0) + 1
for k, v in sorted(counts.items(), key=lambda kv: -kv[1]):
    print('%7d  %s' % (v, k))
print('%7d  total' % sum(counts.values()))
