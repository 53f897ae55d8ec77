"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by source line
and by function (source line ranges are resolved from the .cu file's function headers)."""
import csv, re, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src_file = sys.argv[3] if len(sys.argv) > 3 else None
rows = list(csv.reader(open(path)))
cur_file = None; names = None; out = []; stalls = {}
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No':
        names = r; continue
    if names and len(r) > 8 and r[0] not in ('', 'Line No'):
        try:
            line = int(r[0]); samples = int(r[6]); inst = int(r[7])
        except ValueError:
            continue
        out.append((samples, inst, cur_file, line, r[1].strip()[:110]))
    elif names and len(r) > 48 and r[0] == '':
        for i in range(32, 49):
            try: stalls[names[i]] = stalls.get(names[i], 0) + int(r[i])
            except ValueError: pass
tot_s = sum(o[0] for o in out) or 1; tot_i = sum(o[1] for o in out) or 1
print('total samples', tot_s, 'total warp-inst', tot_i)
T = sum(stalls.values()) or 1
print('stalls:', ', '.join('%s %.1f%%' % (k[6:], 100. * v / T) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]))
if src_file:
    funcs = []
    for i, l in enumerate(open(src_file), 1):
        m = re.match(r'^(?:template.*)?(?:__device__|__global__).*?([A-Za-z_0-9]+)\(', l)
        if m: funcs.append((i, m.group(1)))
    base = src_file.split('/')[-1]
    agg = {}
    for s, i, f, l, _ in out:
        k = f
        if f == base:
            k = 'prologue'
            for a, n in funcs:
                if a <= l: k = n
        x = agg.setdefault(k, [0, 0]); x[0] += s; x[1] += i
    print('--- by function')
    for k, (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
        print('%-34s %5.1f%% samples %5.1f%% inst' % (k, 100. * s / tot_s, 100. * i / tot_i))
print('--- by samples')
for s, i, f, l, src in sorted(out, reverse=True)[:top]:
    print('%5.1f%% smp %5.1f%% inst  %s:%d  %s' % (100. * s / tot_s, 100. * i / tot_i, f, l, src))
