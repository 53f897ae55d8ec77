"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by source line."""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur_file = None; hdr = None; out = []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == 'Line No':
        hdr = r; continue
    if hdr and len(r) > 8 and r[0] not in ('', 'Line No'):
        try:
            line = int(r[0]); samples = int(r[6]); inst = int(r[7])
        except ValueError:
            continue
        out.append((samples, inst, cur_file, line, r[1].strip()[:110]))
tot_s = sum(o[0] for o in out); tot_i = sum(o[1] for o in out)
print('total samples', tot_s, 'total warp-inst', tot_i)
print('--- by samples')
for s, i, f, l, src in sorted(out, reverse=True)[:top]:
    print('%5.1f%% smp %5.1f%% inst  %s:%d  %s' % (100. * s / tot_s, 100. * i / tot_i, f, l, src))
