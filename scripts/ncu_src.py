"""Summarise an ncu `--page source --csv --print-source cuda,sass` dump per CUDA source line.
usage: python scripts/ncu_src.py dump.csv [min_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.8
def f(x):
    try:
        return float(x.replace(',', '')) if x not in ('', '-') else 0.0
    except ValueError:
        return 0.0
lines = []
fname = ''
H = None
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]; continue
    if r[0] == 'Line No':
        H = r; si = H.index('# Samples'); ii = H.index('Instructions Executed'); continue
    if H is None or len(r) <= ii:
        continue
    if r[0] not in ('', '-') and r[0].isdigit():
        lines.append((fname, int(r[0]), r[1], f(r[si]), f(r[ii])))
ts = sum(l[3] for l in lines); ti = sum(l[4] for l in lines)
print("total samples %.0f, warp instructions %.0f" % (ts, ti))
for l in lines:
    if l[3] > ts * minp / 100 or l[4] > ti * minp / 100:
        print("%-16s %4d  samp %5.1f%%  inst %5.1f%%  %s" % (l[0], l[1], l[3] / ts * 100, l[4] / ti * 100, l[2].strip()[:100]))
