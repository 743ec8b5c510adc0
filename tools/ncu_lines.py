"""Aggregate `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per CUDA source line:
instructions executed and stall samples. usage: ncu_lines.py dump.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fname = ''
agg = []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        fname = r[1].split('/')[-1]; continue
    if len(r) >= 8 and r[0].isdigit():
        try:
            agg.append((float(r[7]), int(r[6]) if r[6].isdigit() else 0, fname, int(r[0]), r[1].strip()[:120]))
        except ValueError:
            pass
tot = sum(a[0] for a in agg); smp = sum(a[1] for a in agg)
print("total warp instructions %.4g, samples %d" % (tot, smp))
agg.sort(reverse=True)
for ie, s, f, ln, src in agg[:top]:
    print("%5.1f%% inst %5.1f%% smp  %s:%-4d %s" % (100 * ie / tot, 100.0 * s / max(smp, 1), f, ln, src))
