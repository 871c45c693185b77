#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel from an ncu report
(ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:NAME > f.csv; python tools/ncu_lines.py f.csv)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
cur = None; hdr = None; out = []
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].strip().isdigit():
        ie = hdr.index('Instructions Executed'); sm = hdr.index('# Samples')
        try: out.append((cur, int(r[0]), int(r[ie]), int(r[sm]), r[1].strip()[:110]))
        except ValueError: pass
tot = sum(o[2] for o in out); tots = sum(o[3] for o in out)
print("total warp instructions %d, samples %d" % (tot, tots))
for f, ln, n, s, src in out:
    if 100.0 * n / tot >= minshare or 100.0 * s / max(1, tots) >= minshare:
        print("%-14s %4d %6.2f%% inst %6.2f%% smp | %s" % (f, ln, 100.0 * n / tot, 100.0 * s / max(1, tots), src))
