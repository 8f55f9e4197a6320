#!/usr/bin/env python
"""Per-source-line instruction counts of the first kernel in an .ncu-rep (source page, cuda+sass).  usage: ncu_lines.py rep [topN] [maxthreads-for-'decoder'-split]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 50
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
cur = None; hdr = None; agg = {}; funcs = []
for r in csv.reader(io.StringIO(raw)):
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name':
        if r[1] not in funcs: funcs.append(r[1])
        continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] != '':
        try: ie = int(r[7]); ti = int(r[8]); smp = int(r[6])
        except ValueError: continue
        a = agg.setdefault((cur, int(r[0])), [0, 0, 0, r[1]])
        a[0] += ie; a[1] += smp; a[2] += ti
tot = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print(funcs); print('total warp-instructions', tot, 'samples', ts)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print('%-22s %4d %12d %5.1f%% thr/inst %5.1f smp %5.1f%%  %s' % (k[0][:22], k[1], a[0], 100 * a[0] / tot, a[2] / max(a[0], 1), 100 * a[1] / max(ts, 1), a[3].strip()[:100]))
