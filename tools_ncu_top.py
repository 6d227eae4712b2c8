#!/usr/bin/env python3
"""tools_ncu_top.py REPORT.ncu-rep [N] — the N SASS lines with the most warp-stall samples of the first kernel in a report."""
import csv, io, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h = r[1]; rows = r[2:]
iS = h.index('# Samples'); iI = h.index('Instructions Executed'); iSrc = h.index('Source'); iW = h.index('L1 Wavefronts Shared'); iWi = h.index('L1 Wavefronts Shared Ideal')
cols = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
tot = sum(int(x[iS]) for x in rows)
print('total samples', tot, 'warp instructions', sum(int(x[iI]) for x in rows))
top = sorted(range(len(rows)), key=lambda k: -int(rows[k][iS]))[:n]
for k in sorted(top):
    x = rows[k]
    st = sorted(((int(x[i] or 0), h[i][6:]) for i in cols), reverse=True)[:2]
    print('%5d %-62s samples %5s execs %8s wf %8s/%8s  %s' % (k, x[iSrc].strip()[:62], x[iS], x[iI], x[iW], x[iWi], ' '.join('%s:%d' % (b, a) for a, b in st if a)))
