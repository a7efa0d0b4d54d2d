#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source sass` export: stall reasons over the hot loop
(instructions executed >= 90 % of the maximum) and the hottest instructions."""
import csv
import sys
import collections

rows = list(csv.reader(open(sys.argv[1])))
# the export holds one section per kernel: a "Kernel Name" line, a header line, then the instructions;
# argv[3] (default 0) picks the section
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
lo = starts[which]
hi = starts[which + 1] if which + 1 < len(starts) else len(rows)
print(rows[lo][1][:100])
hdr = rows[lo + 1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[lo + 2:hi] if len(r) >= len(hdr) - 1]
ie = [int(r[ix['Instructions Executed']] or 0) for r in data]
mx = max(ie)
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
reg = collections.Counter()
agg = collections.Counter()
hot = [i for i, v in enumerate(ie) if v >= mx * 0.9]
for i, r in enumerate(data):
    s = int(r[ix['# Samples']] or 0)
    lvl = 'hot' if ie[i] >= mx * 0.9 else ('20-90%' if ie[i] >= mx * 0.2 else ('1-20%' if ie[i] >= mx * 0.01 else 'rare'))
    reg[lvl] += s
    if lvl == 'hot':
        for k in stalls:
            agg[k] += int(r[ix[k]] or 0)
print('samples', tot, 'regions', dict(reg), 'hot instrs', len(hot), 'range', hot[0], hot[-1])
sel = agg['stall_selected'] / max(1, len(hot))
print('samples per issue slot (selected/instr): %.0f' % sel)
for k, v in agg.most_common(10):
    print('  %-28s %8d  %5.1f%%  %6.1f cycles/iter' % (k, v, 100 * v / sum(agg.values()), v / sel))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 6.0
for i in range(hot[0], hot[-1] + 1):
    r = data[i]
    s = int(r[ix['# Samples']] or 0)
    if s >= thr * sel:
        top = sorted(((int(r[ix[k]] or 0), k[6:]) for k in stalls), reverse=True)[:2]
        print(i, '%3d%%' % (ie[i] * 100 // mx), '%6.1f' % (s / sel), r[ix['Source']][:60], top)
