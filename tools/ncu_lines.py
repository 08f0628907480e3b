#!/usr/bin/env python
"""Aggregates an `ncu --page source --print-source cuda,sass --csv` dump by CUDA source line."""
import csv
import sys
import collections

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
cur_file, cur_fn, H = None, None, None
by_line = collections.OrderedDict()
by_fn = collections.Counter()
smp_fn = collections.Counter()
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name':
        cur_fn = r[1]; continue
    if r[0] == 'Line No':
        H = r; continue
    if H and r[0].isdigit() and len(r) >= 8:
        try:
            inst = int(r[7]); smp = int(r[6])
        except ValueError:
            continue
        key = (cur_file, int(r[0]))
        a = by_line.setdefault(key, [0, 0, r[1].strip()[:100]])
        a[0] += inst; a[1] += smp
tot = sum(a[0] for a in by_line.values()); ts = sum(a[1] for a in by_line.values())
print('total inst', tot, 'samples', ts)
files = collections.Counter(); fs = collections.Counter()
for (f, l), a in by_line.items():
    files[f] += a[0]; fs[f] += a[1]
for f, v in files.most_common():
    print('%-20s %6.2f%% inst %6.2f%% samples' % (f, 100.0 * v / tot, 100.0 * fs[f] / max(ts, 1)))
for (f, l), a in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%5.2f%% inst %5.2f%% smp  %s:%d  %s' % (100.0 * a[0] / tot, 100.0 * a[1] / max(ts, 1), f, l, a[2]))
