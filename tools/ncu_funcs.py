#!/usr/bin/env python
"""Aggregates an `ncu --page source --print-source cuda,sass --csv` dump by function of emu_core.cuh
(warp-level instructions executed, thread-level instructions, samples)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
src_path = sys.argv[2] if len(sys.argv) > 2 else '/root/repo/manette_b200/csrc/emu_core.cuh'
cur, H = None, None
by = collections.Counter(); thr = collections.Counter(); smp = collections.Counter()
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No':
        H = r; continue
    if H and r[0].isdigit() and len(r) >= 9:
        try:
            inst = int(r[7]); s = int(r[6]); t = int(r[8])
        except ValueError:
            continue
        by[(cur, int(r[0]))] += inst; smp[(cur, int(r[0]))] += s; thr[(cur, int(r[0]))] += t
src = open(src_path).read().split('\n')
starts = []
for i, l in enumerate(src, 1):
    m = re.match(r'^MN_HD\s+(MN_NOINLINE|MN_INLINE)?\s*[\w:<>\*&\s]+?\s+(\w+)\(', l)
    if m:
        starts.append((i, m.group(2)))


def fn_of(line):
    name = '?'
    for i, n in starts:
        if i <= line:
            name = n
        else:
            break
    return name


agg = collections.Counter(); aggs = collections.Counter(); aggt = collections.Counter()
for (f, l), v in by.items():
    key = fn_of(l) if f == 'emu_core.cuh' else f
    agg[key] += v; aggs[key] += smp[(f, l)]; aggt[key] += thr[(f, l)]
tot = sum(agg.values()); ts = sum(aggs.values())
print('total warp-inst %d, avg active threads %.2f' % (tot, sum(aggt.values()) / max(tot, 1)))
for k, v in agg.most_common(30):
    print('%-26s %6.2f%% inst %6.2f%% smp  active threads %5.2f' % (k, 100.0 * v / tot, 100.0 * aggs[k] / max(ts, 1), aggt[k] / max(v, 1)))
