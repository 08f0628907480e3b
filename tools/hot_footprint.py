#!/usr/bin/env python
"""Instruction-cache footprint of k_round by source function (the footprint is a first-class cost).

    python tools/hot_footprint.py [--sass-csv profile_sass.csv] [--lib manette_b200/libmanette_b200.so] [--kernel k_roundILb0]

Static part (always): bytes of SASS per function of emu_core.cuh / pool.cu, from `cuobjdump -xelf` + `nvdisasm -g -c`
(needs the library built with -lineinfo, which build.py does).
Dynamic part (with --sass-csv, the output of `ncu -i X.ncu-rep --page source --print-source sass --csv` of a
`--set full --import-source on` capture of the SAME build): a 128-byte instruction line counts as "touched t times per
tick" when its most executed instruction ran t x (number of ticks) times, a tick being one pass of the flat loop
(= executions of the loop's CREDUX); per function: static bytes, bytes in lines touched more than once per 100 / per
10 ticks, and the share of executed warp instructions."""
import argparse
import collections
import csv
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def function_ranges(path):
    out = []
    for i, line in enumerate(open(path).read().splitlines(), 1):
        m = re.match(r'^\s*MN_HD\s+(?:MN_INLINE\s+|MN_NOINLINE\s+|MN_NOINLINE_DEV\s+)?(?:const\s+)?[\w:<>]+[\s\*&]+(\w+)\s*\(', line)
        if m:
            out.append((i, m.group(1)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "manette_b200", "libmanette_b200.so"))
    ap.add_argument("--kernel", default="k_roundILb0")
    ap.add_argument("--sass-csv", default=None)
    ap.add_argument("--top", type=int, default=36)
    a = ap.parse_args()
    funcs = function_ranges(os.path.join(ROOT, "manette_b200", "csrc", "emu_core.cuh"))

    def fn_of(file, line):
        if not file.endswith("emu_core.cuh"):
            return os.path.basename(file)
        name = "?"
        for start, n in funcs:
            if start <= line:
                name = n
            else:
                break
        return name

    with tempfile.TemporaryDirectory() as d:
        subprocess.run("cd %s && cuobjdump -xelf all %s >/dev/null 2>&1" % (d, a.lib), shell=True, check=True)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        dis = subprocess.run("cd %s && nvdisasm -g -c %s 2>/dev/null" % (d, cubin), shell=True, capture_output=True, text=True).stdout
    on, cur, amap = False, ("?", 0), {}
    for l in dis.splitlines():
        if l.startswith("//---") and ".text." in l:
            if on and amap:
                break
            on = a.kernel in l
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,6})\*/', l)
        if m:
            amap[int(m.group(1), 16)] = cur
    stat = collections.defaultdict(lambda: [0, 0, 0, 0])
    for addr, where in amap.items():
        stat[fn_of(*where)][0] += 16
    ticks = 0
    if a.sass_csv:
        rows = list(csv.reader(open(a.sass_csv)))
        H = rows[1]
        ix = {n: i for i, n in enumerate(H)}
        base, ex = None, {}
        for r in rows[2:]:
            if len(r) < 10 or not r[0].startswith("0x"):
                continue
            addr = int(r[0], 16)
            base = addr if base is None else base
            ex[addr - base] = int(r[ix["Instructions Executed"]])
            if "CREDUX" in r[1]:
                ticks += int(r[ix["Instructions Executed"]])
        linemax = collections.defaultdict(int)
        for addr, e in ex.items():
            linemax[addr >> 7] = max(linemax[addr >> 7], e)
        for addr, e in ex.items():
            s = stat[fn_of(*amap.get(addr, ("?", 0)))]
            s[3] += e
            t = linemax[addr >> 7] / max(ticks, 1)
            s[1] += 16 if t > 0.01 else 0
            s[2] += 16 if t > 0.1 else 0
        tot = sum(s[3] for s in stat.values())
        print("ticks %d, warp instructions per tick %.1f, distinct lines per tick %.1f"
              % (ticks, tot / max(ticks, 1), sum(v / max(ticks, 1) for v in linemax.values())))
    print("%-24s %8s %9s %8s %7s" % ("function", "static B", "hot>.01 B", "hot>.1 B", "dyn %"))
    tot = max(1, sum(s[3] for s in stat.values()))
    key = (lambda kv: -kv[1][1]) if a.sass_csv else (lambda kv: -kv[1][0])
    for f, s in sorted(stat.items(), key=key)[:a.top]:
        print("%-24s %8d %9d %8d %6.2f%%" % (f, s[0], s[1], s[2], 100.0 * s[3] / tot))
    print("TOTAL %d B static, %d B hot>.01, %d B hot>.1" % tuple(sum(s[i] for s in stat.values()) for i in range(3)))


if __name__ == "__main__":
    main()
