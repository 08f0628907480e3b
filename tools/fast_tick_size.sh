#!/bin/bash
# SASS instruction count of the fast tick (tools/fast_probe.cu: cpu_fast in a bare loop; ~25 instructions are the harness)
set -e
mkdir -p /tmp/probe
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -cubin -o /tmp/probe/p.cubin "$(dirname "$0")/fast_probe.cu"
cuobjdump -sass /tmp/probe/p.cubin > /tmp/probe/p.sass
python3 - <<'PY'
import re
name, n = None, 0
for l in open('/tmp/probe/p.sass'):
    m = re.search(r'Function : (\S+)', l)
    if m:
        if name: print(name, n)
        name, n = m.group(1), 0
    elif re.match(r'\s+/\*[0-9a-f]{4}\*/', l):
        n += 1
if name: print(name, n)
PY
