# scratch script for one-off GPU checks (gpurun -- 'bash tools/gpu_diag.sh'): build, smoke, the default bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_default_final.json 2> gpurun_out/bench_default_final.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_default_final.json').read().strip().splitlines()[-1])
print(int(d['value']), int(d['e2e']['value']), d['steps'], d['warmup'], d['roofline']['frac'], d['roofline']['counters'].get('source_hash'), d['clocks'], d['cpu_baseline']['value'])"
