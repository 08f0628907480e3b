mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_runners.py tests/test_gpu_properties.py tests/test_gpu_parity.py -x -q -k "random_start or sharded or memo or warmed" 2>&1 | tail -3
timeout 900 python bench.py --workload breakout_figar10_n256 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --random-start --steady-state 600 > gpurun_out/steady_c2_rs.json 2> gpurun_out/steady_c2_rs.err; python -c "
import json; d=json.loads(open('gpurun_out/steady_c2_rs.json').read().strip().splitlines()[-1]); print('C2 random_start', int(d['value']), d['steady_state'], d['reset_memo'])"
