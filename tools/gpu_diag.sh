mkdir -p gpurun_out
for v in base nbuf8 poll both; do
  cp tools/_variants/$v.so manette_b200/libmanette_b200.so; touch manette_b200/libmanette_b200.so
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python -c "
import json; d=json.loads(open('gpurun_out/var_$v.json').read().strip().splitlines()[-1]); print('$v ms_pacman bench', int(d['value']), d['ms_per_step'])"
  for g in pong ms_pacman gravitar; do
    MN_DIAG=2 timeout 200 python tools/profile_step.py --game $g --envs 16384 --decorrelate 24 --steps 4 2>&1 | tail -2 | cut -c1-400
  done
done
