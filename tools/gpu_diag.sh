mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/knob_$tag.json 2> gpurun_out/knob_$tag.err; python -c "
import json; d=json.loads(open('gpurun_out/knob_$tag.json').read().strip().splitlines()[-1]); print('$tag', int(d['value']), d['ms_per_step'])"; }
run base MN_EARLY_EMIT=1
run high6 MN_FIFO_HIGH=6
run high8 MN_FIFO_HIGH=8
run high10 MN_FIFO_HIGH=10
run high14 MN_FIFO_HIGH=14
run slack8 MN_SYNC_SLACK=8
