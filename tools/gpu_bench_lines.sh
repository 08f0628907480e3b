#!/bin/bash
# bench lines of the five BASELINE configurations + steady-state legs, with the stamped counters of profiles/ in place
tag=${1:-r2}; mkdir -p gpurun_out
for w in ms_pacman_figar10_n16384 pong_paac_n32 breakout_figar10_n256 seaquest_figar10_rgb_n4096 mixed12_figar10_n16384; do
  extra=""; [ $w = breakout_figar10_n256 ] && extra="--steady-state 200"
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 $extra > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
  tail -c 200 gpurun_out/${tag}_bench_$w.json
done
timeout 900 python bench.py --workload yars_revenge_figar10_n4096 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --steady-state 200 > gpurun_out/${tag}_bench_yars_steady.json 2> gpurun_out/${tag}_bench_yars_steady.err
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/${tag}_reference_arm.json 2> /dev/null
