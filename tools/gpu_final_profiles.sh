#!/bin/bash
# Everything profiles/ quotes for the round, from ONE build: launch list of a bench run, ncu --set full of k_round
# (decorrelated round 0 = every env, round 5 = half of them left) and of k_push_frames, bench lines of the five BASELINE
# configurations, per-game throughput.  usage: tools/gpu_final_profiles.sh <tag>
tag=${1:-r2}; mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_launch_plain.json 2> gpurun_out/${tag}_launch_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_launch_ncu.log 2>&1
bash tools/gpu_prof_decor.sh ${tag}_round0 0 1
bash tools/gpu_prof_decor.sh ${tag}_round5 5 1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_push_frames -s 0 -c 1 -f -o gpurun_out/decor_${tag}_k3 \
    python tools/profile_step.py --envs 16384 --decorrelate 24 --steps 1 > gpurun_out/decor_ncu_${tag}_k3.log 2>&1
for w in pong_paac_n32 breakout_figar10_n256 seaquest_figar10_rgb_n4096 ms_pacman_figar10_n16384 mixed12_figar10_n16384; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
  tail -c 300 gpurun_out/${tag}_bench_$w.json
done
for g in asterix asteroids breakout enduro gopher gravitar montezuma_revenge ms_pacman pong seaquest space_invaders yars_revenge; do
  timeout 300 python tools/profile_step.py --game $g --envs 16384 --decorrelate 24 --steps 4 2>&1 | tail -1
done > gpurun_out/${tag}_per_game_n16384.txt
cat gpurun_out/${tag}_per_game_n16384.txt | cut -c1-140
