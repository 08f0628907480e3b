#!/bin/bash
# ncu capture of k_round launches of the first measured macro step of a decorrelated run (profile_step brackets
# the measured steps with cudaProfilerStart/Stop): skip 0 = round 0 (full warps), skip 5 = round 5 (half of the envs left)
# usage: tools/gpu_prof_decor.sh <tag> [skip] [count] [envs]
tag=${1:-x}; skip=${2:-0}; count=${3:-2}; envs=${4:-16384}
mkdir -p gpurun_out
python tools/profile_step.py --envs $envs --decorrelate 24 --steps 1 > gpurun_out/decor_plain_$tag.log 2>&1 || { cat gpurun_out/decor_plain_$tag.log; exit 1; }
cat gpurun_out/decor_plain_$tag.log
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_round -s $skip -c $count -f -o gpurun_out/decor_$tag \
    python tools/profile_step.py --envs $envs --decorrelate 24 --steps 1 > gpurun_out/decor_ncu_$tag.log 2>&1
tail -2 gpurun_out/decor_ncu_$tag.log
