#!/bin/bash
# ncu capture of k_round launches deep inside a decorrelated run (macro step 25 of profile_step: launches 1 + 24 x 11 on).
# usage: tools/gpu_prof_decor.sh <tag> [skip] [count] [envs]
tag=${1:-x}; skip=${2:-265}; count=${3:-2}; envs=${4:-16384}
mkdir -p gpurun_out
python tools/profile_step.py --envs $envs --decorrelate 24 --steps 1 > gpurun_out/decor_plain_$tag.log 2>&1 || { cat gpurun_out/decor_plain_$tag.log; exit 1; }
cat gpurun_out/decor_plain_$tag.log
ncu --set full --clock-control none --import-source on -k 'regex:k_round<.bool.0>' -s $skip -c $count -f -o gpurun_out/decor_$tag \
    python tools/profile_step.py --envs $envs --decorrelate 24 --steps 1 > gpurun_out/decor_ncu_$tag.log 2>&1
tail -2 gpurun_out/decor_ncu_$tag.log
