#!/bin/bash
# quick parameter sweeps of the emulation kernel on one GPU (device-timed frames/s, Ms Pacman, decorrelated)
mkdir -p gpurun_out
out=gpurun_out/sweep_${1:-x}.log; : > $out
for s in 2 4 8 16 64 100000; do
  echo "slack=$s" >> $out
  MN_SYNC_SLACK=$s python tools/profile_step.py --envs 16384 --decorrelate 24 --steps 3 2>&1 | tail -1 >> $out
done
for w in 16 8; do
  echo "envs_per_warp=$w" >> $out
  python tools/profile_step.py --envs 16384 --decorrelate 24 --steps 3 --envs-per-warp $w 2>&1 | tail -1 >> $out
done
cat $out
