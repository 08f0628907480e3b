#!/bin/bash
# quick parameter sweeps of the emulation kernel on one GPU (device-timed frames/s, Ms Pacman, decorrelated)
mkdir -p gpurun_out
out=gpurun_out/sweep_${1:-x}.log; : > $out
for s in 4 8 16 32 64 256 100000; do
  echo "slack=$s" >> $out
  MN_SYNC_SLACK=$s timeout 300 python tools/profile_step.py --envs 16384 --decorrelate 24 --steps 3 2>&1 | tail -1 >> $out
done
for hh in 4 8; do
  echo "fifo_high=$hh" >> $out
  MN_FIFO_HIGH=$hh timeout 300 python tools/profile_step.py --envs 16384 --decorrelate 24 --steps 3 2>&1 | tail -1 >> $out
done
for s in 16 64; do
  echo "slack=$s fifo_high=8" >> $out
  MN_SYNC_SLACK=$s MN_FIFO_HIGH=8 timeout 300 python tools/profile_step.py --envs 16384 --decorrelate 24 --steps 3 2>&1 | tail -1 >> $out
done
for g in pong seaquest breakout; do
  for s in 4 32; do
    echo "game=$g slack=$s" >> $out
    MN_SYNC_SLACK=$s timeout 300 python tools/profile_step.py --game $g --envs 16384 --decorrelate 24 --steps 3 2>&1 | tail -1 >> $out
  done
done
cut -c1-130 $out
