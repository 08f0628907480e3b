#!/bin/bash
# every README game through pool creation, resets and a few FiGAR macro steps at a given pool size (default 4096)
n=${1:-4096}
for g in asterix asteroids breakout enduro gopher gravitar montezuma_revenge ms_pacman pong seaquest space_invaders yars_revenge; do
  python tools/profile_step.py --game $g --envs $n --decorrelate 3 --steps 2 2>&1 | tail -1 | cut -c1-150
done
