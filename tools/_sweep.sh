for g in asterix asteroids breakout enduro gopher gravitar montezuma_revenge ms_pacman pong seaquest space_invaders yars_revenge; do
timeout 300 python tools/profile_step.py --game $g --envs 16384 --decorrelate 24 --steps 4
done
