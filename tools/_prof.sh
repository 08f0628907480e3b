python tools/profile_step.py --game ms_pacman --envs 16384 --decorrelate 24 --steps 3 --envs-per-warp 32 > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:k_round --launch-skip 390 --launch-count 1 -o gpurun_out/prof_r1f -f \
  python tools/profile_step.py --game ms_pacman --envs 16384 --decorrelate 24 --steps 1 --envs-per-warp 32 > gpurun_out/ncu_r1f.log 2>&1
