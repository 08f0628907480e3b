for g in ms_pacman seaquest breakout; do
timeout 300 python tools/profile_step.py --game $g --envs 16384 --decorrelate 24 --steps 3 --envs-per-warp 32
done
python -m pytest tests/test_gpu_parity.py -x -q -k "ms_pacman or breakout" 2>&1 | tail -3
