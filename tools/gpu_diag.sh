mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
echo "MN_DIAG=2 (wait counters)"; MN_DIAG=2 timeout 300 python tools/profile_step.py --envs 16384 --decorrelate 24 --steps 3 2>&1 | tail -2 | cut -c1-160
echo "MN_DIAG=2 slack 8"; MN_SYNC_SLACK=8 MN_DIAG=2 timeout 300 python tools/profile_step.py --envs 16384 --decorrelate 24 --steps 3 2>&1 | tail -2 | cut -c1-160
echo "MN_DIAG=2 breakout"; MN_DIAG=2 timeout 300 python tools/profile_step.py --game breakout --envs 16384 --decorrelate 24 --steps 3 2>&1 | tail -2 | cut -c1-160
