#!/bin/bash
# quick GPU check of a new k_round build: a few parity tests (bounded), then device-timed frames/s at two pool sizes
tag=${1:-x}; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -x -q > gpurun_out/quick_pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/quick_pytest_$tag.log
tail -4 gpurun_out/quick_pytest_$tag.log
timeout 300 python tools/profile_step.py --envs 16384 --decorrelate 24 --steps 4 2>&1 | tail -1 | tee gpurun_out/quick_step_$tag.log
timeout 300 python tools/profile_step.py --envs 32768 --decorrelate 24 --steps 4 2>&1 | tail -1 | tee -a gpurun_out/quick_step_$tag.log
timeout 300 python tools/profile_step.py --game breakout --envs 256 --decorrelate 24 --steps 8 2>&1 | tail -1 | tee -a gpurun_out/quick_step_$tag.log
