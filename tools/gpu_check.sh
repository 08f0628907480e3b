#!/bin/bash
# One gpurun call: GPU parity tests, a device-timed throughput line, and (optionally) one ncu capture of k_round.
# usage: tools/gpu_check.sh <tag> [tests|notests] [ncu|noncu]
tag=${1:-x}; tests=${2:-tests}; ncu=${3:-ncu}
mkdir -p gpurun_out
if [ "$tests" = tests ]; then
  python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$tag.log
  tail -3 gpurun_out/pytest_$tag.log
fi
python tools/profile_step.py --envs 16384 --decorrelate 24 --steps 4 > gpurun_out/step_$tag.log 2>&1; cat gpurun_out/step_$tag.log
python tools/profile_step.py --envs 32768 --decorrelate 24 --steps 4 >> gpurun_out/step_$tag.log 2>&1; tail -1 gpurun_out/step_$tag.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -c 1500 gpurun_out/bench_$tag.json
if [ "$ncu" = ncu ]; then
  python tools/profile_step.py --envs 16384 --decorrelate 6 --steps 1 > gpurun_out/prof_plain_$tag.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_round -s 102 -c 2 -f -o gpurun_out/prof_$tag \
      python tools/profile_step.py --envs 16384 --decorrelate 6 --steps 1 > gpurun_out/ncu_$tag.log 2>&1
  tail -2 gpurun_out/ncu_$tag.log
fi
