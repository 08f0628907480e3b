#!/bin/bash
# full GPU suite (bounded), then the default bench line
tag=${1:-x}; mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/full_pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/full_pytest_$tag.log
tail -15 gpurun_out/full_pytest_$tag.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -c 600 gpurun_out/bench_$tag.json
