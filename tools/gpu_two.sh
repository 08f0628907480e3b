#!/bin/bash
# two GPUs of one box: the default bench under torchrun (gradient + episode-statistics all-reduce inside the timed
# region, states written into each rank's pinned host array during the step) and the 2-rank training test
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_default.json 2> gpurun_out/bench_2gpu_default.err
tail -c 900 gpurun_out/bench_2gpu_default.json
timeout 900 python -m pytest tests/test_gpu_train_eval.py -x -q -k two_rank 2>&1 | tail -3
