#!/bin/bash
# N GPUs of one box: the default bench under torchrun (collective inside the timed region)
mkdir -p gpurun_out
N=${1:-4}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu_default.json 2> gpurun_out/bench_${N}gpu_default.err
tail -c 1200 gpurun_out/bench_${N}gpu_default.json
