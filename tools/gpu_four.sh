#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu_default.json 2> gpurun_out/bench_${N}gpu_default.err
tail -c 1200 gpurun_out/bench_${N}gpu_default.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29572 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | tail -c 300
