#!/usr/bin/env python
"""PAAC + FiGAR updates per second with the learner in the loop (SURVEY 8(f) rank 2 at the pool sizes of the bench):
net forward -> K4 sampling -> macro step -> K6, x T; bootstrap -> K5 -> loss -> backward -> (NCCL mean) -> clip -> RMSProp.

    python tools/bench_learner.py --game ms_pacman --envs 16384 --arch LSTM [--updates 3] [--warmup 1]
    python -m torch.distributed.run --nproc-per-node 2 tools/bench_learner.py ...

Prints one JSON line (rank 0): preprocessed frames/s (next() calls of all ranks / max-over-ranks device time)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import manette_b200 as mb  # noqa: E402
from manette_b200.learner import PAACLearner  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--game", default="ms_pacman")
    ap.add_argument("--envs", type=int, default=16384)
    ap.add_argument("--arch", default="NIPS")
    ap.add_argument("--rgb", action="store_true")
    ap.add_argument("--updates", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--micro-batch", type=int, default=16384)
    ap.add_argument("--amp", action="store_true", help="bf16 autocast of the convolution stack")
    ap.add_argument("--tf32", action="store_true", help="TF32 matmuls / convolutions")
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    tab = mb.tab_repetitions(10, 11)
    pool = mb.DevicePool([(a.game, mb.load_rom(os.path.join(ROOT, "atari_roms"), a.game), a.envs)], rgb=a.rgb, tab_rep=tab,
                         device=local, env_id_offset=rank * a.envs, history=5 if a.arch.upper() == "LSTM" else 0)
    pool.reset_all()
    torch.manual_seed(0)
    if a.tf32:
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    learner = PAACLearner(pool, arch=a.arch, seed=rank, micro_batch=a.micro_batch, amp=a.amp)
    for _ in range(a.warmup):
        out = learner.train_rollout()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier(device_ids=[local])
    f0 = pool.total_next_calls()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.updates):
        out = learner.train_rollout()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    frames = torch.tensor([float(pool.total_next_calls() - f0)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(frames)
    if rank == 0:
        n_par = sum(p.numel() for p in learner.network.parameters())
        print(json.dumps({"metric": "preprocessed env frames/sec with the PAAC learner in the loop", "value": float(frames) / (float(ms) / 1e3),
                          "unit": "frames/s", "n_gpus": world, "arch": a.arch, "game": a.game, "envs_per_gpu": a.envs,
                          "updates": a.updates, "ms_per_update": float(ms) / a.updates, "macro_steps_per_update": learner.T,
                          "agent_steps_per_s": a.updates * learner.T * a.envs * world / (float(ms) / 1e3),
                          "parameters": n_par, "micro_batch_frames": a.micro_batch, "amp": a.amp, "tf32": a.tf32, "loss": float(out["loss"]),
                          "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30}))
    learner.close(); pool.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
