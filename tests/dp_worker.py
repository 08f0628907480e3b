"""Worker of tests/test_gpu_train_eval.py::test_two_rank_training_keeps_replicas_identical (run under torchrun):
trains Pong for two updates with one process per GPU and writes what the test compares."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if __name__ == "__main__":
    out = sys.argv[1]
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    from manette_b200 import train
    args = train.get_arg_parser().parse_args([
        "-g", "pong", "--rom_path", os.path.join(ROOT, "atari_roms"), "-df", os.path.join(out, "run") + "/", "--arch", "NIPS",
        "--max_repetition", "10", "--nb_choices", "11", "-ec", "32", "--max_global_steps", str(2 * 5 * 32 * dist.get_world_size()),
        "--checkpoint_interval", "100000"])
    torch.manual_seed(100 + rank)                       # different initial weights: rank 0's must win
    learner = train.main(args)
    stats = learner.episode_statistics()
    digest = [float(p.detach().double().sum()) for p in learner.network.parameters()]
    states = int(learner.pool.states.to(torch.int64).sum())
    with open(os.path.join(out, "rank%d.json" % rank), "w") as fh:
        json.dump({"global_step": learner.global_step, "digest": digest, "steps_recorded": float(stats[5]),
                   "states": states, "offset": learner.pool.env_id_offset}, fh)
    dist.barrier()
    dist.destroy_process_group()
