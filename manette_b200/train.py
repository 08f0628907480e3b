"""Command line of the reference's train.py (train.py:84-130) on the device pool.

    python -m manette_b200.train -g pong -df logs/ [the reference's flags]

Every flag, destination and default of `get_arg_parser()` (train.py:84-120) is kept, so an `args.json` written by
either side drives the other; `main(args)` follows train.py:15-29: exploration policy -> network / environment
creators -> PAACLearner -> train() with SIGINT/SIGTERM saving a checkpoint first (train.py:32-43).
`-d/--device` accepts the reference's TensorFlow names ('/gpu:1', '/cpu:0') and maps '/gpu:k' to cuda:k; there is
no CPU path, so '/cpu:0' (the reference's default) selects cuda:0.  Extra flags (absent from the reference, all
optional): --seed (counter-based sampler key), --envs_per_warp (kernel layout override), --micro_batch (frames per
network slice)."""
import argparse
import copy
import logging
import os
import signal
import sys

# (flags, dest, default, type or None for store_true, help)
_FLAGS = [
    (("-g",), "game", "pong", str, "game (ROM file name without .bin)"),
    (("-d", "--device"), "device", "/cpu:0", str, "'/gpu:k' -> cuda:k; '/cpu:0' -> cuda:0 (no CPU path exists)"),
    (("--rom_path",), "rom_path", "./atari_roms", str, "directory holding the ROM images"),
    (("-v", "--visualize"), "visualize", 0, int, "0 none; 1 every emulator calls on_new_frame; 2 one emulator"),
    (("--e",), "e", 0.1, float, "RMSProp epsilon (inside the square root, TensorFlow's form)"),
    (("--alpha",), "alpha", 0.99, float, "RMSProp decay"),
    (("-lr", "--initial_lr"), "initial_lr", 0.0224, float, "initial learning rate"),
    (("-lra", "--lr_annealing_steps"), "lr_annealing_steps", 80000000, int, "global steps over which lr goes to 0"),
    (("--entropy",), "entropy_regularisation_strength", 0.02, float, "entropy bonus weight"),
    (("--clip_norm",), "clip_norm", 3.0, float, "gradient norm bound"),
    (("--clip_norm_type",), "clip_norm_type", "global", str, "ignore | local | global"),
    (("--gamma",), "gamma", 0.99, float, "discount"),
    (("--max_global_steps",), "max_global_steps", 80000000, int, "training length in global steps"),
    (("--max_local_steps",), "max_local_steps", 5, int, "macro steps per update (n-step horizon)"),
    (("--arch",), "arch", "PWYX", str, "NIPS | NATURE | PWYX | LSTM | BAYESIAN"),
    (("--single_life_episodes",), "single_life_episodes", False, None, "losing a life ends the episode"),
    (("-ec", "--emulator_counts"), "emulator_counts", 32, int, "environments (per GPU)"),
    (("-ew", "--emulator_workers"), "emulator_workers", 8, int, "kept for parity: must divide emulator_counts"),
    (("-df", "--debugging_folder"), "debugging_folder", "logs/", str, "run folder (args.json, checkpoints/)"),
    (("-rs", "--random_start"), "random_start", False, None, "0..30 no-op frames after every reset"),
    (("--egreedy",), "egreedy", False, None, "epsilon-greedy instead of sampling the softmax"),
    (("--epsilon",), "epsilon", 0.05, float, "epsilon of --egreedy"),
    (("--softmax_temp",), "softmax_temp", 1.0, float, "temperature of both policy heads"),
    (("--annealed",), "annealed", False, None, "anneal epsilon linearly"),
    (("--annealed_steps",), "annealed_steps", 80000000, int, "global steps of the epsilon annealing"),
    (("--keep_percentage",), "keep_percentage", 0.9, float, "dropout keep probability (BAYESIAN)"),
    (("--rgb",), "rgb", False, None, "colour observations 84x84x12 instead of 84x84x4"),
    (("--max_repetition",), "max_repetition", 0, int, "FiGAR: largest repetition count"),
    (("--nb_choices",), "nb_choices", 1, int, "FiGAR: size of the repetition head"),
    (("--checkpoint_interval",), "checkpoint_interval", 1000000, int, "global steps between checkpoints"),
    (("--activation",), "activation", "relu", str, "relu | leaky_relu"),
    (("--alpha_leaky_relu",), "alpha_leaky_relu", 0.1, float, "leaky slope"),
]
_EXTRA = [
    (("--seed",), "seed", 0, int, "key of the counter-based action sampler"),
    (("--envs_per_warp",), "envs_per_warp", 0, int, "emulation kernel layout override (0 = automatic)"),
    (("--amp",), "amp", False, None, "bf16 autocast of the convolution stack (the reference computes in fp32)"),
    (("--micro_batch",), "micro_batch", 16384, int, "frames per forward/backward slice of the network (memory bound)"),
]


def get_arg_parser():
    parser = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    for flags, dest, default, typ, text in _FLAGS + _EXTRA:
        if typ is None:
            parser.add_argument(*flags, dest=dest, action="store_true", help=text)
        else:
            parser.add_argument(*flags, dest=dest, default=default, type=typ, help=text)
    return parser


def cuda_index(device):
    """'/gpu:3' -> 3; anything else (the reference's '/cpu:0' default) -> LOCAL_RANK or 0."""
    d = str(device)
    if "gpu:" in d:
        return int(d[d.rindex(":") + 1:])
    return int(os.environ.get("LOCAL_RANK", "0"))


def get_network_and_environment_creator(args, explo_policy, random_seed=3):
    """train.py:46-82: sets args.num_actions / args.random_seed, returns (network_creator, env_creator)."""
    from .environment_creator import EnvironmentCreator
    from .networks import PolicyVNetwork
    env_creator = EnvironmentCreator(args)
    args.num_actions = env_creator.num_actions
    args.random_seed = random_seed
    args.cuda_device = cuda_index(args.device)
    conf = {"arch": args.arch if args.arch in PolicyVNetwork.ARCHS else "NATURE",     # train.py:69-78: the else branch
            "num_actions": args.num_actions, "nb_choices": args.nb_choices, "depth": 3 if args.rgb else 1,
            "softmax_temp": explo_policy.softmax_temp, "activation": args.activation,
            "alpha_leaky_relu": args.alpha_leaky_relu, "keep_percentage": explo_policy.keep_percentage,
            "entropy_regularisation_strength": args.entropy_regularisation_strength, "amp": bool(getattr(args, "amp", False))}

    def network_creator(name="local_learning"):
        net = PolicyVNetwork(**copy.copy(conf))
        net.name = name
        return net

    return network_creator, env_creator


def setup_kill_signal_handler(learner):
    owner = os.getpid()

    def on_signal(signum, frame):
        if os.getpid() == owner:
            logging.info("Signal %s detected, cleaning up.", signum)
            learner.cleanup()
            logging.info("Cleanup completed, shutting down...")
            sys.exit(0)

    signal.signal(signal.SIGTERM, on_signal)
    signal.signal(signal.SIGINT, on_signal)


def main(args):
    from .exploration_policy import ExplorationPolicy
    from .paac import PAACLearner
    logging.debug("Configuration: %s", args)
    explo_policy = ExplorationPolicy(args, seed=getattr(args, "seed", 0))
    print("Repetition table : " + str(explo_policy.tab_rep))
    network_creator, env_creator = get_network_and_environment_creator(args, explo_policy)
    learner = PAACLearner(network_creator, env_creator, explo_policy, args)
    setup_kill_signal_handler(learner)
    logging.info("Starting training")
    learner.train()
    logging.info("Finished training")
    return learner


if __name__ == "__main__":
    logging.basicConfig(stream=sys.stdout, level=logging.INFO)
    cli = get_arg_parser().parse_args()
    from . import logger_utils
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:      # torchrun: one process per GPU, synchronous PAAC over NCCL
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    if int(os.environ.get("RANK", "0")) == 0:
        logger_utils.save_args(cli, cli.debugging_folder)
    main(cli)
