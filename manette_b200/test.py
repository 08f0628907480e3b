"""Evaluation of a trained run folder: the reference's test.py (test.py:27-116) on the device pool.

    python -m manette_b200.test -f logs/ -tc 50 [-np 30] [-gn name -gf folder] [-d /gpu:0]

Same flags; the stored `args.json` is overlaid on them (test.py:38-40), episodes are whole games without random
starts (test.py:46-47), the policy samples like in training (the reference builds ExplorationPolicy(args, test=False),
test.py:55), each environment first plays 0..noops no-op next() calls (test.py:76-80), and the loop runs until every
environment has finished one episode, summing rewards (test.py:86-109), then prints mean / min / max / std.

Two ways to run the macro steps, same results (tests/test_gpu_eval.py):
* batched (default): one `DevicePool` macro step for all test environments.  The pool starts the next episode of an
  environment whose game ended, where the reference keeps calling next() on the finished ALE (reward 0 from then
  on, test.py:98-109), so rewards of finished environments are masked out instead;
* per environment (`-gn`, or per_env=True): the reference's loop over `AtariEmulator.next` with `Action` objects;
  every grabbed frame goes to `on_new_frame` (atari_emulator.py:60-62), which is what fills the GIFs."""
import argparse
import os
import random
import time

import numpy as np
import torch

from . import checkpoints, logger_utils
from .exploration_policy import Action, ExplorationPolicy


def get_save_frame(name, fps=30):
    """test.py:12-19 with PIL instead of imageio (absent here): returns on_new_frame; .close() writes <name>.gif."""
    from PIL import Image
    frames = []

    def get_frame(frame):
        frames.append(Image.fromarray(np.asarray(frame, np.uint8)))

    def close():
        if frames:
            frames[0].save(name + ".gif", save_all=True, append_images=frames[1:], duration=int(1000 / fps), loop=0)

    get_frame.close, get_frame.frames = close, frames
    return get_frame


def update_memory(memory, states):
    """test.py:21-24 on a CUDA tensor (N,5,84,84,4D)."""
    memory[:, :-1] = memory[:, 1:].clone()
    memory[:, -1] = states
    return memory


def get_arg_parser():
    p = argparse.ArgumentParser(description="evaluate a run folder (args.json + checkpoints/)")
    p.add_argument("-f", "--folder", type=str, dest="folder", required=True, help="run folder of the training")
    p.add_argument("-tc", "--test_count", default=1, type=int, dest="test_count", help="episodes (= environments)")
    p.add_argument("-np", "--noops", default=30, type=int, dest="noops", help="maximum no-op next() calls at the start")
    p.add_argument("-gn", "--gif_name", default=None, type=str, dest="gif_name", help="write one GIF per environment")
    p.add_argument("-gf", "--gif_folder", default="", type=str, dest="gif_folder", help="where the GIFs go")
    p.add_argument("-d", "--device", default="/gpu:0", type=str, dest="device", help="'/gpu:k'")
    return p


def prepare_args(cli):
    """test.py:36-53: overlay args.json, then force the evaluation settings."""
    args = argparse.Namespace(**vars(cli))
    for k, v in logger_utils.namespace_from(cli.folder).__dict__.items():
        if k != "device":
            setattr(args, k, v)
    args.max_global_steps = 0
    args.debugging_folder = "/tmp/logs"
    args.random_start = False
    args.single_life_episodes = False
    if args.gif_name:
        args.visualize = 1
    args.actor_id = 0
    if getattr(cli, "random_seed", None) is None:
        args.random_seed = int(np.random.RandomState(int(time.time())).randint(1000))
    else:
        args.random_seed = cli.random_seed
    return args


@torch.no_grad()
def evaluate(args, network=None, per_env=None, noop_counts=None, max_macro_steps=None, explo_policy=None):
    """Returns the float32 array of episode rewards, one per test environment (test.py:56-109)."""
    from .train import get_network_and_environment_creator
    explo_policy = explo_policy or ExplorationPolicy(args, test=False, seed=getattr(args, "seed", 0))
    seed = args.random_seed
    network_creator, env_creator = get_network_and_environment_creator(args, explo_policy, random_seed=seed)
    n = args.test_count
    environments = [env_creator.create_environment(i) for i in range(n)]
    hooks = []
    if args.gif_name:
        for i, environment in enumerate(environments):
            environment.on_new_frame = get_save_frame(os.path.join(args.gif_folder, args.gif_name + str(i)))
            hooks.append(environment.on_new_frame)
    per_env = bool(args.gif_name) if per_env is None else per_env
    for environment in environments:
        environment.get_initial_state()
    pool = environments[0].pool
    pool.set_tab_rep(explo_policy.tab_rep)
    if network is None:
        network = network_creator()
        path = checkpoints.latest_checkpoint(os.path.join(args.folder, "checkpoints"))
        if path is None:      # the reference's saver.restore fails loudly on a folder without a loadable checkpoint
            raise FileNotFoundError("no loadable checkpoint (checkpoints/-<step>.pt) under %r: refusing to evaluate a "
                                    "randomly initialised network" % (args.folder,))
        network.load_state_dict(checkpoints.load(path, pool.device))
    network = network.to(pool.device)
    if noop_counts is None:
        noop_counts = [random.randint(0, args.noops) if args.noops != 0 else 0 for _ in range(n)]
    for environment, k in zip(environments, noop_counts):              # test.py:76-80
        for _ in range(k):
            environment.next(0)
    states = pool.states
    lstm = args.arch == "LSTM"
    if lstm:
        memory = torch.zeros((n, 5) + tuple(states.shape[1:]), dtype=torch.uint8, device=pool.device)
        memory[:, -1] = states
    episodes_over = torch.zeros(n, dtype=torch.bool, device=pool.device)
    rewards = torch.zeros(n, dtype=torch.float32, device=pool.device)
    steps = 0
    while not bool(episodes_over.all()) and (max_macro_steps is None or steps < max_macro_steps):
        _, pi, rho = network(memory if lstm else states)
        a_idx, r_idx = explo_policy.choose_next_indices(pi.contiguous(), rho.contiguous(), env_creator.num_actions)
        if per_env:
            a_hot = np.eye(env_creator.num_actions)[a_idx.cpu().numpy()]
            r_hot = np.eye(explo_policy.nb_choices)[r_idx.cpu().numpy()]
            for j, environment in enumerate(environments):            # test.py:98-109
                if bool(episodes_over[j]):
                    continue                                           # a finished ALE only returns reward 0
                macro_action = Action(explo_policy.tab_rep, j, a_hot[j], r_hot[j])
                _, r, over = environment.next(macro_action.current_action)
                rewards[j] += r
                while macro_action.is_repeated() and not over:
                    _, r, over = environment.next(macro_action.repeat())
                    rewards[j] += r
                episodes_over[j] = over
                macro_action.reset()
        else:
            torch.cuda.current_stream(pool.device).synchronize()
            pool.action_idx.copy_(a_idx)
            pool.repetition_idx.copy_(r_idx)
            torch.cuda.current_stream(pool.device).synchronize()
            pool.step_async(use_indices=True)
            pool.wait()
            live = ~episodes_over
            rewards += pool.rewards * live
            episodes_over |= (pool.terminals > 0) & live
        if lstm:
            memory = update_memory(memory, states)
        steps += 1
    for h in hooks:
        h.close()
    return rewards.cpu().numpy()


def main(argv=None):
    cli = get_arg_parser().parse_args(argv)
    args = prepare_args(cli)
    rewards = evaluate(args)
    print("Performed {} tests for {}.".format(args.test_count, args.game))
    print("Mean: {0:.2f}".format(np.mean(rewards)))
    print("Min: {0:.2f}".format(np.min(rewards)))
    print("Max: {0:.2f}".format(np.max(rewards)))
    print("Std: {0:.2f}".format(np.std(rewards)))
    return rewards


if __name__ == "__main__":
    main()
