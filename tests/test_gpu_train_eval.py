"""GPU: the reference's command line end to end on the device pool -- train.py main(args) with checkpoints and
resume, and test.py's evaluation loop, batched against per environment (rank 4)."""
import os

import numpy as np
import pytest
import torch

import util

pytestmark = pytest.mark.gpu


def _train_args(folder, **kw):
    from manette_b200 import train
    argv = ["-g", "pong", "--rom_path", util.ROMS, "-df", str(folder) + "/", "--arch", "NIPS", "--max_repetition", "10",
            "--nb_choices", "11", "-ec", "32", "-ew", "8", "--max_global_steps", "320", "--checkpoint_interval", "160"]
    args = train.get_arg_parser().parse_args(argv)
    for k, v in kw.items():
        setattr(args, k, v)
    return args


def test_train_main_checkpoints_and_resume(tmp_path):
    import manette_b200 as mb
    from manette_b200 import checkpoints, logger_utils, train
    args = _train_args(tmp_path)
    logger_utils.save_args(args, args.debugging_folder)
    learner = train.main(args)
    assert learner.global_step == 320                               # 2 updates x 5 local steps x 32 environments
    ck = os.path.join(args.debugging_folder, "checkpoints")
    assert sorted(f for f in os.listdir(ck) if f.endswith(".pt")) == ["-160.pt", "-320.pt"]
    assert os.listdir(os.path.join(args.debugging_folder, "optimizer_checkpoints")).count("-320.pt") == 1
    assert checkpoints.step_of(checkpoints.latest_checkpoint(ck)) == 320
    st = learner.episode_statistics()
    assert st[5] == 320
    weights = [p.detach().clone() for p in learner.network.parameters()]
    mb.release_pools()
    # resume: the stored step and variables come back (networks.py:162-175), lr continues its annealing
    args2 = _train_args(tmp_path, max_global_steps=480)
    learner2 = train.main(args2)
    assert learner2.global_step == 480 and learner2.last_saving_step == 480
    assert learner2.get_lr() == pytest.approx(0.0224 * (1 - 480 / 80000000))
    assert any(not torch.equal(a, b.detach()) for a, b in zip(weights, learner2.network.parameters()))
    mb.release_pools()
    # a run folder evaluates: args.json + checkpoints/ -> test.py
    from manette_b200 import test as mb_test
    cli = mb_test.get_arg_parser().parse_args(["-f", args.debugging_folder, "-tc", "2", "-np", "3"])
    ev = mb_test.prepare_args(cli)
    rewards = mb_test.evaluate(ev, max_macro_steps=3)
    assert rewards.shape == (2,) and np.all(np.isfinite(rewards))
    mb.release_pools()


@pytest.mark.parametrize("arch", ["NIPS", "LSTM"])
def test_batched_evaluation_equals_the_per_environment_loop(tmp_path, arch):
    import manette_b200 as mb
    from manette_b200 import logger_utils, train
    from manette_b200 import test as mb_test
    from manette_b200.networks import PolicyVNetwork
    args = _train_args(tmp_path, game="breakout", arch=arch)
    logger_utils.save_args(args, args.debugging_folder)
    torch.manual_seed(3)
    net = PolicyVNetwork(arch, 4, 11)
    out = []
    for per_env in (False, True):
        cli = mb_test.get_arg_parser().parse_args(["-f", args.debugging_folder, "-tc", "3"])
        cli.random_seed = 11
        ev = mb_test.prepare_args(cli)
        out.append(mb_test.evaluate(ev, network=net, per_env=per_env, noop_counts=[0, 2, 5], max_macro_steps=8))
        pool = mb.atari_emulator._GROUPS[next(iter(mb.atari_emulator._GROUPS))].pool
        out.append(pool.states.cpu().numpy().copy())
        out.append(np.stack([pool.ram(e) for e in range(3)]))
        mb.release_pools()
    assert np.array_equal(out[0], out[3]) and np.array_equal(out[1], out[4]) and np.array_equal(out[2], out[5])


def test_gif_hook_receives_both_pooled_frames(tmp_path, monkeypatch):
    import manette_b200 as mb
    from manette_b200 import logger_utils
    from manette_b200 import test as mb_test
    args = _train_args(tmp_path, game="breakout")
    logger_utils.save_args(args, args.debugging_folder)
    cli = mb_test.get_arg_parser().parse_args(["-f", args.debugging_folder, "-tc", "1", "-np", "0", "-gn", "ep", "-gf", str(tmp_path)])
    cli.random_seed = 5
    ev = mb_test.prepare_args(cli)
    assert ev.visualize == 1
    # a folder without a checkpoint is refused (the reference's saver.restore fails loudly too) ...
    with pytest.raises(FileNotFoundError):
        mb_test.evaluate(ev, max_macro_steps=1)
    mb.release_pools()
    # ... so the run folder gets one: freshly initialised weights stored the way train.py stores them
    from manette_b200 import checkpoints
    from manette_b200.networks import PolicyVNetwork
    net = PolicyVNetwork("NIPS", len(mb.csrc_info.MINIMAL_ACTIONS["breakout"]), 11)
    checkpoints.save(os.path.join(args.debugging_folder, "checkpoints"), 0, net.state_dict())
    made, orig = [], mb_test.get_save_frame
    monkeypatch.setattr(mb_test, "get_save_frame", lambda name, fps=30: made.append(orig(name, fps)) or made[-1])
    mb_test.evaluate(ev, max_macro_steps=2)
    n = len(made[0].frames)
    assert n >= 2 + 2 * 2 and n % 2 == 0                             # get_initial_state + >= 2 next() calls, 2 frames each
    from PIL import Image
    gif = Image.open(os.path.join(str(tmp_path), "ep0.gif"))
    assert gif.size == (160, 210)
    mb.release_pools()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_training_keeps_replicas_identical(tmp_path):
    """Synchronous PAAC over NCCL: rank 0's initial variables are broadcast, gradients averaged, so the replicas stay
    bit-identical; each rank steps its own environments (global ids offset by rank x emulator_counts)."""
    import json
    import subprocess
    import sys
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(os.path.dirname(__file__), "dp_worker.py"), str(tmp_path)]
    subprocess.run(cmd, check=True, timeout=600)
    r0, r1 = (json.load(open(tmp_path / ("rank%d.json" % r))) for r in (0, 1))
    assert r0["global_step"] == r1["global_step"] == 2 * 5 * 32 * 2
    assert r0["digest"] == r1["digest"]
    assert r0["steps_recorded"] == r1["steps_recorded"] == 2 * 5 * 32 * 2      # all-reduced episode statistics
    assert (r0["offset"], r1["offset"]) == (0, 32)
    assert os.path.exists(tmp_path / "run" / "checkpoints" / "-640.pt")
