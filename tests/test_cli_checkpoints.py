"""CPU: command line, args.json and checkpoint layout against what the reference defines
(tests/golden/cli_golden.json, made by tests/golden/make_cli_golden.py from the reference's train.py / test.py /
pretrained/*/args.json)."""
import json
import os

import pytest
import torch

import util  # noqa: F401
from manette_b200 import checkpoints, logger_utils
from manette_b200 import test as mb_test
from manette_b200 import train as mb_train

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "cli_golden.json")))
TYPES = {"int": int, "float": float, "str": str}


def _by_dest(parser):
    return {a.dest: a for a in parser._actions if a.dest != "help"}


@pytest.mark.parametrize("which,make", [("train", mb_train.get_arg_parser), ("test", mb_test.get_arg_parser)])
def test_parser_has_every_reference_flag_with_its_default(which, make):
    mine = _by_dest(make())
    for spec in GOLDEN[which]:
        a = mine[spec["dest"]]
        assert sorted(a.option_strings) == sorted(spec["flags"]), spec
        if spec.get("action") == "store_true":
            assert a.default is False and a.nargs == 0
        else:
            want = spec.get("default")
            if "type" in spec and want is not None:
                want = TYPES[spec["type"]](want)           # argparse converts string defaults (test.py: -tc default '1')
                assert a.type is TYPES[spec["type"]]
            assert a.default == want and type(a.default) is type(want), spec
        assert bool(a.required) == bool(spec.get("required", False))


def test_default_namespace_equals_the_references():
    ns = vars(mb_train.get_arg_parser().parse_args([]))
    for spec in GOLDEN["train"]:
        want = False if spec.get("action") == "store_true" else spec["default"]
        assert ns[spec["dest"]] == want


@pytest.mark.parametrize("game", sorted(GOLDEN["pretrained"]))
def test_reference_args_json_round_trip(tmp_path, game):
    stored = GOLDEN["pretrained"][game]
    with open(tmp_path / "args.json", "w") as fh:
        json.dump(stored, fh)
    ns = logger_utils.namespace_from(str(tmp_path))
    for k, v in stored.items():
        assert getattr(ns, k) == v
    assert ns.seed == 0                                     # parser defaults fill what the stored file lacks
    out = tmp_path / "again"
    logger_utils.save_args(ns, str(out))
    again = logger_utils.load_args(str(out / "args.json"))
    assert {k: again[k] for k in stored} == stored          # what we write, the reference's load_args reads back
    # test.py:36-53 on top of it
    cli = mb_test.get_arg_parser().parse_args(["-f", str(tmp_path), "-tc", "3"])
    ev = mb_test.prepare_args(cli)
    assert ev.test_count == 3 and ev.game == stored["game"] and ev.nb_choices == stored["nb_choices"]
    assert ev.random_start is False and ev.single_life_episodes is False and ev.max_global_steps == 0
    assert ev.device == "/gpu:0" and 0 <= ev.random_seed < 1000


def test_device_names():
    assert mb_train.cuda_index("/gpu:3") == 3 and mb_train.cuda_index("/cpu:0") == int(os.environ.get("LOCAL_RANK", "0"))


def test_checkpoint_folder_layout(tmp_path):
    folder = str(tmp_path / "checkpoints") + "/"
    assert checkpoints.latest_checkpoint(folder) is None
    for step in (160, 1000160, 2000320, 3000000, 4000000, 5000000, 6000000):
        path = checkpoints.save(folder, step, {"w": torch.full((3,), float(step))})
    assert os.path.basename(path) == "-6000000.pt"          # saver.save(session, 'checkpoints/', global_step) names
    kept = sorted(f for f in os.listdir(folder) if f.endswith(".pt"))
    assert len(kept) == 5 and "-160.pt" not in kept         # tf.train.Saver keeps the last five
    index = open(os.path.join(folder, "checkpoint")).read().splitlines()
    assert index[0] == 'model_checkpoint_path: "-6000000"' and index[-1] == 'all_model_checkpoint_paths: "-6000000"'
    latest = checkpoints.latest_checkpoint(folder)
    assert checkpoints.step_of(latest) == 6000000           # networks.py:173
    assert float(checkpoints.load(latest)["w"][0]) == 6000000.0
    opt = str(tmp_path / "optimizer_checkpoints") + "/"
    for step in (10, 20):
        checkpoints.save(opt, step, {"s": torch.zeros(1)}, max_to_keep=1)
    assert sorted(f for f in os.listdir(opt) if f.endswith(".pt")) == ["-20.pt"]


def test_network_and_optimizer_state_round_trip(tmp_path):
    from manette_b200.networks import PolicyVNetwork, TFRMSProp
    torch.manual_seed(0)
    net = PolicyVNetwork("NIPS", 6, 11)
    opt = TFRMSProp(net.parameters(), 0.0224)
    x = torch.randint(0, 256, (4, 84, 84, 4), dtype=torch.uint8)
    loss, _ = net.loss(x, torch.zeros(4, dtype=torch.long), torch.zeros(4, dtype=torch.long), torch.ones(4), torch.ones(4))
    loss.backward()
    opt.step()
    checkpoints.save(str(tmp_path / "c"), 5, net.state_dict())
    checkpoints.save(str(tmp_path / "o"), 5, opt.state_dict(), max_to_keep=1)
    net2 = PolicyVNetwork("NIPS", 6, 11)
    opt2 = TFRMSProp(net2.parameters(), 0.0224)
    net2.load_state_dict(checkpoints.load(checkpoints.latest_checkpoint(str(tmp_path / "c"))))
    opt2.load_state_dict(checkpoints.load(checkpoints.latest_checkpoint(str(tmp_path / "o"))))
    for a, b in zip(net.parameters(), net2.parameters()):
        assert torch.equal(a, b)
    ms = [opt.state[p]["ms"] for p in net.parameters()]
    ms2 = [opt2.state[p]["ms"] for p in net2.parameters()]
    assert all(torch.equal(a, b) for a, b in zip(ms, ms2)) and not torch.equal(ms[0], torch.ones_like(ms[0]))
