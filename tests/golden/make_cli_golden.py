"""Generates tests/golden/cli_golden.json from the reference tree (run in the build container only):
  * every parser.add_argument(...) of the reference's train.py / test.py, read with `ast` (the modules cannot be
    imported: TensorFlow is absent) -- flags, dest, default, type, action;
  * the args.json files the reference ships under pretrained/ for its Atari games (the on-disk format users have).
    python tests/golden/make_cli_golden.py [/root/reference]"""
import ast
import json
import os
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cli_golden.json")


def parser_spec(path):
    spec = []
    for node in ast.walk(ast.parse(open(path).read())):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr == "add_argument":
            kw = {}
            for k in node.keywords:
                if k.arg == "type":
                    kw["type"] = k.value.id
                elif k.arg in ("default", "dest", "action", "required"):
                    kw[k.arg] = ast.literal_eval(k.value)
            spec.append({"flags": [ast.literal_eval(a) for a in node.args], **kw})
    return spec


golden = {"train": parser_spec(os.path.join(REF, "train.py")), "test": parser_spec(os.path.join(REF, "test.py")),
          "pretrained": {}}
for game in ("breakout", "ms_pacman", "pong", "seaquest", "space_invaders"):
    with open(os.path.join(REF, "pretrained", game, "args.json")) as fh:
        golden["pretrained"][game] = json.load(fh)
with open(OUT, "w") as fh:
    json.dump(golden, fh, indent=1, sort_keys=True)
print(OUT, len(golden["train"]), len(golden["test"]))
