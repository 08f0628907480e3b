"""The on-disk run configuration of the reference: `<debugging_folder>/args.json` is one JSON object with the
parsed command line (reference logger_utils.py:10-21 writes `vars(args)`; train.py:126 and test.py:38-40 are the
writer and the reader).  Files the reference wrote (`pretrained/*/args.json`) load unchanged, and files written
here load in the reference.  Its TensorBoard / matplotlib helpers are out of scope."""
import argparse
import json
import os

ARGS_FILE = "args.json"


def load_args(path):
    """dict of the stored arguments; {} when no path is given (the reference's convention)."""
    if path is None:
        return {}
    with open(path, "r") as fh:
        stored = json.load(fh)
    if not isinstance(stored, dict):
        raise ValueError("%s does not hold a JSON object" % path)
    return stored


def save_args(args, folder, file_name=ARGS_FILE):
    """Writes the namespace (or dict) as <folder>/<file_name>, creating the folder."""
    payload = dict(args) if isinstance(args, dict) else dict(vars(args))
    os.makedirs(folder, exist_ok=True)
    with open(os.path.join(folder, file_name), "w") as fh:
        json.dump(payload, fh)


def namespace_from(folder_or_file, **overrides):
    """argparse.Namespace from a run folder (or an args.json path), parser defaults filling what an older file
    lacks -- how test.py:38-40 overlays the stored arguments on its own."""
    from .train import get_arg_parser
    path = folder_or_file if folder_or_file.endswith(".json") else os.path.join(folder_or_file, ARGS_FILE)
    ns = get_arg_parser().parse_args([])
    for k, v in load_args(path).items():
        setattr(ns, k, v)
    for k, v in overrides.items():
        setattr(ns, k, v)
    return ns if isinstance(ns, argparse.Namespace) else argparse.Namespace(**ns)
