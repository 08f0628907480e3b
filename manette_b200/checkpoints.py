"""Checkpoint folders with the reference's layout (actor_learner.py:18-19,102-106; networks.py:162-175).

    <debugging_folder>/checkpoints/-<global_step>.pt              network variables (tf.train.Saver default: keep 5)
    <debugging_folder>/optimizer_checkpoints/-<global_step>.pt    RMSProp slots (max_to_keep=1)
    <folder>/checkpoint                                           TensorFlow's text index naming the newest file

The reference saves with `saver.save(session, folder, global_step=step)` where `folder` ends in '/', so its files
are named '-<step>.*' and the step is recovered as the text after the last '-' (networks.py:173).  The same names
and the same index file are used here with torch-serialised contents (the TensorFlow weight blobs of `pretrained/`
are not in the reference tree, so there is nothing to import)."""
import os
import re

import torch

INDEX = "checkpoint"


def _index_path(folder):
    return os.path.join(folder, INDEX)


def latest_checkpoint(folder):
    """Path of the newest checkpoint named by the folder's index (tf.train.latest_checkpoint), or None."""
    try:
        with open(_index_path(folder)) as fh:
            m = re.search(r'^model_checkpoint_path:\s*"([^"]*)"', fh.read(), re.M)
    except OSError:
        return None
    if not m:
        return None
    path = os.path.join(folder, m.group(1) + ".pt")
    return path if os.path.exists(path) else None


def step_of(path):
    """networks.py:173: int(path[path.rindex('-')+1:]) on the name without its extension."""
    stem = os.path.basename(path)[:-3] if path.endswith(".pt") else os.path.basename(path)
    return int(stem[stem.rindex("-") + 1:])


def save(folder, step, payload, max_to_keep=5):
    os.makedirs(folder, exist_ok=True)
    name = "-%d" % int(step)
    tmp = os.path.join(folder, name + ".pt.tmp")
    torch.save(payload, tmp)
    os.replace(tmp, os.path.join(folder, name + ".pt"))
    kept = sorted({step_of(f) for f in os.listdir(folder) if re.fullmatch(r"-\d+\.pt", f)})
    for old in kept[:-max_to_keep] if max_to_keep else []:
        os.remove(os.path.join(folder, "-%d.pt" % old))
    kept = kept[-max_to_keep:] if max_to_keep else kept
    with open(_index_path(folder), "w") as fh:
        fh.write('model_checkpoint_path: "%s"\n' % name)
        for s in kept:
            fh.write('all_model_checkpoint_paths: "-%d"\n' % s)
    return os.path.join(folder, name + ".pt")


def load(path, map_location=None):
    return torch.load(path, map_location=map_location, weights_only=True)
