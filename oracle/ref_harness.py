"""ORACLE -- TEST INFRASTRUCTURE ONLY.

Imports the reference's OWN hot-path modules, unmodified, from /root/reference
(runners.py, emulator_runner.py, atari_emulator.py, environment.py,
exploration_policy.py) on top of three shim modules:

* ``ale_python_interface``  -> oracle/shims/ale_python_interface.py (the CPU oracle)
* ``scipy.misc``            -> imresize/imsave restated with PIL (scipy removed them);
                               imresize(img,(84,84),interp='nearest') was
                               ``Image.fromarray(img).resize((84,84), NEAREST)``
* ``tensorflow``            -> empty stub (exploration_policy.py:2 imports it, never uses it
                               on this path)

Only usable where /root/reference exists (this container, not the GPU box): it is
used by tests/golden/make_golden.py to generate the committed fixtures and by the
``not gpu`` tests that are skipped when the reference tree is absent.
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("MANETTE_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "atari_emulator.py"))


def _install_shims():
    shim_dir = os.path.join(_HERE, "shims")
    if shim_dir not in sys.path:
        sys.path.insert(0, shim_dir)
    if "scipy.misc" not in sys.modules or not hasattr(sys.modules["scipy.misc"], "imresize"):
        from PIL import Image
        import scipy

        misc = types.ModuleType("scipy.misc")

        def imresize(arr, size, interp="bilinear", mode=None):
            assert interp == "nearest" and arr.dtype == np.uint8
            im = Image.fromarray(arr)
            return np.asarray(im.resize((size[1], size[0]), Image.NEAREST))

        def imsave(name, arr):
            Image.fromarray(arr).save(name)

        misc.imresize = imresize
        misc.imsave = imsave
        sys.modules["scipy.misc"] = misc
        scipy.misc = misc
    if "tensorflow" not in sys.modules:
        sys.modules["tensorflow"] = types.ModuleType("tensorflow")


def load():
    """Returns a namespace holding the reference modules."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    import atari_emulator
    import emulator_runner
    import environment
    import exploration_policy
    import runners

    return types.SimpleNamespace(atari_emulator=atari_emulator, environment=environment,
                                 emulator_runner=emulator_runner, runners=runners,
                                 exploration_policy=exploration_policy)


class Args(object):
    """The subset of train.py's argparse namespace the hot path reads
    (atari_emulator.py:20,27,33-35,42; exploration_policy.py:44-54)."""

    def __init__(self, game, rom_path, rgb=False, random_start=False, single_life_episodes=False,
                 max_repetition=10, nb_choices=11, random_seed=3):
        self.game = game
        self.rom_path = rom_path
        self.rgb = rgb
        self.random_start = random_start
        self.single_life_episodes = single_life_episodes
        self.visualize = False
        self.random_seed = random_seed
        self.egreedy = False
        self.epsilon = 0.05
        self.softmax_temp = 1.0
        self.keep_percentage = 0.9
        self.annealed = False
        self.max_repetition = max_repetition
        self.nb_choices = nb_choices
