"""Drop-in for the Atari branch of the reference's EnvironmentCreator (environment_creator.py:4-30).
Tetris and gym/PLE environments are out of scope."""
import os

from .atari_emulator import AtariEmulator
from .csrc_info import minimal_action_count


class EnvironmentCreator(object):
    def __init__(self, args):
        path = os.path.join(args.rom_path, args.game + ".bin")
        if not os.path.isfile(path):
            raise ValueError("ROM not found: %s (only Atari ROM environments are supported)" % path)
        self.num_actions = minimal_action_count(args.game)
        self.create_environment = lambda i: AtariEmulator(i, args)
