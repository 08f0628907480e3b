"""manette_b200: B200-native (sm_100a) replacement of Manette's environment hot path -- lock-step Atari 2600
emulation under PAAC with FiGAR action repetition, fused preprocessing, FiGAR sampling and n-step returns --
behind the reference's Runners / EmulatorRunner / AtariEmulator interfaces.  See DESIGN.md."""
from . import _native
from .atari_emulator import AtariEmulator, emulators_for_pool, release_pools
from .emulator_runner import EmulatorRunner
from .environment_creator import EnvironmentCreator
from .exploration_policy import Action, ExplorationPolicy, sample_figar
from .pool import DevicePool, load_rom, palette, start_noops, tab_repetitions
from .preprocess import preprocess
from .returns import nstep_returns
from .rollout import Rollout
from .runners import Runners

__all__ = ["AtariEmulator", "EmulatorRunner", "EnvironmentCreator", "Action", "ExplorationPolicy", "sample_figar",
           "DevicePool", "load_rom", "palette", "start_noops", "tab_repetitions", "preprocess", "nstep_returns", "Runners",
           "release_pools", "emulators_for_pool", "Rollout"]
