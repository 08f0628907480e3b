"""Drop-in for the reference's `AtariEmulator` (atari_emulator.py:17-136) backed by a DevicePool.

Same constructor arguments (`actor_id`, `args` with `random_seed, rom_path, game, random_start,
single_life_episodes, visualize, rgb`), same methods (`get_legal_actions`, `get_initial_state`, `next`,
`get_noop`, `on_new_frame`) and results.  The emulator objects are thin handles: all objects built
with the same settings share ONE device pool, created the first time any of them is used (so the
reference's pattern -- build N emulators, call get_initial_state() on each, hand the list to
`Runners` (actor_learner.py:37-38, paac.py:98,104) -- ends up as one pool of N environments)."""
import numpy as np

from .pool import DevicePool, load_rom

IMG_SIZE_X = 84
IMG_SIZE_Y = 84
NR_IMAGES = 4
ACTION_REPEAT = 4
MAX_START_WAIT = 30
FRAMES_IN_POOL = 2


class _PoolGroup(object):
    """Emulators created with identical settings, and the pool that serves them once materialised."""

    def __init__(self, key, args):
        self.key = key
        self.args = args
        self.members = {}          # actor_id -> AtariEmulator
        self.pool = None
        self.fresh = None          # bool per env: device holds an unconsumed get_initial_state() result
        self.used = False          # any environment of the pool has been stepped or reset
        self.closed_to_new = False  # the pool is sized; later emulators start a new group

    def materialise(self):
        if self.pool is None:
            n = max(self.members) + 1
            a = self.args
            rom = load_rom(a.rom_path, a.game)
            self.pool = DevicePool([(a.game, rom, n)], rgb=bool(getattr(a, "rgb", False)),
                                   single_life_episodes=bool(a.single_life_episodes), random_start=bool(a.random_start),
                                   random_seed=int(a.random_seed), device=getattr(a, "cuda_device", None),
                                   envs_per_warp=int(getattr(a, "envs_per_warp", 0)),
                                   env_id_offset=int(getattr(a, "env_id_offset", 0)),      # rank * emulator_counts (multi-GPU)
                                   history=int(getattr(a, "history", 0)))                  # 5 for the LSTM net (paac.py:107-112)
            self.fresh = np.zeros(n, bool)
        return self.pool


_GROUPS = {}


def _group_for(args):
    key = (args.rom_path, args.game, bool(getattr(args, "rgb", False)), bool(args.single_life_episodes),
           bool(args.random_start), int(args.random_seed), getattr(args, "cuda_device", None),
           int(getattr(args, "env_id_offset", 0)), int(getattr(args, "history", 0)))
    g = _GROUPS.get(key)
    if g is None or (g.pool is not None and g.closed_to_new):
        g = _PoolGroup(key, args)
        _GROUPS[key] = g
    return g


def emulators_for_pool(pool):
    """AtariEmulator handles 0..N-1 over an existing DevicePool (e.g. a mixed-game pool built directly)."""
    g = _PoolGroup(None, None)
    g.pool = pool
    g.fresh = np.zeros(pool.n_envs, bool)
    g.closed_to_new = True
    out = []
    for i in range(pool.n_envs):
        e = AtariEmulator.__new__(AtariEmulator)
        e.actor_id = i
        e.random_start = e.single_life_episodes = None
        e.call_on_new_frame = False
        e.global_step = 0
        e.rgb = pool.rgb
        e.depth = pool.depth
        e.screen_width, e.screen_height = 160, 210
        e._group = g
        e._legal = None
        g.members[i] = e
        out.append(e)
    return out


def release_pools():
    """Frees every device pool created through AtariEmulator objects."""
    for g in list(_GROUPS.values()):
        if g.pool is not None:
            g.pool.close()
    _GROUPS.clear()


class AtariEmulator(object):
    def __init__(self, actor_id, args):
        self.actor_id = int(actor_id)
        self.random_start = args.random_start
        self.single_life_episodes = args.single_life_episodes
        self.call_on_new_frame = getattr(args, "visualize", False)
        self.global_step = 0
        self.rgb = bool(getattr(args, "rgb", False))
        self.depth = 3 if self.rgb else 1
        self.screen_width, self.screen_height = 160, 210
        g = _group_for(args)
        if g.pool is not None:
            if self.actor_id >= g.pool.n_envs:
                # the previous pool of this configuration is already sized: start a new group
                g.closed_to_new = True
                g = _group_for(args)
        g.members[self.actor_id] = self
        self._group = g
        self._legal = None

    # -- pool access
    @property
    def pool(self):
        return self._group.materialise()

    def get_legal_actions(self):
        if self._legal is None:
            self._legal = self.pool.legal_actions(self.actor_id)
        return self._legal

    def get_noop(self):
        return [1.0, 0.0]

    def on_new_frame(self, frame):
        pass

    def _observation(self):
        return self.pool.states[self.actor_id].cpu().numpy()

    def _visualize(self, frames=2):
        """atari_emulator.py:60-62,96-99: the RGB screen of each of the FRAMES_IN_POOL grabbed frames of an
        __action_repeat goes to on_new_frame, older first."""
        if self.call_on_new_frame:
            from .pool import palette
            _, rgb = palette()
            newest = self.pool.screen(self.actor_id)
            if frames > 1:
                both = self.pool.frames[self.actor_id].cpu().numpy()
                older = both[1] if np.array_equal(both[0], newest) else both[0]
                self.on_new_frame(rgb[older >> 1])
            self.on_new_frame(rgb[newest >> 1])

    def get_initial_state(self):
        """atari_emulator.py:102-110.  The first call on a fresh pool resets EVERY environment of the pool in
        one launch sequence (what paac.py:98 asks for, one emulator at a time); the other emulators'
        first calls then just read their slice."""
        g = self._group
        pool = self.pool
        if not g.used:
            pool.reset_all()
            g.fresh[:] = True
            g.used = True
        if g.fresh[self.actor_id]:
            g.fresh[self.actor_id] = False
        else:
            pool.env_reset(self.actor_id)
        self._visualize()
        return self._observation()

    def next(self, action):
        """atari_emulator.py:112-124: (observation, reward, terminal)."""
        g = self._group
        pool = self.pool
        g.used = True
        g.fresh[self.actor_id] = False
        reward, terminal = pool.env_next(self.actor_id, int(action))
        self.global_step += 1
        self._visualize()
        return self._observation(), reward, terminal
