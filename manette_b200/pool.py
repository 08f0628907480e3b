"""DevicePool: the Python face of one `mn_handle` (include/manette_b200.h).

Owns the device-resident emulator state of N environments and exposes the reference's five shared
arrays (paac.py:97-102) as torch CUDA tensors that alias the library's buffers zero-copy."""
import ctypes as C
import os

import numpy as np
import torch

from . import _native

IMG = 84
STACK = 4


class _DevArray(object):
    """Minimal __cuda_array_interface__ carrier so torch.as_tensor can alias a raw device pointer."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}
        self._owner = owner


def _alias(ptr, shape, typestr, device, owner):
    t = torch.as_tensor(_DevArray(ptr, shape, typestr, owner), device=device)
    t._mn_owner = owner   # keep the pool alive as long as a view exists
    return t


def tab_repetitions(max_repetition, nb_choices):
    """ExplorationPolicy.get_tab_repetitions (exploration_policy.py:56-62)."""
    res = [0] * nb_choices
    res[-1] = max_repetition
    if nb_choices > 2:
        for i in range(1, nb_choices - 1):
            res[i] = int(max_repetition / (nb_choices - 1)) * i
    return res


class DevicePool(object):
    """games: list of (name, rom_bytes, n_envs) groups laid out back to back (env ids 0..N-1)."""

    def __init__(self, games, rgb=False, single_life_episodes=False, random_start=False, random_seed=3,
                 env_id_offset=0, tab_rep=None, device=None, envs_per_warp=0, draw_all_frames=False, reset_memo=True, history=0):
        if not torch.cuda.is_available():
            raise _native.NativeError("manette_b200 needs a CUDA device (there is no CPU fallback)")
        self._L = _native.load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else
                                   (device.index if isinstance(device, torch.device) else int(device)))
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.zeros(1, device=self.device)          # make sure the primary context exists
        self._games = [(str(n), bytes(r), int(k)) for (n, r, k) in games]
        arr = (_native.MnGame * len(self._games))()
        for i, (name, rom, k) in enumerate(self._games):
            arr[i].name = name.encode()
            arr[i].rom = rom
            arr[i].rom_size = len(rom)
            arr[i].n_envs = k
        tab = list(tab_rep) if tab_rep is not None else [0]
        ctab = (C.c_int * len(tab))(*tab)
        cfg = _native.MnConfig(device=self.device.index, n_games=len(self._games), games=arr, rgb=int(bool(rgb)),
                               single_life_episodes=int(bool(single_life_episodes)), random_start=int(bool(random_start)),
                               random_seed=int(random_seed), env_id_offset=int(env_id_offset), nb_choices=len(tab),
                               tab_rep=ctab, envs_per_warp=int(envs_per_warp), draw_all_frames=int(bool(draw_all_frames)), no_reset_memo=int(not reset_memo), history=int(history))
        h = C.c_void_p()
        _native.check(self._L.mn_create(C.byref(cfg), C.byref(h)), "mn_create")
        self._h = h
        self.tab_rep = tab
        self.rgb = bool(rgb)
        self.random_seed = int(random_seed)
        self.env_id_offset = int(env_id_offset)
        self._map_buffers()
        self.stream = torch.cuda.Stream(device=self.device)

    # ------------------------------------------------------------------ buffers
    def _map_buffers(self):
        b = _native.MnBuffers()
        _native.check(self._L.mn_get_buffers(self._h, C.byref(b)), "mn_get_buffers")
        self.n_envs, self.num_actions, self.nb_choices, self.depth = b.n_envs, b.num_actions, b.nb_choices, b.depth
        n, d, dev = self.n_envs, self.depth, self.device
        self.states = _alias(b.states, (n, IMG, IMG, STACK * d), "|u1", dev, self)
        self.rewards = _alias(b.rewards, (n,), "<f4", dev, self)
        self.terminals = _alias(b.terminals, (n,), "<f4", dev, self)
        self.actions = _alias(b.actions, (n, self.num_actions), "<f4", dev, self)
        self.repetitions = _alias(b.repetitions, (n, self.nb_choices), "<f4", dev, self)
        self.action_idx = _alias(b.action_idx, (n,), "<i4", dev, self)
        self.repetition_idx = _alias(b.repetition_idx, (n,), "<i4", dev, self)
        self.next_calls = _alias(b.next_calls, (n,), "<i4", dev, self)
        self.frames = _alias(b.frames, (n, 2, 210, 160), "|u1", dev, self)
        self.ring = _alias(b.ring, (n, STACK, IMG, IMG, d), "|u1", dev, self)
        # the learner's observation history (paac.py:79-83,107-112) as a ring over its depth H: slot `history_head`
        # holds the newest state; memory[e][j] of the reference (j oldest -> newest) is slot (head + 1 + j) % H
        self.history_depth = int(b.history_depth)
        self.history = _alias(b.history, (n, self.history_depth, IMG, IMG, STACK * d), "|u1", dev, self) if b.history else None

    @property
    def history_head(self):
        head = C.c_int()
        _native.check(self._L.mn_history_head(self._h, C.byref(head)), "mn_history_head")
        return head.value

    def history_ordered(self, out=None, stream=None):
        """The reference's `memory` array (paac.py:107-112): (N, H, 84, 84, 4D), oldest -> newest."""
        # The gather runs on the CALLER's stream (torch's current one unless told otherwise), ordered after whatever
        # the pool's own stream still has in flight: `out` is allocated, read and freed on that same stream, so neither
        # the consumer nor the caching allocator can get ahead of the copy.
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        stream.wait_stream(self.stream)
        with torch.cuda.stream(stream):
            if out is None:
                out = torch.empty_like(self.history)
            _native.check(self._L.mn_history_gather(self._h, C.c_void_p(out.data_ptr()), self._stream_ptr(stream)), "mn_history_gather")
        return out

    def set_tab_rep(self, tab_rep):
        tab = [int(x) for x in tab_rep]
        _native.check(self._L.mn_set_tab_rep(self._h, (C.c_int * len(tab))(*tab), len(tab)), "mn_set_tab_rep")
        self.tab_rep = tab
        self._map_buffers()

    def shared_variables(self):
        """[states, rewards, terminals, actions, repetitions] in the reference's order (paac.py:97-102)."""
        return [self.states, self.rewards, self.terminals, self.actions, self.repetitions]

    # ------------------------------------------------------------------ stepping
    def _stream_ptr(self, stream):
        s = self.stream if stream is None else stream
        return C.c_void_p(s.cuda_stream)

    def reset_all(self, stream=None, wait=True):
        _native.check(self._L.mn_reset_all(self._h, self._stream_ptr(stream)), "mn_reset_all")
        if wait:
            self.wait()

    def set_host_states(self, pinned):
        """Register `pinned` -- a page-locked uint8 host tensor shaped like `states` -- as the host copy of the states: every
        later reset / macro step writes what it publishes there too, while the step runs (mn_set_host_states).  None
        unregisters.  The pool keeps a reference to the tensor for as long as it is registered."""
        if pinned is None:
            _native.check(self._L.mn_set_host_states(self._h, None), "mn_set_host_states")
            self._host_states = None
            return
        if not (pinned.is_pinned() and pinned.is_contiguous() and pinned.dtype == torch.uint8
                and tuple(pinned.shape) == tuple(self.states.shape)):
            raise ValueError("set_host_states needs a pinned, contiguous uint8 tensor of shape %s" % (tuple(self.states.shape),))
        _native.check(self._L.mn_set_host_states(self._h, C.c_void_p(pinned.data_ptr())), "mn_set_host_states")
        self._host_states = pinned

    def step_async(self, use_indices=False, stream=None):
        _native.check(self._L.mn_step_async(self._h, int(bool(use_indices)), self._stream_ptr(stream)), "mn_step_async")

    def wait(self):
        _native.check(self._L.mn_wait(self._h), "mn_wait")

    def step_host(self, actions, repetitions, states, rewards, terminals, stream=None):
        """One macro step with HOST numpy arrays (copies inside the call)."""
        for a, dt in ((actions, np.float32), (repetitions, np.float32), (states, np.uint8), (rewards, np.float32),
                      (terminals, np.float32)):
            assert a.dtype == dt and a.flags["C_CONTIGUOUS"]
        assert actions.shape == (self.n_envs, self.num_actions) and repetitions.shape == (self.n_envs, self.nb_choices)
        _native.check(self._L.mn_step_host(self._h, actions.ctypes.data, repetitions.ctypes.data, states.ctypes.data,
                                           rewards.ctypes.data, terminals.ctypes.data, self._stream_ptr(stream)),
                      "mn_step_host")

    def env_reset(self, env):
        _native.check(self._L.mn_env_reset(self._h, int(env), self._stream_ptr(None)), "mn_env_reset")

    def env_next(self, env, action_index):
        r, t = C.c_float(), C.c_int()
        _native.check(self._L.mn_env_next(self._h, int(env), int(action_index), C.byref(r), C.byref(t),
                                          self._stream_ptr(None)), "mn_env_next")
        return float(r.value), bool(t.value)

    # ------------------------------------------------------------------ taps
    def legal_actions(self, env=0):
        out = np.zeros(18, np.int32)
        k = _native.check(self._L.mn_legal_actions(self._h, int(env), out.ctypes.data), "mn_legal_actions")
        return out[:k].copy()

    def ram(self, env):
        out = np.zeros(128, np.uint8)
        _native.check(self._L.mn_get_ram(self._h, int(env), out.ctypes.data), "mn_get_ram")
        return out

    def screen(self, env):
        out = np.zeros((210, 160), np.uint8)
        _native.check(self._L.mn_get_screen(self._h, int(env), out.ctypes.data), "mn_get_screen")
        return out

    def cpu_state(self, env):
        out = np.zeros(10, np.int32)
        _native.check(self._L.mn_get_cpu_state(self._h, int(env), out.ctypes.data), "mn_get_cpu_state")
        return out

    def lives(self, env):
        lv, over, fr = C.c_int(), C.c_int(), C.c_int()
        _native.check(self._L.mn_get_lives(self._h, int(env), C.byref(lv), C.byref(over), C.byref(fr)), "mn_get_lives")
        return lv.value, bool(over.value), fr.value

    def total_next_calls(self):
        v = C.c_int64()
        _native.check(self._L.mn_total_next_calls(self._h, C.byref(v)), "mn_total_next_calls")
        return v.value

    def memo_stats(self):
        """(restored, emulated, stored) get_initial_state() calls since creation."""
        v = (C.c_int64 * 3)()
        _native.check(self._L.mn_memo_stats(self._h, v), "mn_memo_stats")
        return int(v[0]), int(v[1]), int(v[2])

    def memo_level1_hits(self):
        """Of the emulated get_initial_state() calls: those restored to the end of the reset unit from the memo's level 1
        (only the four start frames were emulated)."""
        v = C.c_int64()
        _native.check(self._L.mn_memo_level1_hits(self._h, C.byref(v)), "mn_memo_level1_hits")
        return v.value

    def total_instructions(self):
        v = C.c_int64()
        _native.check(self._L.mn_total_instructions(self._h, C.byref(v)), "mn_total_instructions")
        return v.value

    def redo_count(self):
        v = C.c_int64()
        _native.check(self._L.mn_redo_count(self._h, C.byref(v)), "mn_redo_count")
        return v.value

    def launch_count(self):
        v = C.c_int64()
        _native.check(self._L.mn_launch_count(self._h, C.byref(v)), "mn_launch_count")
        return v.value

    def profile_begin(self):
        _native.check(self._L.mn_profile_begin(self._h), "mn_profile_begin")

    def profile_end(self):
        """{kind: (summed ms, launches)} for kinds round / push / emit / other."""
        ms = (C.c_double * 4)()
        cnt = (C.c_int64 * 4)()
        _native.check(self._L.mn_profile_end(self._h, ms, cnt), "mn_profile_end")
        return {k: (ms[i], cnt[i]) for i, k in enumerate(("round", "push", "emit", "other"))}

    def close(self):
        if getattr(self, "_h", None):
            self._L.mn_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def load_rom(rom_path, game):
    with open(os.path.join(rom_path, game + ".bin"), "rb") as f:
        return f.read()


def palette():
    g = np.zeros(128, np.uint8)
    c = np.zeros((128, 3), np.uint8)
    _native.check(_native.load().mn_palette(g.ctypes.data, c.ctypes.data), "mn_palette")
    return g, c


def start_noops(seed, global_env, episode):
    return _native.load().mn_start_noops(seed & 0xFFFFFFFF, global_env & 0xFFFFFFFF, episode & 0xFFFFFFFF)
