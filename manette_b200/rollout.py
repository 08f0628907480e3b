"""Rollout: the per-step bookkeeping of `PAACLearner.train` (paac.py:140-205) on device tensors.

The reference keeps, per local step t of a rollout of T = max_local_steps steps: the one-hot actions /
repetitions taken, the value estimates, the states fed to the net, the clipped rewards and the (1 - episode over)
masks; and per environment the running episode reward / length, `actions_sum`, the action x repetition histogram,
and the lists of finished episodes -- all in Python loops over the environments (paac.py:156-161,178-205).
Here one launch of `mn_rollout_record` per step does that for every environment; the buffers are torch CUDA
tensors aliasing the library's arrays, and `returns()` chains into the n-step kernel (paac.py:226-231).
"""
import ctypes as C

import torch

from . import _native
from .pool import _alias
from .returns import nstep_returns


class Rollout(object):
    def __init__(self, n_envs, max_local_steps, num_actions, tab_rep, device=None, clip=True):
        self._L = _native.load()
        if not torch.cuda.is_available():
            raise _native.NativeError("manette_b200 needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else
                                   (device.index if isinstance(device, torch.device) else int(device)))
        self.tab_rep = [int(x) for x in tab_rep]
        self.clip = bool(clip)
        h = C.c_void_p()
        ctab = (C.c_int * len(self.tab_rep))(*self.tab_rep)
        _native.check(self._L.mn_rollout_create(self.device.index, int(n_envs), int(max_local_steps), int(num_actions),
                                                len(self.tab_rep), ctab, C.byref(h)), "mn_rollout_create")
        self._h = h
        b = _native.MnRolloutBuffers()
        _native.check(self._L.mn_rollout_get_buffers(self._h, C.byref(b)), "mn_rollout_get_buffers")
        n, T, A, K, dev = b.n_envs, b.max_local_steps, b.num_actions, b.nb_choices, self.device
        self.n_envs, self.max_local_steps, self.num_actions, self.nb_choices = n, T, A, K
        self.rewards = _alias(b.rewards, (T, n), "<f4", dev, self)
        self.masks = _alias(b.masks, (T, n), "<f4", dev, self)
        self.actions = _alias(b.actions, (T, n), "<i4", dev, self)
        self.repetitions = _alias(b.repetitions, (T, n), "<i4", dev, self)
        self.episode_reward = _alias(b.episode_reward, (n,), "<f8", dev, self)
        self.episode_steps = _alias(b.episode_steps, (n,), "<i4", dev, self)
        self.actions_sum = _alias(b.actions_sum, (n, A), "<f4", dev, self)
        self.action_rep = _alias(b.action_rep, (A, K), "<i8", dev, self)
        self.stats = _alias(b.stats, (6,), "<f8", dev, self)
        self.finished_reward = _alias(b.finished_reward, (n,), "<f4", dev, self)
        self.finished_steps = _alias(b.finished_steps, (n,), "<i4", dev, self)
        self.finished_count = _alias(b.finished_count, (1,), "<i4", dev, self)
        # written by the learner (paac.py:164): value estimates of the states each step started from
        self.values = torch.zeros((T, n), dtype=torch.float32, device=dev)

    def _stream(self, stream):
        st = torch.cuda.current_stream(self.device) if stream is None else stream
        return C.c_void_p(st.cuda_stream)

    def begin(self, stream=None):
        """Start of a rollout (paac.py:142): the action x repetition histogram starts from zero."""
        _native.check(self._L.mn_rollout_begin(self._h, self._stream(stream)), "mn_rollout_begin")

    def record(self, t, rewards, terminals, action_idx, repetition_idx, stream=None):
        """After `wait_updated()` of local step t (paac.py:173-205)."""
        for x, dt in ((rewards, torch.float32), (terminals, torch.float32), (action_idx, torch.int32), (repetition_idx, torch.int32)):
            assert x.is_cuda and x.dtype == dt and x.is_contiguous() and x.numel() == self.n_envs
        with torch.cuda.device(self.device):
            _native.check(self._L.mn_rollout_record(self._h, int(t), C.c_void_p(rewards.data_ptr()), C.c_void_p(terminals.data_ptr()),
                                                    C.c_void_p(action_idx.data_ptr()), C.c_void_p(repetition_idx.data_ptr()),
                                                    int(self.clip), self._stream(stream)), "mn_rollout_record")

    def returns(self, bootstrap, gamma, stream=None):
        """paac.py:226-231 on the recorded rows: rewards are already clipped, masks = 1 - terminal."""
        return nstep_returns(self.rewards, 1.0 - self.masks, self.values, bootstrap, gamma, clip=False, stream=stream)

    def finished(self):
        """(rewards, lengths) of the episodes that ended in the last recorded step, in environment order."""
        k = int(self.finished_count.item())
        return self.finished_reward[:k].cpu().numpy(), self.finished_steps[:k].cpu().numpy()

    def close(self):
        if self._h is not None:
            self._L.mn_rollout_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
