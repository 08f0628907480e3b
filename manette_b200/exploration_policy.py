"""FiGAR action + repetition sampling on the device (reference: exploration_policy.py:5-116).

`ExplorationPolicy(args, test=False)` keeps the reference's constructor and `get_tab_repetitions`,
`choose_next_actions(pi, rho, num_actions)`; the draw itself is one launch of `mn_sample_figar`
(Philox-4x32-10, counter = (env, step), key = seed) instead of two Python list comprehensions over N.
The reference's draws come from numpy's global unseeded RNG, so only the distribution -- not the
individual draws -- can match; the CPU oracle restates the same counter-based generator."""
import ctypes as C

import numpy as np
import torch

from . import _native
from .pool import tab_repetitions

MODE_MULTINOMIAL, MODE_EGREEDY, MODE_ARGMAX = 0, 1, 2


class Action(object):
    """FiGAR macro-action countdown (exploration_policy.py:5-36), kept for test.py-style callers."""

    def __init__(self, tab_rep, i, a, r):
        self.tab_rep = tab_rep
        self.id = i
        self.repeated = False
        self.current_action = int(np.argmax(a))
        self.nb_repetitions_left = tab_rep[int(np.argmax(r))]
        if self.nb_repetitions_left > 0:
            self.repeated = True

    def repeat(self):
        self.nb_repetitions_left -= 1
        if self.nb_repetitions_left == 0:
            self.repeated = False
        return self.current_action

    def reset(self):
        self.repeated = False
        self.current_action = 0
        self.nb_repetitions_left = 0

    def is_repeated(self):
        return self.repeated


def sample_figar(pi, rho, mode=MODE_MULTINOMIAL, epsilon=0.0, seed=0, step=0, onehot=True, stream=None):
    """pi (N,A), rho (N,K): float32 CUDA tensors.  Returns (action_idx, rep_idx, action_onehot, rep_onehot)."""
    assert pi.is_cuda and rho.is_cuda and pi.dtype == torch.float32 and rho.dtype == torch.float32
    pi, rho = pi.contiguous(), rho.contiguous()
    n, a = pi.shape
    k = rho.shape[1]
    a_idx = torch.empty(n, dtype=torch.int32, device=pi.device)
    r_idx = torch.empty(n, dtype=torch.int32, device=pi.device)
    a_hot = torch.empty(n, a, dtype=torch.float32, device=pi.device) if onehot else None
    r_hot = torch.empty(n, k, dtype=torch.float32, device=pi.device) if onehot else None
    st = torch.cuda.current_stream(pi.device) if stream is None else stream
    with torch.cuda.device(pi.device):
        _native.check(_native.load().mn_sample_figar(
            pi.data_ptr(), rho.data_ptr(), n, a, k, int(mode), float(epsilon), int(seed) & 0xFFFFFFFFFFFFFFFF,
            int(step) & 0xFFFFFFFF, a_idx.data_ptr(), r_idx.data_ptr(), a_hot.data_ptr() if onehot else None,
            r_hot.data_ptr() if onehot else None, C.c_void_p(st.cuda_stream)), "mn_sample_figar")
    return a_idx, r_idx, a_hot, r_hot


class ExplorationPolicy(object):
    def __init__(self, args, test=False, seed=0):
        self.test = test
        self.global_step = 0
        self.egreedy_policy = args.egreedy
        self.initial_epsilon = args.epsilon
        self.epsilon = args.epsilon
        self.softmax_temp = args.softmax_temp
        self.keep_percentage = args.keep_percentage
        self.annealed = args.annealed
        self.annealing_steps = 80000000                # hard-coded in the reference too (:50); --annealed_steps is unused
        self.max_repetition = args.max_repetition
        self.nb_choices = args.nb_choices
        self.tab_rep = self.get_tab_repetitions()
        self.seed = seed
        self._calls = 0

    def get_tab_repetitions(self):
        return tab_repetitions(self.max_repetition, self.nb_choices)

    def get_epsilon(self):
        if self.global_step <= self.annealing_steps:
            return self.initial_epsilon - (self.global_step * self.initial_epsilon / self.annealing_steps)
        return 0.0

    def _mode(self):
        return MODE_ARGMAX if self.test else (MODE_EGREEDY if self.egreedy_policy else MODE_MULTINOMIAL)

    def choose_next_indices(self, pi, rho, num_actions):
        """choose_next_actions for device-resident callers: CUDA (N,A), (N,K) -> int32 index tensors (no one-hots)."""
        assert pi.shape[1] == num_actions and rho.shape[1] == self.nb_choices
        a_idx, r_idx, _, _ = sample_figar(pi, rho, self._mode(), self.epsilon, self.seed, self._calls, onehot=False)
        self._calls += 1
        self.global_step += len(pi)
        if self.annealed:
            self.epsilon = self.get_epsilon()
        return a_idx, r_idx

    def choose_next_actions(self, network_output_pi, network_output_rep, num_actions):
        """Returns (new_actions (N,A), new_repetitions (N,K)) one-hot.  CUDA tensors in -> CUDA tensors out;
        numpy in -> numpy out (float64 like np.eye in the reference)."""
        as_numpy = not torch.is_tensor(network_output_pi)
        pi = torch.as_tensor(np.asarray(network_output_pi, np.float32)).cuda() if as_numpy else network_output_pi
        rho = torch.as_tensor(np.asarray(network_output_rep, np.float32)).cuda() if as_numpy else network_output_rep
        assert pi.shape[1] == num_actions and rho.shape[1] == self.nb_choices
        _, _, a_hot, r_hot = sample_figar(pi, rho, self._mode(), self.epsilon, self.seed, self._calls)
        self._calls += 1
        self.global_step += len(pi)
        if self.annealed:
            self.epsilon = self.get_epsilon()    # the reference's unqualified get_epsilon() raises NameError here
        if as_numpy:
            return a_hot.cpu().numpy().astype(np.float64), r_hot.cpu().numpy().astype(np.float64)
        return a_hot, r_hot
