"""Fused reward clip + n-step discounted return / advantage (reference: paac.py:176,180,226-231 and
actor_learner.py:108-114) as one launch of `mn_nstep`."""
import ctypes as C

import torch

from . import _native


def nstep_returns(rewards, terminals, values, bootstrap, gamma, clip=True, stream=None):
    """rewards (raw macro-step sums), terminals (0/1), values: (T,N) float32 CUDA; bootstrap V(s_T): (N,).
    R <- bootstrap; for t = T-1..0: R = clip(r[t]) + gamma * R * (1 - terminal[t]); y[t] = R; adv[t] = R - V[t].
    Returns (y, adv), float32 (T,N)."""
    for t in (rewards, terminals, values, bootstrap):
        assert t.is_cuda and t.dtype == torch.float32
    rewards, terminals, values, bootstrap = [t.contiguous() for t in (rewards, terminals, values, bootstrap)]
    T, n = rewards.shape
    y = torch.empty_like(rewards)
    adv = torch.empty_like(rewards)
    st = torch.cuda.current_stream(rewards.device) if stream is None else stream
    with torch.cuda.device(rewards.device):
        _native.check(_native.load().mn_nstep(rewards.data_ptr(), terminals.data_ptr(), values.data_ptr(),
                                              bootstrap.data_ptr(), float(gamma), int(bool(clip)), T, n, y.data_ptr(),
                                              adv.data_ptr(), C.c_void_p(st.cuda_stream)), "mn_nstep")
    return y, adv
