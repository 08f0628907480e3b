"""Drop-in for the reference's env-pool facade `Runners` (runners.py:7-50).

    Runners(tab_rep, EmulatorRunner, emulators, workers, variables)
    start() / stop() / get_shared_variables() / update_environments() / wait_updated()

`variables` is the reference's list [states u8 (N,84,84,4D), rewards f32 (N,), terminals f32 (N,),
actions f32 (N,A) one-hot, repetitions f32 (N,K) one-hot] (paac.py:97-102).  Instead of W forked worker
processes over RawArray shared memory, the environments live on the GPU; `update_environments()`
enqueues one FiGAR macro step of every environment (emulator_runner.py:19-42) and `wait_updated()`
blocks until it has finished.

Two ways to read the results:
* host mirror (default, what paac.py expects): get_shared_variables() returns pinned host numpy arrays;
  update_environments() uploads actions/repetitions before the step and downloads rewards/terminals after it; the
  states array is registered with the pool (mn_set_host_states), whose kernels write each environment's new state into
  it as soon as that environment has finished its repeats -- the 28 KB x N of a step cross PCIe underneath the remaining
  FiGAR rounds instead of in one copy after them;
* zero-copy: get_device_variables() returns torch CUDA tensors aliasing the pool's buffers; with
  `host_mirror=False` no host copies are made at all (the learner reads and writes the device tensors).
"""
import ctypes as C

import numpy as np
import torch

from . import _native


class Runners(object):
    def __init__(self, tab_rep, EmulatorRunner, emulators, workers, variables, host_mirror=True):
        emulators = list(emulators)
        n = len(emulators)
        if workers <= 0 or n % workers != 0:
            # np.split(emulators, workers) in the reference (runners.py:17) raises for an uneven split
            raise ValueError("array split does not result in an equal division")
        pool = emulators[0].pool
        for i, e in enumerate(emulators):
            if e.pool is not pool or e.actor_id != i:
                raise ValueError("Runners needs emulators 0..N-1 of one device pool, in order")
        if pool.n_envs != n:
            raise ValueError("the device pool holds %d environments, %d emulators were given" % (pool.n_envs, n))
        self.pool = pool
        self.workers = workers
        self.emulator_runner_cls = EmulatorRunner      # accepted for signature parity; no processes are forked
        self.tab_rep = [int(x) for x in tab_rep]
        pool.set_tab_rep(self.tab_rep)
        self.host_mirror = bool(host_mirror)
        self._started = False
        self._group = getattr(emulators[0], "_group", None)
        self.variables = None
        if self.host_mirror:
            shapes = [(tuple(pool.states.shape), torch.uint8), ((n,), torch.float32), ((n,), torch.float32),
                      ((n, pool.num_actions), torch.float32), ((n, pool.nb_choices), torch.float32)]
            self._pinned = [torch.zeros(s, dtype=d).pin_memory() for s, d in shapes]
            self.variables = [t.numpy() for t in self._pinned]
            if variables is not None:
                for dst, src in zip(self.variables, variables):
                    src = np.asarray(src)
                    if src.shape != dst.shape:
                        raise ValueError("shared variable of shape %s, expected %s" % (src.shape, dst.shape))
                    dst[...] = src
                # the learner's copies of actions / repetitions are the truth until the first step
                pool.actions.copy_(self._pinned[3])
                pool.repetitions.copy_(self._pinned[4])

    def start(self):
        self._started = True

    def stop(self):
        self._started = False
        torch.cuda.synchronize(self.pool.device)
        if self.host_mirror and getattr(self.pool, "_host_states", None) is self._pinned[0]:
            self.pool.set_host_states(None)

    def get_shared_variables(self):
        if not self.host_mirror:
            return self.get_device_variables()
        return self.variables

    def get_device_variables(self):
        return self.pool.shared_variables()

    def update_environments(self, use_indices=False):
        pool = self.pool
        if self._group is not None:
            self._group.used = True
            self._group.fresh[:] = False
        stream = pool.stream
        if self.host_mirror:
            if getattr(pool, "_host_states", None) is not self._pinned[0]:
                pool.set_host_states(self._pinned[0])
            with torch.cuda.stream(stream):
                if not use_indices:
                    pool.actions.copy_(self._pinned[3], non_blocking=True)
                    pool.repetitions.copy_(self._pinned[4], non_blocking=True)
                pool.step_async(use_indices, stream)     # publishes the states into self._pinned[0] as it goes
                self._pinned[1].copy_(pool.rewards, non_blocking=True)
                self._pinned[2].copy_(pool.terminals, non_blocking=True)
        else:
            stream.wait_stream(torch.cuda.current_stream(pool.device))
            pool.step_async(use_indices, stream)

    def wait_updated(self):
        self.pool.stream.synchronize()
        self.pool.wait()
        if not self.host_mirror:
            torch.cuda.current_stream(self.pool.device).wait_stream(self.pool.stream)
