"""`PAACLearner(network_creator, environment_creator, explo_policy, args)` with the reference's constructor, `train()`
and `cleanup()` (paac.py:16-31,86-301 on actor_learner.py:11-136), on device tensors.

What the reference's loop does in numpy and Python per environment happens in kernels here: the networks read the
pool's states where they are, FiGAR sampling is K4, a macro step of every environment is `Runners.update_environments`
(zero-copy mode), the per-step bookkeeping is K6, returns are K5; `learner.PAACLearner.train_rollout` chains them.
This class adds what surrounds the loop: environment construction through the creator (one device pool), the
checkpoint folders (checkpoints.py), learning-rate annealing on `global_step`, periodic progress lines, and -- under
torchrun -- one process per GPU with the gradient and episode-statistics all-reduce (not in the reference).
TensorBoard summaries are out of scope."""
import collections
import logging
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from . import checkpoints
from .emulator_runner import EmulatorRunner
from .learner import PAACLearner as _DeviceLearner
from .runners import Runners


class PAACLearner(object):
    def __init__(self, network_creator, environment_creator, explo_policy, args):
        # actor_learner.py:15-38
        self.checkpoint_interval = args.checkpoint_interval
        self.debugging_folder = args.debugging_folder
        self.network_checkpoint_folder = os.path.join(self.debugging_folder, "checkpoints/")
        self.optimizer_checkpoint_folder = os.path.join(self.debugging_folder, "optimizer_checkpoints/")
        self.last_saving_step = 0
        self.game, self.max_global_steps, self.max_local_steps = args.game, args.max_global_steps, args.max_local_steps
        self.emulator_counts, self.workers = args.emulator_counts, args.emulator_workers
        self.explo_policy, self.tab_rep = explo_policy, explo_policy.tab_rep
        self.lstm_bool = args.arch == "LSTM"
        self.rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        explo_policy.seed = int(getattr(explo_policy, "seed", 0)) + 0x9E3779B1 * self.rank   # ranks draw different streams
        args.history = 5 if self.lstm_bool else 0                      # paac.py:107-112 -> the pool's history ring
        args.env_id_offset = self.rank * self.emulator_counts          # ALE seeds stay random_seed * (global id + 1)
        self.emulators = [environment_creator.create_environment(i) for i in range(self.emulator_counts)]
        self.pool = self.emulators[0].pool
        self.pool.set_tab_rep(self.tab_rep)                            # sizes the repetition head's buffers (K = nb_choices)
        self.network = network_creator()
        self.core = _DeviceLearner(self.pool, arch=args.arch, gamma=args.gamma, initial_lr=args.initial_lr,
                                   lr_annealing_steps=args.lr_annealing_steps, alpha=args.alpha, e=args.e,
                                   clip_norm=args.clip_norm, clip_norm_type=args.clip_norm_type,
                                   max_local_steps=args.max_local_steps, network=self.network, explo_policy=explo_policy,
                                   on_step=self._after_step, micro_batch=int(getattr(args, "micro_batch", 16384)))
        if self.world > 1:                                             # every replica starts from rank 0's variables
            for p in self.network.parameters():
                dist.broadcast(p.data, 0)
        # finished episodes: logged on the device per local step (no host sync inside the rollout), drained once per
        # rollout; only a bounded window is kept (the reference's lists grow for the whole 80 M-step run, paac.py:143-144,
        # and only their last ten entries are ever read, paac.py:268)
        self.total_rewards, self.total_steps = collections.deque(maxlen=1000), collections.deque(maxlen=1000)
        T, n, dev = int(args.max_local_steps), self.emulator_counts, self.pool.device
        self._fin_count = torch.zeros(T, dtype=self.core.rollout.finished_count.dtype, device=dev)
        self._fin_reward = torch.zeros((T, n), dtype=self.core.rollout.finished_reward.dtype, device=dev)
        self._fin_steps = torch.zeros((T, n), dtype=self.core.rollout.finished_steps.dtype, device=dev)
        self.runners = None

    # -- actor_learner.py:102-136
    @property
    def global_step(self):
        return self.core.global_step

    def get_lr(self):
        return self.core.get_lr()

    def rescale_reward(self, reward):
        return max(-1.0, min(1.0, reward))

    def save_vars(self, force=False):
        if force or self.global_step - self.last_saving_step >= self.checkpoint_interval:
            self.last_saving_step = self.global_step
            if self.rank == 0:
                checkpoints.save(self.network_checkpoint_folder, self.last_saving_step, self.network.state_dict())
                checkpoints.save(self.optimizer_checkpoint_folder, self.last_saving_step,
                                 self.core.optimizer.state_dict(), max_to_keep=1)

    def init_network(self):
        os.makedirs(self.network_checkpoint_folder, exist_ok=True)
        os.makedirs(self.optimizer_checkpoint_folder, exist_ok=True)
        last_saving_step = 0
        path = checkpoints.latest_checkpoint(self.network_checkpoint_folder)
        if path is None:
            logging.info("Initializing all variables")
        else:
            logging.info("Restoring network variables from previous run")
            self.network.load_state_dict(checkpoints.load(path, self.pool.device))
            last_saving_step = checkpoints.step_of(path)
        path = checkpoints.latest_checkpoint(self.optimizer_checkpoint_folder)
        if path is not None:
            logging.info("Restoring optimizer variables from previous run")
            self.core.optimizer.load_state_dict(checkpoints.load(path, self.pool.device))
        return last_saving_step

    # -- paac.py:86-297
    def _after_step(self, t):
        ro = self.core.rollout                                         # device-to-device, stream ordered: no host sync
        self._fin_count[t].copy_(ro.finished_count.reshape(()))
        self._fin_reward[t].copy_(ro.finished_reward)
        self._fin_steps[t].copy_(ro.finished_steps)

    def _after_rollout(self):
        """The episodes that ended in the rollout, in (local step, environment) order -- the order of the reference's
        appends (paac.py:186-193).  One host synchronisation per rollout."""
        counts = self._fin_count.cpu().numpy()
        if not counts.any():
            return
        rewards, steps = self._fin_reward.cpu().numpy(), self._fin_steps.cpu().numpy()
        for t, k in enumerate(counts):
            self.total_rewards.extend(float(r) for r in rewards[t, :int(k)])
            self.total_steps.extend(int(x) for x in steps[t, :int(k)])

    def train(self):
        self.core.global_step = self.init_network()
        self.last_saving_step = self.core.global_step
        global_step_start = self.global_step
        logging.debug("Starting training at Step %d", self.global_step)
        # paac.py:98 asks every emulator for its initial state; the first call resets the whole pool in one launch
        # sequence and the states stay on the device, so the other N-1 calls (one host copy each) are not made
        self.emulators[0].get_initial_state()
        self.runners = Runners(self.tab_rep, EmulatorRunner, self.emulators, self.workers, None, host_mirror=False)
        self.runners.start()
        counter, start_time, out = 0, time.time(), None
        every = max(1, int(2048 / self.emulator_counts))
        while self.global_step < self.max_global_steps:
            loop_start_time = time.time()
            out = self.core.train_rollout()
            self._after_rollout()
            counter += 1
            if counter % every == 0:
                torch.cuda.synchronize(self.pool.device)
                now = time.time()
                last_ten = 0.0 if not self.total_rewards else float(np.mean(list(self.total_rewards)[-10:]))
                steps_per_sec = self.max_local_steps * self.emulator_counts * self.world / (now - loop_start_time)
                average_steps_per_sec = (self.global_step - global_step_start) / (now - start_time)
                if self.rank == 0:
                    logging.info("Ran {} steps, at {} steps/s ({} steps/s avg), last 10 rewards avg {}"
                                 .format(self.global_step, steps_per_sec, average_steps_per_sec, last_ten))
            self.save_vars()
        self.cleanup()
        return out

    def episode_statistics(self):
        """(count, sum reward, sum length, min, max, global steps recorded) over all ranks: the rollout's double
        precision running sums (K6), all-reduced under torchrun."""
        st = self.core.rollout.stats.clone()
        if self.world > 1:
            sums, lo, hi = st[[0, 1, 2, 5]].clone(), st[3:4].clone(), st[4:5].clone()
            dist.all_reduce(sums)
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            st = torch.stack([sums[0], sums[1], sums[2], lo[0], hi[0], sums[3]])
        return st.cpu().numpy()

    def cleanup(self):
        self.save_vars(True)
        if self.runners is not None:
            self.runners.stop()
