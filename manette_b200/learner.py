"""PAACLearner: one synchronous PAAC + FiGAR update on device tensors (paac.py:88-300, actor_learner.py:40-63).

    rollout of T = max_local_steps macro steps:   net -> K4 sampling -> pool macro step -> K6 bookkeeping
    then:                                          bootstrap value -> K5 returns -> loss -> grads
                                                   -> (NCCL all-reduce, mean) -> global-norm clip -> RMSProp

Nothing leaves the GPU: states are read where the pool publishes them, the rollout's state rows are device copies
(or, for the LSTM net, gathers of the pool's history ring)."""
import torch
import torch.distributed as dist

from .exploration_policy import sample_figar
from .networks import PolicyVNetwork, TFRMSProp
from .rollout import Rollout


class PAACLearner(object):
    def __init__(self, pool, arch="NIPS", gamma=0.99, initial_lr=0.0224, lr_annealing_steps=80000000, alpha=0.99, e=0.1,
                 clip_norm=3.0, clip_norm_type="global", max_local_steps=5, entropy_regularisation_strength=0.02,
                 softmax_temp=1.0, mode="multinomial", epsilon=0.05, seed=0, world_envs=None, network=None, explo_policy=None,
                 on_step=None, micro_batch=16384, amp=False):
        self.pool, self.device = pool, pool.device
        self.T, self.gamma = int(max_local_steps), float(gamma)
        self.lstm = arch.upper() == "LSTM"
        if self.lstm and pool.history is None:
            raise ValueError("the LSTM architecture needs a pool created with history=5 (paac.py:107-112)")
        if network is None:
            network = PolicyVNetwork(arch, pool.num_actions, pool.nb_choices, pool.depth, softmax_temp,
                                     entropy_regularisation_strength=entropy_regularisation_strength, amp=amp)
        self.network = network.to(self.device)
        self.explo_policy, self.on_step = explo_policy, on_step
        # frames per forward / backward chunk: activations of the PWYX stack are ~1.3 MB per frame in fp32, so pools of
        # 16,384 environments (81,920 frames per acting pass of the LSTM net, 409,600 per update) go through in slices
        self.micro_rows = max(1, int(micro_batch) // (5 if self.lstm else 1))
        self.optimizer = TFRMSProp(self.network.parameters(), initial_lr, alpha, e)
        self.initial_lr, self.lr_annealing_steps = float(initial_lr), int(lr_annealing_steps)
        self.clip_norm, self.clip_norm_type = float(clip_norm), clip_norm_type
        self.mode, self.epsilon, self.seed = {"multinomial": 0, "egreedy": 1, "argmax": 2}[mode], float(epsilon), int(seed)
        n = pool.n_envs
        self.rollout = Rollout(n, self.T, pool.num_actions, pool.tab_rep, device=self.device)
        self.states = torch.empty((self.T,) + ((n, 5) if self.lstm else (n,)) + tuple(pool.states.shape[1:]),
                                  dtype=torch.uint8, device=self.device)          # paac.py:123 / whole_memory :109
        self.global_step, self.draws = 0, 0
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.step_increment = n * self.world if world_envs is None else int(world_envs)   # global_step += 1 per env (paac.py:184)
        self._flat = None

    def get_lr(self):                                                              # paac.py get_lr: linear to zero
        if self.global_step <= self.lr_annealing_steps:
            return self.initial_lr - (self.global_step * self.initial_lr / self.lr_annealing_steps)
        return 0.0

    def _forward(self, x):
        """(v, pi, rho) of the network over all rows of x, in slices of `micro_rows`."""
        if x.shape[0] <= self.micro_rows:
            return self.network(x)
        outs = [self.network(x[lo:lo + self.micro_rows]) for lo in range(0, x.shape[0], self.micro_rows)]
        return tuple(torch.cat(o) for o in zip(*outs))

    def _net_input(self):
        return self.pool.history_ordered() if self.lstm else self.pool.states

    @torch.no_grad()
    def _act(self, t):
        pool, ro = self.pool, self.rollout
        x = self._net_input()
        v, pi, rho = self._forward(x)
        if self.explo_policy is not None:              # the reference's object decides mode / epsilon / annealing
            a_idx, r_idx = self.explo_policy.choose_next_indices(pi, rho, pool.num_actions)
        else:
            a_idx, r_idx, _, _ = sample_figar(pi.contiguous(), rho.contiguous(), mode=self.mode, epsilon=self.epsilon,
                                              seed=self.seed, step=self.draws, onehot=False)
        self.draws += 1
        self.states[t].copy_(x)
        ro.values[t].copy_(v)
        pool.action_idx.copy_(a_idx)
        pool.repetition_idx.copy_(r_idx)

    def train_rollout(self):
        """One pass of the `while` body of PAACLearner.train (paac.py:140-262).  Returns a dict of scalars (tensors)."""
        pool, ro, st = self.pool, self.rollout, torch.cuda.current_stream(self.device)
        ro.begin()
        for t in range(self.T):
            self._act(t)
            pool.stream.wait_stream(st)
            pool.step_async(use_indices=True)
            st.wait_stream(pool.stream)
            ro.record(t, pool.rewards, pool.terminals, pool.action_idx, pool.repetition_idx)
            self.global_step += self.step_increment
            if self.on_step is not None:
                self.on_step(t)
        pool.wait()
        with torch.no_grad():
            boot, _, _ = self._forward(self._net_input())                          # paac.py:217-222
        y, adv = ro.returns(boot.contiguous(), self.gamma)
        n_rows = self.T * pool.n_envs
        flat_states = self.states.reshape((n_rows,) + tuple(self.states.shape[2:]))
        self.network.zero_grad(set_to_none=False)
        fa, fr, fy, fadv = ro.actions.reshape(-1), ro.repetitions.reshape(-1), y.reshape(-1), adv.reshape(-1)
        loss, parts = None, None
        for lo in range(0, n_rows, self.micro_rows):                               # the loss is a mean over rows: slices add up
            hi = min(n_rows, lo + self.micro_rows)
            l_c, p_c = self.network.loss(flat_states[lo:hi], fa[lo:hi], fr[lo:hi], fy[lo:hi], fadv[lo:hi])
            w = (hi - lo) / n_rows
            (l_c * w).backward()
            loss = l_c.detach() * w if loss is None else loss + l_c.detach() * w
            parts = {k: v * w for k, v in p_c.items()} if parts is None else {k: parts[k] + v * w for k, v in p_c.items()}
        grads = [p.grad for p in self.network.parameters()]
        if self.world > 1:                                                         # synchronous PAAC across GPUs: mean gradient
            flat = torch._utils._flatten_dense_tensors(grads)
            dist.all_reduce(flat)
            flat.div_(self.world)
            for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
                g.copy_(f)
        global_norm = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in grads]))
        if self.clip_norm_type == "global":                                        # tf.clip_by_global_norm (actor_learner.py:57-60)
            scale = self.clip_norm / torch.maximum(global_norm, torch.as_tensor(self.clip_norm, device=self.device))
            torch._foreach_mul_(grads, scale)
        elif self.clip_norm_type == "local":          # tf.clip_by_norm per tensor (:62-65); 'ignore' leaves them alone
            for g in grads:
                g.mul_(self.clip_norm / torch.maximum(torch.linalg.vector_norm(g), torch.as_tensor(self.clip_norm, device=self.device)))
        for grp in self.optimizer.param_groups:
            grp["lr"] = self.get_lr()
        self.optimizer.step()
        return {"loss": loss, "global_norm": global_norm, "lr": self.get_lr(), **parts}

    def close(self):
        self.rollout.close()
