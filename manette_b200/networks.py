"""Policy / value / repetition networks and the FiGAR actor-critic loss in PyTorch.

The north star keeps the small net in PyTorch; this module restates the reference's graphs so the pool's
states can be consumed where they are:
  networks.py:10-125   Operations (conv2d / fc / softmax with the "torch" uniform(+-1/sqrt(fan_in)) initialiser)
  networks.py:178-190  NIPSNetwork      networks.py:266-281 NatureNetwork      networks.py:205-225 PpwwyyxxNetwork
  networks.py:227-263  LSTMNetwork (memory (N,5,84,84,4D) -> conv stack per step -> LSTM(32) -> fc 128)
  networks.py:194-204  BayesianNetwork (NIPS + dropout(keep_percentage) + fc 256; the dropout is always on)
  policy_v_network.py:19-74  critic / actor / repetition heads and the loss
Input is the pool's uint8 NHWC tensor with channel c = d*4 + k; `permute(0,3,1,2)` of it is a channels-last view,
so nothing is copied before the first convolution.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def _torch_init(module, fan_in, weight_fan=None):
    d = 1.0 / math.sqrt(fan_in)                       # networks.py:47-50,83-85: uniform(-d, d) for weights and biases
    dw = d if weight_fan is None else 1.0 / math.sqrt(weight_fan)
    nn.init.uniform_(module.weight, -dw, dw)
    nn.init.uniform_(module.bias, -d, d)
    return module


def _conv(cin, cout, size, stride, same=False):
    m = nn.Conv2d(cin, cout, size, stride, padding=(size // 2 if same else 0) if size % 2 else 0)
    m._same_even = bool(same and size % 2 == 0)       # TF 'SAME' with an even kernel pads (1, 2): done in forward
    # The reference's conv_weight_variable reads `input_channels = shape[3]` of an [h, w, cin, cout] kernel
    # (networks.py:41-46): its weights are U(+-1/sqrt(COUT*h*w)); only the bias (networks.py:52-60) uses cin.
    # Restated as it behaves, not as it reads.
    return _torch_init(m, cin * size * size, weight_fan=cout * size * size)


def _fc(cin, cout):
    return _torch_init(nn.Linear(cin, cout), cin)


class _Act(nn.Module):
    def __init__(self, activation, alpha):
        super().__init__()
        self.leaky, self.alpha = activation == "leaky_relu", alpha

    def forward(self, x):
        return torch.maximum(x, self.alpha * x) if self.leaky else F.relu(x)


def _apply_conv(conv, x):
    if conv._same_even:                               # kernel 4, stride 1, 'SAME': pad left/top 1, right/bottom 2
        x = F.pad(x, (1, 2, 1, 2))
    return conv(x)


class PolicyVNetwork(nn.Module):
    ARCHS = ("NIPS", "NATURE", "PWYX", "LSTM", "BAYESIAN")

    def __init__(self, arch, num_actions, nb_choices, depth=1, softmax_temp=1.0, activation="relu", alpha_leaky_relu=0.1,
                 entropy_regularisation_strength=0.02, keep_percentage=0.9, amp=False):
        super().__init__()
        arch = arch.upper()
        assert arch in self.ARCHS, arch
        self.arch, self.num_actions, self.nb_choices = arch, int(num_actions), int(nb_choices)
        self.softmax_temp, self.beta, self.loss_scaling = float(softmax_temp), float(entropy_regularisation_strength), 5.0
        self.act = _Act(activation, alpha_leaky_relu)
        c = 4 * depth
        self.keep = float(keep_percentage)
        self.amp = bool(amp)      # bf16 autocast of the convolution stack (the reference computes in fp32: off by default)
        if arch in ("NIPS", "BAYESIAN"):
            self.convs = nn.ModuleList([_conv(c, 16, 8, 4), _conv(16, 32, 4, 2)])
            self.pool_after, flat, hidden = (), 32 * 9 * 9, 256
        elif arch == "NATURE":
            self.convs = nn.ModuleList([_conv(c, 32, 8, 4), _conv(32, 64, 4, 2), _conv(64, 64, 3, 1)])
            self.pool_after, flat, hidden = (), 64 * 7 * 7, 512
        else:                                          # PWYX, and the per-step stack of LSTM
            self.convs = nn.ModuleList([_conv(c, 32, 5, 1, True), _conv(32, 32, 5, 1, True), _conv(32, 64, 4, 1, True),
                                        _conv(64, 64, 3, 1, True)])
            self.pool_after, flat, hidden = (0, 1, 2), 64 * 10 * 10, 512
        if arch == "LSTM":
            self.n_steps, n_hidden, hidden = 5, 32, 128
            self.lstm = nn.LSTM(flat, n_hidden, batch_first=True)
            with torch.no_grad():                      # BasicLSTMCell(forget_bias=1.0)
                self.lstm.bias_ih_l0[n_hidden:2 * n_hidden].fill_(1.0)
                self.lstm.bias_hh_l0.zero_()
            self.lstm_out = nn.Linear(n_hidden, n_hidden)          # networks.py:118-121: random_normal w, b
            nn.init.normal_(self.lstm_out.weight); nn.init.normal_(self.lstm_out.bias)
            self.fc = _fc(n_hidden, hidden)
        else:
            self.fc = _fc(flat, hidden)
        if arch == "BAYESIAN":                         # networks.py:194-204: dropout(keep) on fc3, then fc4 256
            self.fc4 = _fc(hidden, 256)
        self.critic = _fc(hidden, 1)
        self.actor = _fc(hidden, self.num_actions)
        self.repetition = _fc(hidden, self.nb_choices)
        self.to(memory_format=torch.channels_last)

    def _features(self, x_u8_nhwc):
        x = x_u8_nhwc.permute(0, 3, 1, 2).float().mul_(1.0 / 255.0)          # networks.py:157
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.amp and x.is_cuda):
            for i, conv in enumerate(self.convs):
                x = self.act(_apply_conv(conv, x))
                if i in self.pool_after:
                    x = F.max_pool2d(x, 2, 2)
        return x.float().permute(0, 2, 3, 1).reshape(x.shape[0], -1)         # flatten in NHWC order like the reference

    def forward(self, states):
        """states: uint8 (N,84,84,4D), or for LSTM the memory (N,5,84,84,4D) oldest -> newest.
        Returns (value (N,), pi (N,A), rho (N,K)) -- probabilities, as ExplorationPolicy expects."""
        if self.arch == "LSTM":
            n = states.shape[0]
            f = self._features(states.reshape((n * self.n_steps,) + tuple(states.shape[2:]))).reshape(n, self.n_steps, -1)
            out, _ = self.lstm(f)
            h = self.act(self.fc(self.lstm_out(out[:, -1])))
        else:
            h = self.act(self.fc(self._features(states)))
        if self.arch == "BAYESIAN":                    # tf.nn.dropout sits in the graph: active when acting too
            h = self.act(self.fc4(F.dropout(h, 1.0 - self.keep, training=True)))
        v = self.critic(h).reshape(-1)
        pi = F.softmax(self.actor(h) / self.softmax_temp, dim=1)
        rho = F.softmax(self.repetition(h) / self.softmax_temp, dim=1)
        return v, pi, rho

    def loss(self, states, action_idx, repetition_idx, y, adv):
        """policy_v_network.py:24-74 with index targets instead of one-hots (the same sums).  Returns (loss, parts)."""
        v, pi, rho = self.forward(states)
        critic_loss_mean = (0.25 * (y - v) ** 2).mean()
        log_pi, log_rho = torch.log(pi + 1e-30), torch.log(rho + 1e-30)
        ent = -(pi * log_pi).sum(1) - (rho * log_rho).sum(1)
        log_sel = log_pi.gather(1, action_idx.long().unsqueeze(1)).squeeze(1) + \
            log_rho.gather(1, repetition_idx.long().unsqueeze(1)).squeeze(1)
        actor_objective_mean = (-(log_sel * adv + self.beta * ent)).mean()
        loss = self.loss_scaling * (actor_objective_mean + critic_loss_mean)
        return loss, {"critic_loss_mean": critic_loss_mean.detach(), "actor_objective_mean": actor_objective_mean.detach()}


class TFRMSProp(torch.optim.Optimizer):
    """tf.train.RMSPropOptimizer(lr, decay, epsilon) (actor_learner.py:47-48): the mean square starts at ONE and epsilon
    sits inside the square root -- with the reference's epsilon = 0.1 that is not what torch.optim.RMSprop computes."""

    def __init__(self, params, lr, decay=0.99, epsilon=0.1):
        super().__init__(params, dict(lr=lr, decay=decay, epsilon=epsilon))

    @torch.no_grad()
    def step(self):
        for g in self.param_groups:
            ps = [p for p in g["params"] if p.grad is not None]
            if not ps:
                continue
            ms = []
            for p in ps:
                st = self.state[p]
                if "ms" not in st:
                    st["ms"] = torch.ones_like(p)
                ms.append(st["ms"])
            grads = [p.grad for p in ps]
            torch._foreach_mul_(ms, g["decay"])
            torch._foreach_addcmul_(ms, grads, grads, value=1.0 - g["decay"])
            denom = torch._foreach_add(ms, g["epsilon"])
            torch._foreach_sqrt_(denom)
            torch._foreach_addcdiv_(ps, grads, denom, value=-g["lr"])
