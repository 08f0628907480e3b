"""CPU: the PyTorch restatement of the reference's networks / loss / optimiser (manette_b200/networks.py) against
numpy restatements of policy_v_network.py:24-74 and of TensorFlow's RMSProp update (actor_learner.py:47-48)."""
import numpy as np
import pytest
import torch

import util  # noqa: F401
from manette_b200.networks import PolicyVNetwork, TFRMSProp


@pytest.mark.parametrize("arch,shape", [("NIPS", (6, 84, 84, 4)), ("NATURE", (3, 84, 84, 12)), ("PWYX", (2, 84, 84, 4)), ("BAYESIAN", (3, 84, 84, 4)),
                                        ("LSTM", (2, 5, 84, 84, 4))])
def test_shapes_and_loss_formula(arch, shape):
    torch.manual_seed(0)
    A, K = 9, 11
    net = PolicyVNetwork(arch, A, K, depth=shape[-1] // 4)
    x = torch.randint(0, 256, shape, dtype=torch.uint8)
    torch.manual_seed(7)                                                  # BAYESIAN: the dropout mask is part of the graph
    v, pi, rho = net(x)
    n = shape[0]
    assert v.shape == (n,) and pi.shape == (n, A) and rho.shape == (n, K)
    assert torch.allclose(pi.sum(1), torch.ones(n), atol=1e-5) and torch.allclose(rho.sum(1), torch.ones(n), atol=1e-5)
    a, r = torch.randint(0, A, (n,)), torch.randint(0, K, (n,))
    y, adv = torch.randn(n), torch.randn(n)
    torch.manual_seed(7)
    loss, _ = net.loss(x, a, r, y, adv)
    # policy_v_network.py:24-74 in numpy, with one-hot targets like the reference feeds
    V, P, R = v.detach().numpy().astype(np.float64), pi.detach().numpy().astype(np.float64), rho.detach().numpy().astype(np.float64)
    critic = np.mean(0.25 * (y.numpy() - V) ** 2)
    lp, lr = np.log(P + 1e-30), np.log(R + 1e-30)
    ent = 0.02 * (-(P * lp).sum(1)) + 0.02 * (-(R * lr).sum(1))
    sel = (lp * np.eye(A)[a.numpy()]).sum(1) + (lr * np.eye(K)[r.numpy()]).sum(1)
    actor = np.mean(-1.0 * (sel * adv.numpy() + ent))
    assert abs(float(loss.detach()) - 5.0 * (actor + critic)) < 1e-4 * max(1.0, abs(float(loss.detach())))


def test_initialiser_bounds_follow_the_reference():
    net = PolicyVNetwork("NIPS", 4, 1)
    w = net.convs[0].weight
    assert float(w.detach().abs().max()) <= 1.0 / np.sqrt(4 * 8 * 8) + 1e-7        # networks.py:47-50
    assert float(net.fc.weight.detach().abs().max()) <= 1.0 / np.sqrt(32 * 9 * 9) + 1e-7


def test_rmsprop_is_tensorflows():
    torch.manual_seed(1)
    p = torch.nn.Parameter(torch.randn(7))
    opt = TFRMSProp([p], lr=0.0224, decay=0.99, epsilon=0.1)
    w, ms = p.detach().numpy().astype(np.float64).copy(), np.ones(7)
    for i in range(4):
        g = torch.randn(7)
        p.grad = g.clone()
        opt.step()
        gn = g.numpy().astype(np.float64)
        ms = 0.99 * ms + 0.01 * gn * gn                                   # rms slot starts at one
        w = w - 0.0224 * gn / np.sqrt(ms + 0.1)                           # epsilon inside the square root
        assert np.allclose(p.detach().numpy(), w, rtol=1e-5, atol=1e-6), i
