"""GPU parity of the stand-alone kernels through the C ABI: K3 preprocess (exact), K4 FiGAR sampling (index for
index against the oracle's Philox restatement), K5 n-step returns (1e-6 relative, the north-star tolerance)."""
import numpy as np
import pytest

import util
from util import host_path

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rgb", [False, True])
@pytest.mark.parametrize("n", [1, 7, 300])
def test_preprocess_exact(rgb, n):
    import torch
    import manette_b200 as mb
    rng = np.random.RandomState(n)
    frames = (rng.randint(0, 128, size=(n, 2, 210, 160)) * 2).astype(np.uint8)
    got = mb.preprocess(torch.as_tensor(frames).cuda(), rgb=rgb).cpu().numpy()
    for e in range(n if n < 20 else 20):
        want = host_path.preprocess_indices(frames[e, 0], frames[e, 1], rgb)
        assert np.array_equal(got[e], want), (rgb, e)       # exact (the tolerance allowed is +-1)
    g, c = mb.palette()
    og, oc = host_path.palettes()
    assert np.array_equal(g, og) and np.array_equal(c, oc)


def test_preprocess_empty_batch():
    import torch
    import manette_b200 as mb
    out = mb.preprocess(torch.zeros(0, 2, 210, 160, dtype=torch.uint8, device="cuda"))
    assert out.shape == (0, 84, 84, 1)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("a,k", [(4, 11), (18, 6), (6, 1)])
def test_sample_figar_matches_oracle(mode, a, k):
    import torch
    import manette_b200 as mb
    n = 5000
    rng = np.random.RandomState(a * 100 + k)
    pi = rng.dirichlet(np.ones(a), size=n).astype(np.float32)
    rho = rng.dirichlet(np.ones(k), size=n).astype(np.float32)
    for step in (0, 1, 77):
        wa, wr, wah, wrh = host_path.choose_next_actions(pi, rho, mode, seed=0x1234567890, step=step, eps=0.3)
        ga, gr, gah, grh = mb.sample_figar(torch.as_tensor(pi).cuda(), torch.as_tensor(rho).cuda(), mode=mode, epsilon=0.3,
                                           seed=0x1234567890, step=step)
        assert np.array_equal(ga.cpu().numpy(), wa) and np.array_equal(gr.cpu().numpy(), wr)
        assert np.array_equal(gah.cpu().numpy(), wah) and np.array_equal(grh.cpu().numpy(), wrh)


def test_sample_figar_degenerate_rows():
    import torch
    import manette_b200 as mb
    pi = torch.tensor([[1.0, 0.0, 0.0], [0.0, 0.0, 1.0], [0.0, 1.0, 0.0]], device="cuda")
    rho = torch.tensor([[1.0], [1.0], [1.0]], device="cuda")
    a, r, _, _ = mb.sample_figar(pi, rho, mode=0, seed=5, step=2)
    assert a.tolist() == [0, 2, 1] and r.tolist() == [0, 0, 0]


@pytest.mark.parametrize("T,n", [(5, 32), (5, 16384), (1, 3), (20, 1000)])
def test_nstep_matches_float64_reference(T, n):
    import torch
    import manette_b200 as mb
    rng = np.random.RandomState(T * n)
    r = rng.randint(-3, 8, size=(T, n)).astype(np.float32)
    term = (rng.rand(T, n) < 0.1).astype(np.float32)
    v = rng.randn(T, n).astype(np.float32)
    boot = rng.randn(n).astype(np.float32)
    for clip in (True, False):
        wy, wadv = host_path.nstep_returns(r, term, v, boot, 0.99, clip=clip)
        gy, gadv = mb.nstep_returns(*[torch.as_tensor(x).cuda() for x in (r, term, v, boot)], 0.99, clip=clip)
        gy, gadv = gy.cpu().numpy().astype(np.float64), gadv.cpu().numpy().astype(np.float64)
        assert np.all(np.abs(gy - wy) <= 1e-6 * np.maximum(1.0, np.abs(wy)))       # 1e-6 relative fp32
        assert np.all(np.abs(gadv - wadv) <= 1e-6 * np.maximum(1.0, np.abs(wadv)))
