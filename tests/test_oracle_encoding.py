"""Known-answer / structural tests that validate the encoding oracle itself (SURVEY.md App. B KATs i-vii).
The upstream package is absent (parity unpinned), so these are the anchors of the restatement."""
import math

import numpy as np
import pytest
import torch

from oracle import encoding_oracle as eo


def setup(D=3, L=4, T=2 ** 14, seed=0, dtype=torch.float32):
    torch.manual_seed(seed)
    scale = eo.scale_factor(np.geomspace(1.0, 1e-2, L), D)
    shift = torch.randn(L, D) * 10
    lattice = torch.randn(L, T, 2)
    pos = (torch.rand(500, D) - 0.5).to(dtype)
    return pos, lattice, scale, shift


@pytest.mark.parametrize("D", [3, 4])
def test_barycentrics_partition_of_unity(D):
    pos, lattice, scale, shift = setup(D)
    for l in range(scale.shape[0]):
        rem0, rank, bary = eo.simplex(pos, scale[l], shift[l])
        assert torch.allclose(bary.sum(1), torch.ones(pos.shape[0]), atol=1e-5)
        assert float(bary.min()) > -1e-5
        assert sorted(rank[0].tolist()) == list(range(D + 1))       # rank is a permutation
        assert bool((rem0.sum(1) == 0).all())                        # remainder-0 point lies on the hyperplane


def test_constant_table_gives_constant_times_window():
    pos, lattice, scale, shift = setup()
    lattice = torch.full_like(lattice, 0.75)
    window = eo.coarse2fine(4, 0.55)
    out = eo.encode(pos, lattice, scale, shift, window)
    exp = (0.75 * window).repeat_interleave(2)[None].expand(pos.shape[0], -1)
    assert torch.allclose(out, exp, atol=1e-5)


def test_concat_columns_and_layout():
    pos, lattice, scale, shift = setup()
    out = eo.encode(pos, lattice, scale, shift, None, True, 1e-3)
    assert out.shape == (500, (4 + 2) * 2)
    assert torch.allclose(out[:, 8:11], pos * 1e-3) and float(out[:, 11].abs().max()) == 0.0


def test_vertex_hit_returns_vertex_value():
    """a point whose elevated coordinates are exactly a lattice vertex gets that vertex's row (weight 1)"""
    D, T = 3, 2 ** 12
    scale = eo.scale_factor([1.0], D)
    shift = torch.zeros(1, D)
    lattice = torch.randn(1, T, 2)
    # elevation is linear and invertible on the hyperplane: elevated = E cf ; choose cf so that elevated = (4,0,0,-4) * k
    # from elevate(): e0 = c0+c1+c2, e1 = c1+c2 - c0, e2 = c2 - 2 c1, e3 = -3 c2
    target = torch.tensor([4.0, 0.0, 0.0, -4.0])
    c2 = -target[3] / 3; c1 = (c2 - target[2]) / 2; c0 = c1 + c2 - target[1]
    assert abs(float(c0 + c1 + c2 - target[0])) < 1e-6
    cf = torch.stack([c0, c1, c2])
    pos = (cf / scale[0])[None]
    rem0, rank, bary = eo.simplex(pos, scale[0], shift[0])
    assert torch.allclose(bary.max(1).values, torch.ones(1), atol=1e-5)
    out = eo.encode(pos, lattice, scale, shift)
    r = int(bary.argmax(1))
    idx = eo.vertex_indices(rem0, rank, T)[0, r]
    assert torch.allclose(out[0], lattice[0, idx], atol=1e-5)


def test_hash_formula_known_answer():
    # key (1,2,3): h = ((0+1)*2531011 + 2)*2531011 + 3)*2531011 mod 2^32, then mod capacity
    h = 0
    for k in (1, 2, 3):
        h = ((h + k) * 2531011) & 0xFFFFFFFF
    rem0 = torch.tensor([[1, 2, 3, -6]])
    rank = torch.tensor([[0, 0, 0, 3]])          # with r=0 no key is shifted
    assert int(eo.vertex_indices(rem0, rank, 2 ** 18)[0, 0]) == h % (2 ** 18)
    # negative keys wrap like uint32 arithmetic
    h2 = 0
    for k in (-1, -2, 3):
        h2 = ((h2 + k) * 2531011) & 0xFFFFFFFF
    assert int(eo.vertex_indices(torch.tensor([[-1, -2, 3, 0]]), rank, 1000003)[0, 0]) == h2 % 1000003


def test_continuity_across_simplex_faces():
    pos, lattice, scale, shift = setup(seed=3)
    a = pos[:50].double()
    b = a + 0.05 * torch.nn.functional.normalize(pos[50:100].double(), dim=1)      # 5 finest cells long
    ts = torch.linspace(0, 1, 4000, dtype=torch.float64)[:, None, None]
    seg = (a[None] * (1 - ts) + b[None] * ts).reshape(-1, 3)
    out = eo.encode(seg, lattice.double(), scale, shift).reshape(4000, 50, -1)
    jump = (out[1:] - out[:-1]).abs().max()
    # step = 1.25e-5 = 1/800 of the finest cell: a discontinuity at a simplex face would show up as O(1)
    assert float(jump) < 0.05, "features must be continuous along a segment (no jumps at simplex faces)"


def test_gradients_match_finite_differences_fp64():
    pos, lattice, scale, shift = setup(L=3, dtype=torch.float64, seed=5)
    lattice = lattice.double().requires_grad_(True)
    Wm = torch.randn(6, 1, dtype=torch.float64)
    p = pos[:50].clone().requires_grad_(True)

    def f(pp):
        return torch.tanh(eo.encode(pp, lattice, scale, shift) @ Wm)
    y = f(p)
    (g,) = torch.autograd.grad(y.sum(), p, create_graph=True)
    eps = 1e-7
    for j in range(3):
        dp = torch.zeros_like(p); dp[:, j] = eps
        fd = (f(p + dp) - f(p - dp)) / (2 * eps)
        ok = (fd[:, 0] - g[:, j]).abs() < 1e-4 * (1 + g[:, j].abs())      # points next to a face may cross it
        assert ok.float().mean() > 0.9
    # double backward: d/d lattice of sum(g*v) against finite differences of the analytic gradient
    v = torch.randn_like(g)
    s = (g * v).sum()
    (gl,) = torch.autograd.grad(s, lattice)
    nz = gl.abs().reshape(-1).argmax()
    d = torch.zeros_like(lattice).reshape(-1); d[nz] = 1e-6
    d = d.reshape(lattice.shape)

    def grad_at(lat):
        pp = p.detach().clone().requires_grad_(True)
        yy = torch.tanh(eo.encode(pp, lat, scale, shift) @ Wm)
        return torch.autograd.grad(yy.sum(), pp)[0]
    fd = ((grad_at(lattice.detach() + d) - grad_at(lattice.detach() - d)) * v).sum() / 2e-6
    assert abs(float(fd) - float(gl.reshape(-1)[nz])) < 1e-4 * (1 + abs(float(fd)))


def test_coarse2fine_window():
    w = eo.coarse2fine(16, 0.3)
    assert w.shape == (16,) and float(w[0]) == 1.0 and float(w[-1]) == 0.0
    assert abs(float(w[4]) - 0.5 * (1 + math.cos(math.pi * 0.8 + math.pi))) < 1e-6
    assert torch.allclose(eo.coarse2fine(8, 1.0), torch.ones(8))
    from permutohedral_encoding import Coarse2Fine
    c = Coarse2Fine(16)
    assert torch.allclose(c(0.3).cpu(), w) and c.get_last_t() == 0.3


def test_vertex_hash_is_linear_in_the_key():
    """The fused kernels hash the four vertices of a simplex from one shared h_0 (csrc/fused_common.cuh vindex3):
    h = ((k0 P + k1) P + k2) P (mod 2^32) with k_i = rem0[i] + r - 4 [rank[i] > 3 - r] equals
    rem0 . (P^3, P^2, P) + r (P^3 + P^2 + P) - 4 sum_i [rank[i] > 3 - r] P^(3-i) over Z / 2^32. Checked here on random keys."""
    import random
    rnd = random.Random(7)
    P, M = 2531011, 2 ** 32
    P1, P2, P3 = P, P * P % M, P * P * P % M
    for _ in range(20000):
        rem = [rnd.randint(-2 ** 24, 2 ** 24) for _ in range(3)]
        rank = rnd.sample(range(4), 4)
        for r in range(4):
            h = 0
            for i in range(3):
                key = rem[i] + r - (4 if rank[i] > 3 - r else 0)
                h = (h + key) % M
                h = h * P % M
            g = (rem[0] * P3 + rem[1] * P2 + rem[2] * P1 + r * (P3 + P2 + P1)) % M
            for i, q in enumerate((P3, P2, P1)):
                if rank[i] > 3 - r:
                    g = (g - 4 * q) % M
            assert g == h
