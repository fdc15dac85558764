"""TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

CPU (PyTorch) restatement of the permutohedral-lattice hash encoding and of the small MLP / NeuS
weight formulas of the hot path. Used only as the parity oracle by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline leg; nothing under permuto_sdf_b200/ may import it.

PARITY UNPINNED for the encoding: the algorithm lives in the external package
`permutohedral_encoding` (github.com/RaduAlexandru/permutohedral_encoding, version not pinned by the
reference: README.md:40-49, no submodule, no requirements file), which is absent from /root/reference and
from this image. This file restates its published algorithm (elevation, remainder-0
rounding, rank sort, barycentric weights, key hash h=(h+key)*2531011 mod capacity, window multiply,
concat-points columns) anchored on the reference's call sites permuto_sdf_py/models/models.py:149,186,
and is validated by structural known-answer tests (tests/test_oracle_encoding.py), not by upstream vectors.

Gradients come from autograd: rem0/rank/hash are integer decisions (no gradient), the barycentric
weights are differentiable piecewise-linear functions of the position, so first and second order
derivatives (double backward) are exact inside a simplex.
"""
import math

import numpy as np
import torch


def scale_factor(sigmas, pos_dim):
    sf = np.zeros((len(sigmas), pos_dim), dtype=np.float32)
    for l, s in enumerate(sigmas):
        for i in range(pos_dim):
            v = np.float32(1.0) / np.float32(math.sqrt(float((i + 1) * (i + 2))))
            sf[l, i] = np.float32(v / np.float32(s))
    return torch.from_numpy(sf)


def _fma(a, b, c):
    """fused multiply-add for float32 tensors (exact product and sum in float64, one rounding)"""
    if c.dtype == torch.float64:
        return a * b + c
    return (a.double() * b.double() + c.double()).to(c.dtype)


def simplex(pos, scale_l, shift_l):
    """elevated coords, remainder-0 point, rank and barycentric weights for one level.
    pos [N,D]; returns rem0 [N,D+1] int64, rank [N,D+1] int64, bary [N,D+1] (differentiable wrt pos)."""
    N, D = pos.shape
    dt = pos.dtype
    cf = (pos + shift_l.to(dt)) * scale_l.to(dt)
    elev = [None] * (D + 1)
    sm = torch.zeros(N, dtype=dt)
    for i in range(D, 0, -1):
        elev[i] = _fma(torch.full((N,), -float(i), dtype=dt), cf[:, i - 1], sm)
        sm = sm + cf[:, i - 1]
    elev[0] = sm
    E = torch.stack(elev, 1)
    with torch.no_grad():
        Ed = E.detach()
        inv = torch.tensor(1.0 / (D + 1), dtype=torch.float32).to(dt)
        v = Ed * inv
        up = torch.ceil(v) * (D + 1)
        down = torch.floor(v) * (D + 1)
        rem0 = torch.where((up - Ed) < (Ed - down), up, down).to(torch.int64)
        s = rem0.sum(1) // (D + 1)
        diff = Ed - rem0.to(dt)
        rank = torch.zeros(N, D + 1, dtype=torch.int64)
        for i in range(D):
            for j in range(i + 1, D + 1):
                lt = diff[:, i] < diff[:, j]
                rank[:, i] += lt.long()
                rank[:, j] += (~lt).long()
        rank = rank + s[:, None]
        low = rank < 0
        high = rank > D
        rem0 = rem0 + low.long() * (D + 1) - high.long() * (D + 1)
        rank = rank + low.long() * (D + 1) - high.long() * (D + 1)
    delta = (E - rem0.to(dt)) * torch.tensor(1.0 / (D + 1), dtype=torch.float32).to(dt)
    bary = torch.zeros(N, D + 2, dtype=dt)
    bary = bary.scatter_add(1, D - rank, delta)
    bary = bary.scatter_add(1, D + 1 - rank, -delta)
    b0 = bary[:, 0] + (1.0 + bary[:, D + 1])
    bary = torch.cat([b0[:, None], bary[:, 1:D + 1]], 1)
    return rem0, rank, bary


def vertex_indices(rem0, rank, capacity):
    """hashed table index of each of the D+1 simplex vertices -> [N, D+1] int64"""
    N, D1 = rem0.shape
    D = D1 - 1
    out = []
    for r in range(D + 1):
        h = torch.zeros(N, dtype=torch.int64)
        for i in range(D):
            key = rem0[:, i] + r - (D + 1) * (rank[:, i] > D - r).long()
            h = (h + key) & 0xFFFFFFFF
            h = (h * 2531011) & 0xFFFFFFFF
        out.append(h % capacity)
    return torch.stack(out, 1)


def encode(pos, lattice, scale, shift, window=None, concat_points=False, points_scaling=1.0):
    """pos [N,D], lattice [L,T,F], scale [L,D], shift [L,D] or None, window [L] or None -> [N,(L+E)*F]"""
    N, D = pos.shape
    L, T, F = lattice.shape
    dt = pos.dtype
    cols = []
    for l in range(L):
        sh = shift[l] if shift is not None else torch.zeros(D)
        rem0, rank, bary = simplex(pos, scale[l], sh)
        idx = vertex_indices(rem0, rank, T)
        w = 1.0 if window is None else window[l].to(dt)
        vals = lattice[l].to(dt)[idx]                       # [N, D+1, F]
        cols.append((vals * (bary * w)[:, :, None]).sum(1))
    if concat_points:
        E = int(math.ceil(D / F))
        extra = torch.zeros(N, E * F, dtype=dt)
        extra = torch.cat([pos * points_scaling, extra[:, : E * F - D]], 1)
        cols.append(extra)
    return torch.cat(cols, 1)


def all_indices(pos, scale, shift, capacity):
    """[N, L, D+1] hashed indices (integer parity target)"""
    L = scale.shape[0]
    res = []
    for l in range(L):
        sh = shift[l] if shift is not None else torch.zeros(pos.shape[1])
        rem0, rank, _ = simplex(pos.detach(), scale[l], sh)
        res.append(vertex_indices(rem0, rank, capacity))
    return torch.stack(res, 1)


def cosine_easing_window(num_freqs, alpha):
    x = torch.clip(alpha - torch.arange(num_freqs, dtype=torch.float32), 0.0, 1.0)
    return 0.5 * (1 + torch.cos(math.pi * x + math.pi))


def coarse2fine(nr_levels, t):
    return cosine_easing_window(nr_levels, float(t) * nr_levels)


# ------------------------------------------------------------------------------------------- small MLPs
def mlp_forward(x, weights, biases, last_linear=True):
    """Linear-GELU stack exactly like torch.nn.Sequential(Linear, GELU, ..., Linear) (models.py:153-161)"""
    h = x
    n = len(weights)
    for i, (W, b) in enumerate(zip(weights, biases)):
        h = torch.nn.functional.linear(h, W, b)
        if i < n - 1 or not last_linear:
            h = torch.nn.functional.gelu(h)
    return h


def sdf_and_gradient(pos, lattice, scale, shift, window, weights, biases, concat_points=True, points_scaling=1e-3,
                     create_graph=False):
    """SDF.get_sdf_and_gradient (models.py:199-259): sdf, d sdf/d pos, geometric feature"""
    p = pos.detach().clone().requires_grad_(True)
    feat = encode(p, lattice, scale, shift, window, concat_points, points_scaling)
    y = mlp_forward(feat, weights, biases)
    sdf = y[:, 0:1]
    grad = torch.autograd.grad(sdf, p, torch.ones_like(sdf), create_graph=create_graph, retain_graph=True)[0]
    return sdf, grad, y[:, 1:]


# ------------------------------------------------------------------------------------------- NeuS weights
def neus_alpha(sdf, grads, dirs, dt, inv_s, cos_anneal_ratio):
    """VolumeRenderingNeus.compute_weights up to alpha (volume_rendering_modules.py:129-162)"""
    true_cos = (dirs * grads).sum(-1, keepdim=True)
    iter_cos = -(torch.relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_anneal_ratio) + torch.relu(-true_cos) * cos_anneal_ratio)
    nxt = sdf + iter_cos * dt.reshape(-1, 1) * 0.5
    prv = sdf - iter_cos * dt.reshape(-1, 1) * 0.5
    pc = torch.sigmoid(prv * inv_s)
    nc = torch.sigmoid(nxt * inv_s)
    return ((pc - nc + 1e-5) / (pc + 1e-5)).clip(0.0, 1.0)
