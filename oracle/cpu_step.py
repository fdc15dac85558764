"""TEST INFRASTRUCTURE / BASELINE -- NOT PRODUCT CODE.

CPU (PyTorch-only) port of the model part of one PermutoSDF training step, timed by bench.py as the
`cpu_baseline` / `--impl reference` arm. The reference itself has no CPU path (every native op hard-codes
torch::kCUDA and its encoding is a CUDA-only external package); this is the oracle
restatement (oracle/encoding_oracle.py) of
   SDF.get_sdf_and_gradient (encoding + 4-layer GELU MLP + autograd.grad with create_graph)
   + eikonal loss + a colour-like loss on the geometric feature, then backward through the double backward
(permuto_sdf_py/models/models.py:176-259, train_permuto_sdf.py:342-363,415) on `n_rays * samples_per_ray` points.
"""
import time

import numpy as np
import torch

from . import encoding_oracle as eo


class CpuSdfStep:
    def __init__(self, nr_levels=16, capacity=2 ** 18, hidden=64, geom=32, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.L, self.T = nr_levels, capacity
        self.lattice = (torch.randn(nr_levels, capacity, 2, generator=g) * 1e-2).requires_grad_(True)
        self.scale = eo.scale_factor(np.geomspace(1.0, 1e-4, nr_levels), 3)
        self.shift = torch.randn(nr_levels, 3, generator=g) * 10
        dims = [(nr_levels + 2) * 2, hidden, hidden, hidden, 1 + geom]
        self.W = [(torch.randn(dims[i + 1], dims[i], generator=g) * (2.0 / (dims[i] + dims[i + 1])) ** 0.5).requires_grad_(True) for i in range(4)]
        self.B = [torch.zeros(dims[i + 1], requires_grad=True) for i in range(4)]
        self.window = eo.coarse2fine(nr_levels, 1.0)

    def step(self, n_rays, samples_per_ray, seed=0):
        """forward + gradient + losses + backward on n_rays*samples_per_ray points; returns the loss value"""
        g = torch.Generator().manual_seed(seed)
        o = torch.randn(n_rays, 3, generator=g)
        o = 1.2 * o / o.norm(dim=1, keepdim=True)
        d = -o / o.norm(dim=1, keepdim=True)
        z = torch.linspace(0.85, 1.0, samples_per_ray)
        pos = (o[:, None, :] + z[None, :, None] * d[:, None, :]).reshape(-1, 3)
        sdf, grad, geom = eo.sdf_and_gradient(pos, self.lattice, self.scale, self.shift, self.window, self.W, self.B, True, 1e-3,
                                              create_graph=True)
        loss = ((grad.norm(dim=-1) - 1.0) ** 2).mean() * 0.04 + sdf.abs().mean() + (geom ** 2).mean()
        params = [self.lattice] + self.W + self.B
        grads = torch.autograd.grad(loss, params)
        return float(loss), grads


def time_cpu_step(n_rays, samples_per_ray, steps, warmup, threads, nr_levels=16, hidden=64):
    torch.set_num_threads(threads)
    st = CpuSdfStep(nr_levels=nr_levels, hidden=hidden)
    for i in range(warmup):
        st.step(n_rays, samples_per_ray, i)
    ts = []
    for i in range(steps):
        t0 = time.perf_counter()
        st.step(n_rays, samples_per_ray, 100 + i)
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), ts
