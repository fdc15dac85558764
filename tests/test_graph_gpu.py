"""CUDA-graph replay of the training iteration (Trainer.enable_cuda_graph): static-capacity containers, device-resident pcg32
generators / iteration number / AdamW step must give the same trajectory as eager iterations with exactly-sized containers.
Curvature and off-surface terms draw from torch's Philox generator, whose offsets differ under capture, so the comparison
runs past the curvature phase with the off-surface weight at zero; everything else (samplers, jitter, importance resampling,
compositing, losses, optimizer, occupancy refresh) is driven by the pcg32 streams and is reproducible."""
import numpy as np
import pytest
import torch

import scenes

pytestmark = pytest.mark.gpu


def _fresh_rngs():
    from permuto_sdf import OccupancyGrid, RaySampler, RaySamplesPacked, VolumeRendering
    from permuto_sdf_b200.permuto_sdf import _Pcg32Host
    OccupancyGrid.m_rng, RaySampler.m_rng, VolumeRendering.m_rng = _Pcg32Host(), _Pcg32Host(), _Pcg32Host()
    RaySamplesPacked.static_capacity = False


def _run(mode, steps=8, start_iter=60000):
    from permuto_sdf import PermutoSDF
    from permuto_sdf_b200.train import HyperParams, Trainer
    _fresh_rngs()
    hp = HyperParams()
    hp.max_nr_samples_per_ray = 32
    hp.nr_samples_imp_sampling = 8
    hp.min_dist_between_samples = 1e-3
    hp.offsurface_weight = 0.0
    tr = Trainer(hp, nr_levels=8, capacity=2 ** 14, sdf_hidden=64, occupancy_resolution=128, nr_images=4, seed=3, optimizer="fused")
    tr.set_analytic_scene()
    tr.iter_nr = start_iter
    if mode == "graph":
        tr.enable_cuda_graph(warmup_steps=2)

    class Reel:
        pass
    rgb, mask, K, tf = scenes.synthetic_reel(nimg=4, H=60, W=80)
    reel = Reel()
    reel.rgb_reel, reel.mask_reel, reel.K_reel, reel.tf_world_cam_reel = [torch.from_numpy(a).cuda() for a in (rgb, mask, K, tf)]
    gen = torch.Generator().manual_seed(11)
    losses, counts = [], []
    for i in range(steps):
        pix = torch.randint(0, 60 * 80, (256,), generator=gen, dtype=torch.int32).cuda()
        img = torch.randint(0, 4, (256,), generator=gen, dtype=torch.int32).cuda()
        with torch.no_grad():
            o, d, gt, gm, idx = PermutoSDF.rays_from_reel_indices(reel, pix, img)
        losses.append(float(tr.step(o, d, gt, gm, idx)))
        counts.append(int(tr.last["nr_samples_dev"]))
    params = [p.detach().clone() for p in tr.params]
    launches = tr.graph_launches_per_step()
    tr.disable_cuda_graph()
    _fresh_rngs()
    return losses, counts, params, launches


def test_graph_replay_matches_eager_trajectory(cuda):
    l_e, n_e, p_e, _ = _run("eager")
    l_g, n_g, p_g, launches = _run("graph")
    assert n_e == n_g, "static-capacity containers must hold the same samples: %s vs %s" % (n_e, n_g)
    assert all(np.isfinite(l_g))
    for a, b in zip(l_e, l_g):
        assert abs(a - b) / abs(a) < 2e-3, (l_e, l_g)
    assert launches is not None and launches > 10
    # parameters after 8 AdamW steps: Adam moves every touched parameter by ~lr per step whatever the gradient's size, so entries
    # whose gradient is round-off noise may step in opposite directions; the trajectories must agree in norm, not entry by entry
    p_0 = _run("eager", steps=0)[2]
    num = sum(float(((a - b) ** 2).sum()) for a, b in zip(p_e, p_g)) ** 0.5
    den = sum(float(((a - b) ** 2).sum()) for a, b in zip(p_e, p_0)) ** 0.5
    assert den > 1e-2, "sanity: the parameters do move over the compared steps"
    assert num / den < 0.05, (num, den)


def test_graph_recaptures_when_schedule_branch_changes(cuda):
    """crossing a schedule boundary (end of the curvature phase at iteration 51001) invalidates the captured graph"""
    l, n, _, _ = _run("graph", steps=7, start_iter=50997)
    assert all(np.isfinite(l)) and min(n) > 1000
