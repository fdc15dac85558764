"""Fused NeuS compositing + loss kernels (csrc/neus_fused.cu) against the modular path: the mirrored reference modules
(VolumeRenderingNeus.compute_weights / integrate on the per-op kernels + torch losses, differentiated by autograd).
Tolerance 1e-3 relative on values and gradients (north_star); in practice the two agree to ~1e-5."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import scenes

pytestmark = pytest.mark.gpu
V = 128


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def make_rsp(nr_rays, seed, compact=True):
    from permuto_sdf import OccupancyGrid, Sphere
    o, d = scenes.make_rays(nr_rays, seed=seed)
    _, occ = scenes.analytic_occupancy(V)
    grid = OccupancyGrid(V, 1.0, [0, 0, 0])
    grid.set_grid_occupancy(T(occ.astype(np.uint8)).bool())
    sph = Sphere(0.5, [0, 0, 0])
    _, te, _, tx, hit = sph.ray_intersection(T(o), T(d))
    rsp = grid.compute_samples_in_occupied_regions(T(o), T(d), te, tx, 1e-3, 64, True)
    if compact:
        rsp = rsp.compact_to_valid_samples()
    return rsp, hit


def relerr(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


@pytest.mark.parametrize("with_mask,with_bg,learn_s,cos_anneal", [(True, False, False, 0.3), (False, True, True, 1.0), (True, True, True, 0.0)])
def test_fused_neus_matches_modular(cuda, with_mask, with_bg, learn_s, cos_anneal):
    from permuto_sdf_b200.fused import neus_render_loss
    from permuto_sdf_b200.volume_rendering import VolumeRenderingNeus
    torch.manual_seed(1)
    rsp, hit = make_rsp(300, seed=3)
    N, R = rsp.samples_pos.shape[0], rsp.ray_start_end_idx.shape[0]
    assert N > 3000
    pos = rsp.samples_pos
    # a smooth synthetic field around the analytic sphere so that alpha is neither 0 nor 1 everywhere
    sdf0 = (pos.norm(dim=-1, keepdim=True) - 0.3) + 0.01 * torch.randn(N, 1, device="cuda")
    grad0 = F.normalize(pos, dim=-1) * (1.0 + 0.1 * torch.randn(N, 1, device="cuda")) + 0.05 * torch.randn(N, 3, device="cuda")
    rgb0 = torch.rand(N, 3, device="cuda")
    gt_rgb = torch.rand(R, 3, device="cuda")
    gt_mask = (torch.rand(R, 1, device="cuda") > 0.4).float() if with_mask else None
    bg0 = torch.rand(R, 3, device="cuda") if with_bg else None
    vr = VolumeRenderingNeus().cuda()
    with torch.no_grad():
        vr.deviation_network.variance.fill_(0.45)
    forced = None if learn_s else 0.4
    w_eik, w_mask = 0.04, 0.1

    def leaves():
        a = [t.clone().requires_grad_(True) for t in (sdf0, grad0, rgb0)]
        b = bg0.clone().requires_grad_(True) if with_bg else None
        return a + [b]

    # ---- modular
    sdf, grad, rgb, bg = leaves()
    weights, wsum, bgT, inv_s = vr.compute_weights(rsp, sdf, grad, cos_anneal, forced)
    pred = vr.integrate(rsp, rgb, weights)
    if with_bg:
        pred = pred + bgT.view(-1, 1) * bg
    l_rgb = ((gt_rgb - pred).abs() * hit.view(-1, 1) * 1.0).mean()
    l_eik = ((torch.linalg.norm(grad, ord=2, dim=-1) - 1.0) ** 2).mean()
    loss_m = l_rgb + w_eik * l_eik
    if with_mask:
        loss_m = loss_m + w_mask * F.binary_cross_entropy(wsum.clip(1e-3, 1.0 - 1e-3), gt_mask)
    vr.zero_grad()
    loss_m.backward()
    gm = [sdf.grad, grad.grad, rgb.grad, bg.grad if with_bg else None]
    gs_m = vr.deviation_network.variance.grad.clone() if learn_s else None

    # ---- fused
    sdf, grad, rgb, bg = leaves()
    inv_s_f = vr.deviation_network(forced)
    loss_f, pred_f, wsum_f, w_f, terms = neus_render_loss(rsp, sdf, grad, rgb, inv_s_f, cos_anneal, gt_rgb, gt_mask, hit, w_eik, w_mask, bg_rgb=bg)
    vr.zero_grad()
    loss_f.backward()
    gf = [sdf.grad.view(-1, 1), grad.grad, rgb.grad, bg.grad if with_bg else None]

    assert relerr(pred_f, pred.detach()) < 1e-4
    assert relerr(wsum_f, wsum.detach()) < 1e-4
    assert relerr(w_f, weights.detach()) < 1e-4
    assert abs(float(loss_f) - float(loss_m)) / abs(float(loss_m)) < 1e-4
    assert abs(float(terms[2]) / N - float(l_eik)) / float(l_eik) < 1e-4
    for name, a, b in zip(("sdf", "grad", "rgb", "bg"), gf, gm):
        if b is None:
            continue
        assert relerr(a, b) < 1e-3, "d loss / d %s: %g" % (name, relerr(a, b))
    if learn_s:
        gs_f = vr.deviation_network.variance.grad
        assert abs(float(gs_f) - float(gs_m)) / (abs(float(gs_m)) + 1e-12) < 1e-3


def test_fused_neus_uncompacted_and_empty_rays(cuda):
    """slot-strided containers (rays with zero samples, gaps between rays) give the same per-ray outputs"""
    from permuto_sdf_b200.fused import neus_render_loss
    torch.manual_seed(2)
    rsp, hit = make_rsp(200, seed=5, compact=True)
    N, R = rsp.samples_pos.shape[0], rsp.ray_start_end_idx.shape[0]
    n_per_ray = (rsp.ray_start_end_idx[:, 1] - rsp.ray_start_end_idx[:, 0])
    assert int((n_per_ray == 0).sum()) > 0, "scene must contain rays without samples"
    sdf = (rsp.samples_pos.norm(dim=-1) - 0.3).requires_grad_(True)
    grad = F.normalize(rsp.samples_pos, dim=-1).clone().requires_grad_(True)
    rgb = torch.rand(N, 3, device="cuda", requires_grad=True)
    gt = torch.rand(R, 3, device="cuda")
    inv_s = torch.tensor(30.0, device="cuda")
    loss, pred, wsum, w, _ = neus_render_loss(rsp, sdf, grad, rgb, inv_s, 0.5, gt, None, hit, 0.04, 0.0)
    loss.backward()
    empty = n_per_ray == 0
    assert float(pred[empty].abs().max()) == 0.0 and float(wsum[empty].abs().max()) == 0.0
    assert torch.isfinite(sdf.grad).all() and torch.isfinite(grad.grad).all() and torch.isfinite(rgb.grad).all()
    assert float(wsum.max()) <= 1.0 + 1e-5
