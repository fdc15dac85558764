"""Colour-network helper kernels (csrc/rgb_misc.cu) against the PyTorch formulas of the reference modules they replace
(LipshitzMLP.normalization, models.py:96-110; Colorcal.calib_RGB_samples_packed + sigmoid, models.py:395-414,677-741).
Tolerance 1e-5 relative: same fp32 formulas, different summation order."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-20))


@pytest.mark.parametrize("rows,cols,cval", [(128, 96, 2.0), (64, 128, 0.5), (3, 64, -1.0), (128, 128, 30.0)])
def test_lipschitz_normalization_matches_torch(cuda, rows, cols, cval):
    from permuto_sdf_b200.fused import LipschitzNormFn
    from permuto_sdf_b200.models import LipshitzMLP
    torch.manual_seed(rows + cols)
    w0 = torch.randn(rows, cols, device="cuda") * 0.2
    g = torch.randn(rows, cols, device="cuda")
    res = []
    for fused in (True, False):
        w = w0.clone().requires_grad_(True)
        c = torch.full((1,), cval, device="cuda", requires_grad=True)
        out = LipschitzNormFn.apply(w, c) if fused else LipshitzMLP.normalization(w, F.softplus(c))
        (out * g).sum().backward()
        res.append((out.detach(), w.grad, c.grad))
    scale = float((res[1][0].abs().sum(1) / w0.abs().sum(1)).min())
    assert rel(res[0][0], res[1][0]) < 1e-5
    assert rel(res[0][1], res[1][1]) < 1e-4
    assert float((res[0][2] - res[1][2]).abs()) < 1e-4 * max(1.0, float(res[1][2].abs())), (res[0][2], res[1][2])
    if cval < 1.0:
        assert scale < 1.0, "the clamp must be active for some rows in this case"


@pytest.mark.parametrize("with_calib", [True, False])
def test_calib_sigmoid_matches_torch(cuda, with_calib):
    from permuto_sdf_b200.fused import CalibSigmoidFn
    from permuto_sdf_b200.models import Colorcal
    torch.manual_seed(5)
    R, nimg = 200, 6
    counts = torch.randint(0, 40, (R,))
    counts[::7] = 0
    ends = torch.cumsum(counts, 0)
    se = torch.stack([ends - counts, ends], 1).to(torch.int32).cuda()
    N = int(ends[-1]) + 13                      # padding rows past the last ray
    x0 = torch.randn(N, 3, device="cuda")
    g = torch.randn(N, 3, device="cuda")
    img = torch.randint(0, nimg, (R,), dtype=torch.int32, device="cuda")
    cal = Colorcal(nimg, 0)
    with torch.no_grad():
        cal.weight_delta.normal_(0, 0.2); cal.bias.normal_(0, 0.2)
    res = []
    for fused in (True, False):
        x = x0.clone().requires_grad_(True)
        cal.zero_grad()
        if fused:
            out = CalibSigmoidFn.apply(x, se, img if with_calib else None, cal.weight_delta if with_calib else None,
                                       cal.bias if with_calib else None, 0 if with_calib else -1)
        else:
            y = cal.calib_RGB_samples_packed(x, img, se) if with_calib else x
            out = torch.sigmoid(y)
        valid = torch.zeros(N, 1, device="cuda")
        valid[:int(ends[-1])] = 1.0
        ((out * g) * valid).sum().backward()
        res.append((out.detach() * valid, x.grad.clone(), cal.weight_delta.grad.clone() if with_calib else None,
                    cal.bias.grad.clone() if with_calib else None))
    assert rel(res[0][0], res[1][0]) < 1e-5
    assert rel(res[0][1], res[1][1]) < 1e-4
    if with_calib:
        assert rel(res[0][2], res[1][2]) < 1e-4 and rel(res[0][3], res[1][3]) < 1e-4
        assert float(res[0][2][0].abs().max()) == 0.0, "the fixed-calibration image gets no gradient"


@pytest.mark.parametrize("static", [False, True])
def test_curvature_loss_matches_torch(cuda, static):
    """fused curvature loss (shifted points, acos of the normal agreement, masked mean; forward + backward) against the PyTorch chain of
    SDF.get_sdf_and_curvature_1d_precomputed_gradient_normal_based (models.py:261-294)"""
    import math
    from permuto_sdf_b200.fused import CurvatureLossFn, curvature_shifted_points
    torch.manual_seed(3)
    N, nv = 5000, 4321
    g0 = torch.randn(N, 3, device="cuda")
    gs0 = g0 + 0.3 * torch.randn(N, 3, device="cuda")
    gs0[:50] = g0[:50] * 2.0                                # exactly parallel normals: the clamp region
    n_dev = torch.tensor([nv], dtype=torch.int32, device="cuda") if static else None
    res = []
    for fused in (True, False):
        g, gs = g0.clone().requires_grad_(True), gs0.clone().requires_grad_(True)
        if fused:
            loss = CurvatureLossFn.apply(g, gs, n_dev)
        else:
            dot = (F.normalize(g, dim=-1) * F.normalize(gs, dim=-1)).sum(dim=-1, keepdim=True)
            curv = torch.acos(torch.clamp(dot, -1.0 + 1e-6, 1.0 - 1e-6)) / math.pi
            loss = curv[:nv].mean() if static else curv.mean()
        (loss * 3.0).backward()
        res.append((float(loss), g.grad.clone(), gs.grad.clone()))
    assert abs(res[0][0] - res[1][0]) < 1e-5 * abs(res[1][0])
    assert rel(res[0][1], res[1][1]) < 1e-3 and rel(res[0][2], res[1][2]) < 1e-3
    if static:
        assert float(res[0][1][nv:].abs().max()) == 0.0
    # shifted points: eps * unit tangent, orthogonal to the normal
    p = torch.rand(N, 3, device="cuda")
    torch.manual_seed(9)
    sh = curvature_shifted_points(p, g0, 1e-4)
    d = sh - p
    assert float((d * F.normalize(g0, dim=-1)).sum(-1).abs().max()) < 3e-7       # tangent to the normal (up to the rounding of p + d)
    assert float(d.norm(dim=-1).max()) <= 1.01e-4
