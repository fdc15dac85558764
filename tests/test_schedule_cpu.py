"""Host-side logic of the CUDA-graph mode: DeviceIter validity ranges and map_range_val on plateaus / ramps
(permuto_sdf_py/utils/common_utils.py:156-160 semantics), checked on CPU tensors."""
import math

import pytest
import torch

from permuto_sdf_b200.models import DeviceIter, map_range_val


def test_map_range_val_host():
    assert map_range_val(5, 0.0, 10.0, 0.0, 1.0) == pytest.approx(0.5)
    assert map_range_val(-3, 0.0, 10.0, 0.3, 1.0) == pytest.approx(0.3)
    assert map_range_val(50, 0.0, 10.0, 0.3, 1.0) == pytest.approx(1.0)


@pytest.mark.parametrize("it", [0, 1, 17, 9999, 10000, 10001, 34999, 35000, 50000, 50500, 51001, 70000])
def test_device_iter_matches_host_schedule(it):
    d = DeviceIter(it, torch.tensor(float(it)))
    for (a, b, c, e) in [(0.0, 10000, 0.3, 1.0), (0.0, 35000, 0.0, 1.0), (50000, 51001, 1.0, 0.0)]:
        want = map_range_val(it, a, b, c, e)
        got = map_range_val(d, a, b, c, e)
        got_f = float(got)
        assert abs(got_f - want) < 1e-6 * max(1.0, abs(want))
        # plateaus are host constants, ramps are device tensors
        assert isinstance(got, torch.Tensor) == (a < it < b)
    # every iteration inside [lo, hi) takes the same branches and the device formula gives the host value there
    assert d.lo <= it < d.hi
    for other in {max(d.lo, it - 3), it, min(d.hi - 1, it + 3)}:
        if other == float("-inf") or other == float("inf"):
            continue
        o = int(other)
        d2 = DeviceIter(o, torch.tensor(float(o)))
        for (a, b, c, e) in [(0.0, 10000, 0.3, 1.0), (0.0, 35000, 0.0, 1.0), (50000, 51001, 1.0, 0.0)]:
            g1, g2 = map_range_val(d, a, b, c, e), map_range_val(d2, a, b, c, e)
            assert isinstance(g1, torch.Tensor) == isinstance(g2, torch.Tensor), "branch changed inside the validity range"
            assert abs(float(g2) - map_range_val(o, a, b, c, e)) < 1e-6


def test_device_iter_comparisons_narrow_range():
    d = DeviceIter(100, torch.tensor(100.0))
    assert (d >= 50) and not (d >= 200) and (d < 200) and not (d < 100) and (d > 99) and (d <= 100)
    assert d.lo == 100 and d.hi == 101       # d < 100 is False -> lo = 100; d <= 100 True -> hi = 101
    with pytest.raises(TypeError):
        d == 100
    assert int(d) == 100 and math.isclose(float(d), 100.0)


def test_splitk_weight_gradient_matches_plain_gemm():
    """host logic of the split-K weight-gradient product (fused.splitk_tn / SplitKLinearFn) on CPU tensors"""
    from permuto_sdf_b200.fused import SplitKLinearFn, splitk_tn
    torch.manual_seed(0)
    for rows in (10, 2048, 4100):
        a, b = torch.randn(rows, 7, dtype=torch.float64), torch.randn(rows, 5, dtype=torch.float64)
        assert torch.allclose(splitk_tn(a, b), a.t() @ b, rtol=1e-12, atol=1e-12)
    x = torch.randn(4100, 6, dtype=torch.float64, requires_grad=True)
    w = torch.randn(3, 6, dtype=torch.float64, requires_grad=True)
    bias = torch.randn(3, dtype=torch.float64, requires_grad=True)
    y = SplitKLinearFn.apply(x, w, bias)
    y.pow(2).sum().backward()
    g = [t.grad.clone() for t in (x, w, bias)]
    for t in (x, w, bias):
        t.grad = None
    torch.nn.functional.linear(x, w, bias).pow(2).sum().backward()
    for a, t in zip(g, (x, w, bias)):
        assert torch.allclose(a, t.grad, rtol=1e-10, atol=1e-10)


def test_lr_schedule_matches_warmup_then_multistep():
    """GradualWarmupScheduler(multiplier=1, total_epoch=3000, after_scheduler=MultiStepLR(milestones, gamma=0.3)) as driven by
    train_permuto_sdf.py:417-421, replayed with torch's own MultiStepLR and the reference's warm-up rule"""
    from permuto_sdf_b200.train import HyperParams, lr_schedule
    hp = HyperParams()
    hp.lr_milestones = (40, 70)
    hp.lr_warmup_iters = 10
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=hp.lr)
    decay = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=list(hp.lr_milestones), gamma=hp.lr_decay_gamma)
    # iteration 0 steps with the base lr; the warm-up scheduler is created afterwards (lr -> 0) and stepped once per iteration
    used = [hp.lr]
    last_epoch, finished = 0, False
    for it in range(1, 120):
        last_epoch += 1
        if finished:
            decay.step()
            lr = decay.get_last_lr()[0]
        elif last_epoch > hp.lr_warmup_iters:
            finished = True
            lr = decay.get_last_lr()[0]
        else:
            lr = hp.lr * last_epoch / hp.lr_warmup_iters
        used.append(lr)
    for it, want in enumerate(used):
        assert lr_schedule(hp, it) == pytest.approx(want, rel=1e-12), it
    hp.use_lr_schedule = False
    assert lr_schedule(hp, 5) == hp.lr


def test_checkpoint_keys_match_reference_models(tmp_path):
    """state_dict key names of the reference's classes (models.py:75-79,149-161,330-340,449-470; SingleVarianceNetwork
    volume_rendering_modules.py:95) so that checkpoints written by either side load on the other; save() file layout of
    models.py:296-307"""
    import os
    from permuto_sdf_b200.models import RGB, SDF, Colorcal, NerfHash
    sdf = SDF(3, None, 32, 1000, nr_levels=4, capacity=2 ** 8)
    rgb = RGB(3, None, 32, 1, nr_levels=4, capacity=2 ** 8)
    bg = NerfHash(4, None, 1, nr_levels=4, capacity=2 ** 8)
    cal = Colorcal(3, 0)
    ks = set(sdf.state_dict())
    assert ks == {"encoding.lattice_values", "encoding.random_shift_per_level"} | {"mlp_sdf.%d.%s" % (i, n) for i in (0, 2, 4, 6) for n in ("weight", "bias")}
    kr = set(rgb.state_dict())
    want = {"volume_renderer_neus.deviation_network.variance", "encoding.lattice_values", "encoding.random_shift_per_level"}
    want |= {"mlp.layers.%d.%s" % (i, n) for i in range(4) for n in ("weight", "bias")}
    want |= {"mlp.%s.%d" % (n, i) for i in range(4) for n in ("weights_per_layer", "biases_per_layer", "lipshitz_bound_per_layer")}
    assert kr == want
    assert rgb.mlp.weights_per_layer[2] is rgb.mlp.layers[2].weight          # the same Parameter, registered twice like the reference
    assert len(list(rgb.parameters())) == 1 + 1 + 1 + 8 + 4                   # no duplicates reach the optimizer
    kb = set(bg.state_dict())
    assert {"mlp_feat_and_density.6.weight", "mlp_rgb.4.bias", "encoding.lattice_values"} <= kb
    assert set(cal.state_dict()) == {"weight_delta", "bias"}
    path = sdf.save(str(tmp_path), "exp", 300)
    rgb.save(str(tmp_path), "exp", 300); bg.save(str(tmp_path), "exp", 300, additional_name="_bg"); cal.save(str(tmp_path), "exp", 300)
    assert path == os.path.join(str(tmp_path), "exp", "300", "models")
    assert sorted(os.listdir(path)) == ["colorcal_model.pt", "nerf_hash_model_bg.pt", "rgb_model.pt", "sdf_model.pt"]
    sdf2 = SDF(3, None, 32, 1000, nr_levels=4, capacity=2 ** 8)
    sdf2.load_state_dict(torch.load(os.path.join(path, "sdf_model.pt")))
    assert all(torch.equal(a, b) for a, b in zip(sdf.state_dict().values(), sdf2.state_dict().values()))
