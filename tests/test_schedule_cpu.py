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
