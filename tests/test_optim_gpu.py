"""Fused dense AdamW kernel (csrc/optim.cu) against torch.optim.AdamW with the reference's hyper-parameters
(train_permuto_sdf.py:293-304): same trajectories within fp32 round-off over several steps, per-group weight decay,
gradient zeroing folded into the step, flat buffers shared with autograd."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fused_adamw_matches_torch(cuda):
    from permuto_sdf_b200.optim import FusedAdamW
    torch.manual_seed(0)
    shapes = [(1000, 7), (33,), (64, 64), (5,)]
    pa = [torch.nn.Parameter(torch.randn(*s, device="cuda")) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    ga = [{"params": pa[:2], "weight_decay": 0.0, "lr": 1e-3, "name": "a"}, {"params": pa[2:], "weight_decay": 0.1, "lr": 2e-3, "name": "b"}]
    gb = [{"params": pb[:2], "weight_decay": 0.0, "lr": 1e-3}, {"params": pb[2:], "weight_decay": 0.1, "lr": 2e-3}]
    ours = FusedAdamW(ga, betas=(0.9, 0.99), eps=1e-15)
    ref = torch.optim.AdamW(gb, betas=(0.9, 0.99), eps=1e-15, amsgrad=False)
    assert all(p.grad is not None and p.grad.data_ptr() >= ours.flat_grad.data_ptr() for p in pa)
    for it in range(6):
        coef = [torch.randn_like(p) for p in pa]
        ours.zero_grad()
        loss_a = sum(((p * c) ** 2).sum() + (p * c).sum() for p, c in zip(pa, coef))
        loss_a.backward()
        ours.step()
        assert float(ours.flat_grad.abs().max()) == 0.0, "step must leave the gradient buffer zeroed"
        ref.zero_grad()
        loss_b = sum(((p * c) ** 2).sum() + (p * c).sum() for p, c in zip(pb, coef))
        loss_b.backward()
        ref.step()
        for a, b in zip(pa, pb):
            assert torch.allclose(a, b, rtol=2e-5, atol=2e-6), "diverged at step %d" % it
    # grad_scale folds the data-parallel mean into the step
    pa[0].grad.fill_(2.0)
    before = pa[0].detach().clone()
    ours.step(grad_scale=0.5)
    assert not torch.equal(before, pa[0])


def test_fused_adamw_device_step_under_cuda_graph(cuda):
    """step count kept on the device + the step captured in a CUDA graph: replays track an eagerly stepped twin"""
    from permuto_sdf_b200.optim import FusedAdamW
    torch.manual_seed(1)
    pa = [torch.nn.Parameter(torch.randn(4096, device="cuda")), torch.nn.Parameter(torch.randn(33, 7, device="cuda"))]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    oa = FusedAdamW([{"params": pa, "weight_decay": 0.01, "lr": 1e-2}], betas=(0.9, 0.99), eps=1e-15)
    ob = FusedAdamW([{"params": pb, "weight_decay": 0.01, "lr": 1e-2}], betas=(0.9, 0.99), eps=1e-15)
    oa.device_step = True
    oa.step_dev = torch.zeros(1, dtype=torch.int32, device="cuda")
    grads = [torch.randn_like(oa.flat_grad) for _ in range(5)]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                      # warm-up outside the capture
        oa.flat_grad.copy_(grads[0]); oa.step()
    torch.cuda.current_stream().wait_stream(side)
    ob.flat_grad.copy_(grads[0]); ob.step()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        oa.step()
    oa.step_count -= 1
    for k in range(1, 5):
        oa.flat_grad.copy_(grads[k]); g.replay(); oa.step_count += 1
        ob.flat_grad.copy_(grads[k]); ob.step()
        assert int(oa.step_dev) == ob.step_count == k + 1
        assert torch.allclose(oa.flat_param, ob.flat_param, rtol=1e-5, atol=1e-6), "diverged at replay %d" % k


def test_scheduler_edits_take_effect_under_graph_replay(cuda):
    """lr / weight_decay live in a device array read by the captured kernel (FusedAdamW.hyper_dev): editing param_groups between
    replays (MultiStepLR, warm-up, the wd = 1.0 switch of train_permuto_sdf.py:400-403) changes the update WITHOUT re-capturing"""
    from permuto_sdf_b200.optim import FusedAdamW
    torch.manual_seed(2)
    pa = [torch.nn.Parameter(torch.randn(2048, device="cuda"))]
    pb = [torch.nn.Parameter(pa[0].detach().clone())]
    oa = FusedAdamW([{"params": pa, "weight_decay": 0.0, "lr": 1e-2, "name": "g"}], betas=(0.9, 0.99), eps=1e-15)
    ob = FusedAdamW([{"params": pb, "weight_decay": 0.0, "lr": 1e-2, "name": "g"}], betas=(0.9, 0.99), eps=1e-15)
    oa.device_step = True
    oa.step_dev = torch.zeros(1, dtype=torch.int32, device="cuda")
    oa.sync_hyper()
    grads = [torch.randn_like(oa.flat_grad) for _ in range(6)]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        oa.flat_grad.copy_(grads[0]); oa.step()
    torch.cuda.current_stream().wait_stream(side)
    ob.flat_grad.copy_(grads[0]); ob.step()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        oa.step()
    oa.step_count -= 1
    sched = [(1e-2, 0.0), (3e-3, 0.0), (3e-3, 1.0), (1e-4, 1.0), (0.0, 0.0)]
    for k, (lr, wd) in enumerate(sched, start=1):
        for o in (oa, ob):
            o.param_groups[0]["lr"], o.param_groups[0]["weight_decay"] = lr, wd
        oa.sync_hyper()
        before = oa.flat_param.clone()
        oa.flat_grad.copy_(grads[k]); g.replay(); oa.step_count += 1
        ob.flat_grad.copy_(grads[k]); ob.step()
        assert torch.allclose(oa.flat_param, ob.flat_param, rtol=1e-5, atol=1e-6), "replay %d ignored lr=%g wd=%g" % (k, lr, wd)
        if lr == 0.0:
            assert torch.equal(before, oa.flat_param)


def test_zero_grad_after_skipped_step_and_state_roundtrip(cuda):
    """a backward that is not followed by a step must not leak into the next iteration (same as torch.optim.AdamW + zero_grad);
    state_dict / load_state_dict round trip"""
    from permuto_sdf_b200.optim import FusedAdamW
    torch.manual_seed(3)
    p = torch.nn.Parameter(torch.randn(100, device="cuda"))
    o = FusedAdamW([{"params": [p], "weight_decay": 0.0, "lr": 1e-2}], betas=(0.9, 0.99), eps=1e-15)
    o.zero_grad(); (p * 2).sum().backward(); o.step()
    o.zero_grad(); (p * 3).sum().backward()            # no step
    o.zero_grad(); (p * 5).sum().backward()
    assert torch.allclose(p.grad, torch.full_like(p, 5.0))
    o.step()
    sd = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in o.state_dict().items()}
    q = torch.nn.Parameter(p.detach().clone())
    o2 = FusedAdamW([{"params": [q], "weight_decay": 0.0, "lr": 5e-1}], betas=(0.9, 0.99), eps=1e-15)
    o2.load_state_dict(sd)
    assert o2.step_count == o.step_count and o2.param_groups[0]["lr"] == 1e-2
    for opt, par in ((o, p), (o2, q)):
        opt.zero_grad(); (par * 7).sum().backward(); opt.step()
    assert torch.allclose(p, q, rtol=1e-6, atol=1e-7)
    o.check_aliasing()
    p.data = p.data.clone()
    with pytest.raises(RuntimeError):
        o.check_aliasing()
