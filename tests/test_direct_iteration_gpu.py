"""The direct (autograd-free) training iteration (permuto_sdf_b200/iteration.py + csrc/iter_glue.cu) against the torch.autograd
formulation of the same iteration (Trainer.losses + loss.backward, train_permuto_sdf.py:311-422): identical samples and random draws,
loss and every parameter gradient compared. Plus unit checks of the folded glue kernels against the per-op kernels they replace."""
import ctypes

import numpy as np
import pytest
import torch

import scenes

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def _fresh_rngs():
    from permuto_sdf import OccupancyGrid, RaySampler, RaySamplesPacked, VolumeRendering
    from permuto_sdf_b200.permuto_sdf import _Pcg32Host
    OccupancyGrid.m_rng, RaySampler.m_rng, VolumeRendering.m_rng = _Pcg32Host(), _Pcg32Host(), _Pcg32Host()
    RaySamplesPacked.static_capacity = False


def _trainer(direct, seed=5):
    from permuto_sdf_b200.train import HyperParams, Trainer
    hp = HyperParams()
    hp.max_nr_samples_per_ray = 32
    hp.nr_samples_imp_sampling = 8
    hp.min_dist_between_samples = 1e-3
    tr = Trainer(hp, nr_levels=8, capacity=2 ** 14, sdf_hidden=64, occupancy_resolution=128, nr_images=4, seed=seed, optimizer="fused",
                 direct=direct)
    tr.set_analytic_scene()
    with torch.no_grad():       # calibration parameters start at zero: give them values so that their path is exercised
        tr.model_colorcal.weight_delta.copy_(torch.linspace(-0.2, 0.3, 12).view(4, 3))
        tr.model_colorcal.bias.copy_(torch.linspace(0.1, -0.1, 12).view(4, 3))
    return tr


def _rays(n=256):
    from permuto_sdf import PermutoSDF

    class Reel:
        pass
    rgb, mask, K, tf = scenes.synthetic_reel(nimg=4, H=60, W=80)
    reel = Reel()
    reel.rgb_reel, reel.mask_reel, reel.K_reel, reel.tf_world_cam_reel = [torch.from_numpy(a).cuda() for a in (rgb, mask, K, tf)]
    gen = torch.Generator().manual_seed(3)
    pix = torch.randint(0, 60 * 80, (n,), generator=gen, dtype=torch.int32).cuda()
    img = torch.randint(0, 4, (n,), generator=gen, dtype=torch.int32).cuda()
    with torch.no_grad():
        return PermutoSDF.rays_from_reel_indices(reel, pix, img)


@pytest.mark.parametrize("it", [3000, 20000, 60000])
def test_direct_iteration_equals_autograd_iteration(cuda, it):
    """3000: coarse-to-fine ramp + curvature term; 20000: the bench iteration (curvature, all levels); 60000: no curvature,
    Lipschitz-bound loss active"""
    o, d, gt, gm, img = _rays()
    res = {}
    fixed = {}

    def draw(name, fn):
        # the two torch-RNG draws of the iteration (off-surface points, curvature directions) are consumed in a different order by the
        # two formulations: the first run draws them, the second reuses them
        if name not in fixed:
            fixed[name] = fn()
        return fixed[name]
    for direct in (False, True):
        _fresh_rngs()
        tr = _trainer(direct)
        assert (tr._direct is not None) == direct and tr.execution["iteration"].startswith("direct kernel sequence" if direct else "torch.autograd")
        tr.draw = draw
        tr.optimizer.zero_grad(set_to_none=False)
        loss = tr.forward_backward(o, d, gt, gm, img, it)
        torch.cuda.synchronize()
        groups = {g["name"]: tr.optimizer.flat_grad[g["_off"]:g["_off"] + g["_n"]].clone() for g in tr.optimizer.param_groups}
        named = {n: p.grad.clone() for m in (tr.model_sdf, tr.model_rgb, tr.model_colorcal) for n, p in m.named_parameters() if p.grad is not None}
        res[direct] = dict(loss=float(loss), n=int(tr.last["nr_samples"]), groups=groups, named=named,
                           lr=float(tr.last["loss_rgb"]), le=float(tr.last["loss_eikonal"]), lc=float(tr.last["loss_curvature"]))
    a, b = res[True], res[False]
    assert a["n"] == b["n"] and a["n"] > 1000
    assert abs(a["loss"] - b["loss"]) <= 1e-5 * abs(b["loss"]), (a["loss"], b["loss"])
    assert abs(a["lr"] - b["lr"]) <= 1e-5 * abs(b["lr"]) and abs(a["le"] - b["le"]) <= 1e-5 * abs(b["le"]) + 1e-9
    assert abs(a["lc"] - b["lc"]) <= 1e-5 * abs(b["lc"]) + 1e-9
    for name in b["groups"]:
        assert float(b["groups"][name].abs().max()) > 0, name
        assert rel(a["groups"][name], b["groups"][name]) < 2e-4, name
    for name in b["named"]:
        if float(b["named"][name].abs().max()) > 0:
            assert rel(a["named"][name], b["named"][name]) < 1e-3, name


def test_direct_iteration_trains_under_graph_replay(cuda):
    """the replayed direct iteration: finite decreasing losses, far fewer launches than the autograd graph, optimizer counters advanced
    by the fused pack launch"""
    from permuto_sdf import PermutoSDF
    _fresh_rngs()
    tr = _trainer(True, seed=9)
    tr.iter_nr = 20000
    tr.enable_cuda_graph(warmup_steps=2)

    class Reel:
        pass
    rgb, mask, K, tf = scenes.synthetic_reel(nimg=4, H=60, W=80)
    reel = Reel()
    reel.rgb_reel, reel.mask_reel, reel.K_reel, reel.tf_world_cam_reel = [torch.from_numpy(a).cuda() for a in (rgb, mask, K, tf)]
    gen = torch.Generator().manual_seed(11)
    losses = []
    for i in range(8):
        pix = torch.randint(0, 60 * 80, (256,), generator=gen, dtype=torch.int32).cuda()
        img = torch.randint(0, 4, (256,), generator=gen, dtype=torch.int32).cuda()
        h = tr.step_from_reel(reel, pix.cpu().pin_memory() if i % 2 else pix, img.cpu().pin_memory() if i % 2 else img, loss_to_host=True)
        losses.append(float(h))                       # waits for the 4-byte copy behind the forward/backward graph only
        assert losses[-1] == float(h.detach()), "host copy of the loss differs from the device value"
    launches = tr.graph_launches_per_step()
    step_dev, it_dev = int(tr.optimizer.step_dev), float(tr._cg["it_dev"])
    tr.disable_cuda_graph()
    _fresh_rngs()
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
    assert step_dev == 8 and it_dev == 20008.0, (step_dev, it_dev)
    assert launches is not None and launches < 70, launches


def test_glue_kernels_match_the_per_op_kernels(cuda):
    from permuto_sdf_b200 import call
    from permuto_sdf_b200._lib import load_library
    lib = load_library()
    torch.manual_seed(0)
    dev = "cuda"
    # ---- Lipschitz normalisation + pack in one launch == 4 x normalize + pack
    dims = [100, 128, 128, 64, 3]
    W = [torch.randn(dims[l + 1], dims[l], device=dev) * 0.3 for l in range(4)]
    b = [torch.randn(dims[l + 1], device=dev) for l in range(4)]
    c = [torch.tensor([v], device=dev) for v in (5.0, 40.0, 3.0, 0.5)]          # clamp active on some rows, softplus threshold crossed
    nbytes = int(lib.psdf_rgb_mlp_blob_bytes(*dims))
    blob_a, blob_b = torch.zeros(nbytes, dtype=torch.uint8, device=dev), torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    weff = []
    for l in range(4):
        o = torch.empty_like(W[l])
        call("psdf_lipschitz_normalize", dims[l + 1], dims[l], W[l], c[l], o)
        weff.append(o)
    call("psdf_rgb_mlp_pack", *dims, weff[0], b[0], weff[1], b[1], weff[2], b[2], weff[3], b[3], blob_a)
    call("psdf_lipschitz_pack4", *dims, W[0], b[0], c[0], W[1], b[1], c[1], W[2], b[2], c[2], W[3], b[3], c[3], blob_b)
    assert torch.equal(blob_a, blob_b)
    # ---- backward of the four layers in one launch (accumulating, resetting G) == per-layer kernels
    G = [torch.randn_like(w) for w in W]
    ref_gw, ref_gc = [], []
    for l in range(4):
        gw, gc = torch.empty_like(W[l]), torch.zeros(1, device=dev)
        call("psdf_lipschitz_normalize_backward", dims[l + 1], dims[l], W[l], c[l], G[l], gw, gc)
        ref_gw.append(gw); ref_gc.append(gc)
    acc_w = [torch.full_like(w, 0.5) for w in W]
    acc_c = [torch.full((1,), 0.25, device=dev) for _ in range(4)]
    Gc = [g.clone() for g in G]
    args = []
    for l in range(4):
        args += [W[l], c[l], Gc[l], acc_w[l], acc_c[l]]
    lip_w = 3e-6
    call("psdf_lipschitz_backward4", *dims, *args, lip_w)
    sp = [torch.nn.functional.softplus(x.double()) for x in c]
    prod = sp[0] * sp[1] * sp[2] * sp[3]
    for l in range(4):
        assert rel(acc_w[l] - 0.5, ref_gw[l]) < 1e-5
        lip_g = lip_w * prod / sp[l] * torch.sigmoid(c[l].double())
        assert abs(float(acc_c[l]) - 0.25 - float(ref_gc[l]) - float(lip_g)) <= 1e-5 * (abs(float(ref_gc[l])) + abs(float(lip_g))) + 1e-7
        assert float(Gc[l].abs().max()) == 0.0
    # ---- two sample sets in one forward launch == two launches
    from permuto_sdf_b200.models import SDF
    from permuto_sdf import Sphere
    m = SDF(3, Sphere(0.5, [0, 0, 0]), 32, 10000, nr_levels=8, capacity=2 ** 14, hidden=64).to(dev)
    fs = m.enable_fused_inference()
    enc = m.encoding
    p0, p1 = (torch.rand(1000, 3, device=dev) - 0.5) * 0.8, (torch.rand(300, 3, device=dev) - 0.5) * 0.8
    win = m.window(20000).view(-1).contiguous()
    base = (enc.nr_levels, enc.capacity, enc.lattice_values.detach(), enc.scale_factor, enc.shift_tensor(), win, enc.concat_points_scaling,
            fs.hidden, fs.out_dim, fs.blob)
    s0, g0, f0 = torch.empty(1000, 1, device=dev), torch.empty(1000, 3, device=dev), torch.empty(1000, 32, device=dev)
    s1 = torch.empty(300, 1, device=dev)
    call("psdf_sdf_fused_forward_multi", *base, 1000, p0, s0, g0, f0, 300, p1, s1, None, None)
    r0, rg0, rf0 = fs(p0, 20000, with_gradient=True)
    r1, _, _ = fs(p1, 20000, with_gradient=True)
    assert torch.equal(s0, r0) and torch.equal(g0, rg0) and torch.equal(f0, rf0) and torch.equal(s1, r1)
    # ---- AdamW over several groups in one launch == one launch per group (device-resident counters)
    ns = [4096, 20, 1024 * 64 + 4]
    P = [torch.randn(n, device=dev) for n in ns]; Gd = [torch.randn(n, device=dev) for n in ns]
    M = [torch.rand(n, device=dev) * 0.1 for n in ns]; V = [torch.rand(n, device=dev) * 0.01 for n in ns]
    hyper = torch.tensor([[1e-3, 0.0], [5e-4, 0.1], [2e-3, 1.0]], device=dev)
    step = torch.tensor([6], dtype=torch.int32, device=dev)
    Pa, Ga, Ma, Va = [[t.clone() for t in X] for X in (P, Gd, M, V)]
    step_b = torch.tensor([7], dtype=torch.int32, device=dev)
    for i, n in enumerate(ns):
        call("psdf_adamw_step", n, Pa[i], Ga[i], Ma[i], Va[i], 0.0, 0.9, 0.99, 1e-15, 0.0, 7, step_b, hyper[i], 0.5, 1)
    k = len(ns)
    tab = lambda X: (ctypes.c_uint64 * k)(*[t.data_ptr() for t in X])
    nn = (ctypes.c_longlong * k)(*ns)
    tp, tg, tm, tv = tab(P), tab(Gd), tab(M), tab(V)
    th = (ctypes.c_uint64 * k)(*[hyper.data_ptr() + 8 * i for i in range(k)])
    call("psdf_adamw_multi_step", k, ctypes.addressof(nn), ctypes.addressof(tp), ctypes.addressof(tg), ctypes.addressof(tm),
         ctypes.addressof(tv), ctypes.addressof(th), 0.9, 0.99, 1e-15, step, 1, 0.5, 0)
    for i in range(k):
        assert torch.equal(P[i], Pa[i]) and torch.equal(M[i], Ma[i]) and torch.equal(V[i], Va[i]) and float(Gd[i].abs().max()) == 0.0
    # the low-occupancy variant (4 quadruples in flight per thread, 2 blocks per SM) computes the same thing
    P2, G2, M2, V2 = [[t.clone() for t in X] for X in (Pa, Ga, Ma, Va)]
    for i, n in enumerate(ns):
        G2[i].copy_(torch.randn(n, device=dev)); Ga[i].copy_(G2[i])
        call("psdf_adamw_step", n, Pa[i], Ga[i], Ma[i], Va[i], 0.0, 0.9, 0.99, 1e-15, 0.0, 8, torch.tensor([8], dtype=torch.int32, device=dev),
             hyper[i], 0.5, 1)
    t2 = lambda X: (ctypes.c_uint64 * k)(*[t.data_ptr() for t in X])
    a, b_, c_, d_ = t2(P2), t2(G2), t2(M2), t2(V2)
    call("psdf_adamw_multi_step", k, ctypes.addressof(nn), ctypes.addressof(a), ctypes.addressof(b_), ctypes.addressof(c_),
         ctypes.addressof(d_), ctypes.addressof(th), 0.9, 0.99, 1e-15, step, 2, 0.5, 1)
    for i in range(k):
        assert torch.equal(P2[i], Pa[i]) and torch.equal(M2[i], Ma[i]) and torch.equal(V2[i], Va[i])
    # ---- schedule scalars == map_range_val on a float32 device iteration
    from permuto_sdf_b200.models import DeviceIter, map_range_val
    it_dev = torch.tensor(12345.0, device=dev)
    prm = (ctypes.c_float * 8)(0.0, 35000.0, 0.0, 1.0, 0.0, 35000.0, 0.3, 0.8)
    kinds = (ctypes.c_int * 2)(0, 1)
    out = torch.zeros(2, device=dev)
    call("psdf_iter_scalars", 2, ctypes.addressof(prm), ctypes.addressof(kinds), it_dev, 0.0, out)
    dit = DeviceIter(12345, it_dev)
    cos = map_range_val(dit, 0.0, 35000, 0.0, 1.0)
    fv = map_range_val(dit, 0.0, 35000, 0.3, 0.8)
    inv_s = torch.exp(torch.ones((), device=dev) * fv * 10.0)
    assert float(out[0]) == float(cos) and abs(float(out[1]) - float(inv_s)) <= 2e-6 * float(inv_s)
