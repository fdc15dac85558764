"""Sphere-tracer statistics on the bench scene: time and number of network evaluations of the fused ray-queue kernel for several
iteration caps, with and without the occupancy grid. usage (under gpurun): python tools/trace_stats.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "compat")]


def main():
    import permuto_sdf_b200.train as tr
    from permuto_sdf_b200.train import HyperParams, Trainer
    t = Trainer(HyperParams(), nr_levels=16, capacity=2 ** 18, sdf_hidden=64, nr_images=8, occupancy_resolution=256, seed=0, optimizer="fused")
    t.set_analytic_scene()
    m = t.model_sdf
    m.last_iter_nr = 20000
    for i in range(400):
        loss, _, _ = tr.loss_sphere_init(30000, t.aabb, m, 20000)
        t.optimizer.zero_grad(set_to_none=False)
        loss.backward()
        t.optimizer_step()
    W, H = 1920, 1080
    f = 1.2 * W
    u, v = torch.meshgrid(torch.arange(W, device="cuda", dtype=torch.float32), torch.arange(H, device="cuda", dtype=torch.float32), indexing="xy")
    d = torch.stack([(u + 0.5 - W / 2) / f, (v + 0.5 - H / 2) / f, torch.ones_like(u)], -1).reshape(-1, 3)
    d = torch.nn.functional.normalize(d, dim=-1).contiguous()
    o = torch.tensor([0.0, 0.0, -1.5], device="cuda").expand_as(d).contiguous()
    torch.set_grad_enabled(False)
    tr.FUSED_SPHERE_TRACE_MAX_RAYS = 1 << 30
    out = []
    for grid in (t.occupancy_grid, None):
        for iters in ((16, 64, 256) if grid is not None else (256,)):
            for thresh in (1e-3,):
                for _ in range(2):
                    pts, sdf, grads, geom, rsp = tr.sphere_trace(iters, o, d, m, False, 0.9, thresh, grid)
                torch.cuda.synchronize()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                pts, sdf, grads, geom, rsp = tr.sphere_trace(iters, o, d, m, False, 0.9, thresh, grid)
                e.record()
                torch.cuda.synchronize()
                st = m.fused.last_trace_stats.tolist()
                out.append({"grid": grid is not None, "iters": iters, "thresh": thresh, "ms": s.elapsed_time(e), "rays": int(pts.shape[0]),
                            "evaluations": st[1], "evals_per_ray": st[1] / max(int(pts.shape[0]), 1), "rounds_sum": st[2], "net_rounds_sum": st[3],
                            "rounds_max_cta": st[4]})
                print(json.dumps(out[-1]), flush=True)


if __name__ == "__main__":
    main()
