"""torch.profiler breakdown of one C2 training iteration (which ops / kernels the step spends its CPU and GPU time in).
usage (under gpurun): python tools/profile_step.py > gpurun_out/step_profile.txt"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "compat")]


def main():
    import bench
    from permuto_sdf import PermutoSDF
    from permuto_sdf_b200.train import HyperParams, Trainer
    dev = torch.device("cuda", 0)
    hp = HyperParams()
    hp.max_nr_samples_per_ray = bench.SAMPLES_PER_RAY - 32
    hp.nr_samples_imp_sampling = 16
    hp.nr_rays = bench.NR_RAYS
    tr = Trainer(hp, nr_levels=16, capacity=2 ** 18, sdf_hidden=64, nr_images=8, occupancy_resolution=256, seed=0, optimizer="fused")
    tr.set_analytic_scene()
    tr.iter_nr = 20000
    if os.environ.get("STATIC", "1") == "1":     # shapes of the CUDA-graph mode (kernel times as inside the replayed graph)
        from permuto_sdf import RaySamplesPacked
        RaySamplesPacked.static_capacity = True
    reel = bench.analytic_reel(8, 600, 800, 1000.0, dev)
    gen = torch.Generator().manual_seed(1)

    def step(i):
        pix = bench.central_pixels(bench.NR_RAYS, 600, 800, 300, gen).to(dev)
        img = torch.randint(0, 8, (bench.NR_RAYS,), generator=gen, dtype=torch.int32).to(dev)
        with torch.no_grad():
            o, d, gt, gm, img_idx = PermutoSDF.rays_from_reel_indices(reel, pix, img)
        return tr.step(o, d, gt, gm, img_idx, update_occupancy=False)
    for i in range(5):
        step(i)
    torch.cuda.synchronize()
    n = 5
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for i in range(n):
            step(i)
        torch.cuda.synchronize()
    ka = prof.key_averages()
    print("=== by CUDA time (per step = total / %d)" % n)
    print(ka.table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
    print("=== by self CPU time")
    print(ka.table(sort_by="self_cpu_time_total", row_limit=45, max_name_column_width=70))
    nk = sum(e.count for e in ka if e.device_type == torch.autograd.DeviceType.CUDA)
    print("device kernels+memops per step:", nk / n)
    print("=== device kernels by launch count (per step), with mean duration")
    rows = [(e.count / n, e.device_time_total / max(e.count, 1), e.key) for e in ka if e.device_type == torch.autograd.DeviceType.CUDA]
    for c, t, k in sorted(rows, reverse=True)[:60]:
        print("%7.1f  %8.2f us  %s" % (c, t, k[:110]))
    print("=== torch ops by call count (per step)")
    rows = [(e.count / n, e.key) for e in ka if e.device_type != torch.autograd.DeviceType.CUDA and e.count >= n]
    for c, k in sorted(rows, reverse=True)[:50]:
        print("%7.1f  %s" % (c, k[:100]))


if __name__ == "__main__":
    main()
