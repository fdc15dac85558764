"""One training iteration of the bench workload (C2, static-capacity containers, every kernel launched eagerly) between
cudaProfilerStart / Stop, for ncu:
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'^(?!.*at::).*k_' \\
      -o gpurun_out/prof_r2_iteration python tools/profile_iteration.py
Only this library's kernels (names k_*) are captured; summarise with tools/ncu_summary.py / tools/ncu_hotspots.py."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "compat")]


def main():
    import bench
    from permuto_sdf_b200.train import HyperParams, Trainer
    dev = torch.device("cuda", 0)
    hp = HyperParams()
    hp.max_nr_samples_per_ray = bench.SAMPLES_PER_RAY - 32
    hp.nr_samples_imp_sampling = 16
    hp.nr_rays = bench.NR_RAYS
    tr = Trainer(hp, nr_levels=16, capacity=2 ** 18, sdf_hidden=64, nr_images=8, occupancy_resolution=256, seed=0, optimizer="fused")
    tr.set_analytic_scene()
    tr.iter_nr = 20001
    tr.enable_cuda_graph(warmup_steps=1 << 30)          # static shapes, device-resident generators / schedule, never captured
    reel = bench.analytic_reel(8, 600, 800, 1000.0, dev)
    gen = torch.Generator().manual_seed(1)

    def step(update_occ):
        pix = bench.central_pixels(bench.NR_RAYS, 600, 800, 300, gen).to(dev)
        img = torch.randint(0, 8, (bench.NR_RAYS,), generator=gen, dtype=torch.int32).to(dev)
        return tr.step_from_reel(reel, pix, img, update_occupancy=update_occ)
    for _ in range(4):
        step(False)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step(int(os.environ.get("WITH_OCC", "0")) != 0)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
