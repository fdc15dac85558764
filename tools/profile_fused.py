"""Driver for ncu captures of the fused SDF kernels and the encoding kernels at the C2 size (65 536 samples along rays).
usage (under gpurun):
  ncu --set full --clock-control none --import-source on -k regex:'k_sdf_fused|k_enc_' -s 6 -c 6 -o gpurun_out/prof_fused python tools/profile_fused.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "compat")]
import scenes  # noqa: E402


def main():
    from permuto_sdf import Sphere
    from permuto_sdf_b200.models import SDF
    torch.manual_seed(0)
    R, S = 512, 128
    o, d = scenes.make_rays(R, seed=0, miss_fraction=0.0, axis_aligned=0)
    z = np.linspace(0.75, 1.0, S, dtype=np.float32)
    pos = torch.from_numpy((o[:, None, :] + z[None, :, None] * d[:, None, :]).reshape(-1, 3).astype(np.float32)).cuda()
    m = SDF(3, Sphere(0.5, [0, 0, 0]), 32, 10000, nr_levels=16, capacity=2 ** 18, hidden=64).to("cuda")
    m.enable_fused_training()
    reps = int(os.environ.get("REPS", "3"))
    for _ in range(reps):
        with torch.no_grad():
            m(pos, 20000)                                        # value only
        sdf, grad, geom = m.get_sdf_and_gradient(pos, 20000)     # value + tangents
        loss = ((grad.norm(dim=-1) - 1) ** 2).mean() + sdf.mean() + geom.mean()
        loss.backward()                                          # fused backward
        # modular encoding kernels for comparison
        feat = m.encoding(pos.clone().requires_grad_(True))
        feat.sum().backward()
    torch.cuda.synchronize()
    # timing outside the profiler
    if os.environ.get("TIME", "0") == "1":
        def t(fn, n=20):
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(n):
                fn()
            e.record()
            torch.cuda.synchronize()
            return s.elapsed_time(e) / n * 1e3
        with torch.no_grad():
            print("fused value-only us", t(lambda: m(pos, 20000)))
            print("fused value+tangents us", t(lambda: m.fused(pos, 20000, with_gradient=True)))
        def fb():
            sdf, grad, geom = m.get_sdf_and_gradient(pos, 20000)
            (sdf.mean() + grad.sum() + geom.mean()).backward()
        print("fused fwd+bwd (incl. autograd glue, dW GEMMs) us", t(fb))


if __name__ == "__main__":
    main()
