"""Replay tests/golden/ref_rayops.npz (outputs of the reference's own CUDA kernels, see make_ref_golden.py)
through the C oracle on the CPU. Integer / index results must match exactly; float results exactly unless
the op goes through a transcendental (exp/pow/sin: libm vs CUDA differ by ulps), where the tolerance is given."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)]
import scenes  # noqa: E402

V = 32
TR = [0.0, 0.0, 0.0]


def same(a, b, what, tol=0.0):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if tol == 0.0:
        assert np.array_equal(a, b), "%s: not bit exact, max err %g" % (what, np.abs(a.astype(np.float64) - b.astype(np.float64)).max())
    else:
        err = np.abs(a.astype(np.float64) - b.astype(np.float64)).max() if a.size else 0
        assert err <= tol, "%s: max err %g > %g" % (what, err, tol)


def canon(p, names=("z", "dt", "pos", "dirs")):
    se = p.start_end
    out = {"n": (se[:, 1] - se[:, 0]).astype(np.int32), "fixed_dt": p.fixed_dt}
    for nm in names:
        a = getattr(p, nm)
        out[nm] = np.concatenate([a[s:e] for s, e in se] + [a[:0]], 0)
    return out


def cmp_canon(G, prefix, p, names=("z", "dt", "pos", "dirs"), tol=0.0):
    c = canon(p, names)
    for k, v in c.items():
        same(v, G[prefix + "." + k], prefix + "." + k, 0.0 if k == "n" else tol)


def check(path, orc):
    G = dict(np.load(path))
    o, d = G["in.o"], G["in.d"]
    pe, te, px, tx, hit = orc.sphere_ray_intersection(0.5, TR, o, d)
    for nm, v in zip(("pe", "te", "px", "tx"), (pe, te, px, tx)):
        same(v, G["sphere." + nm], "sphere." + nm)
    same(hit.astype(np.uint8), G["sphere.hit"], "sphere.hit")
    same(orc.sphere_rand_points_inside(0.5, G["randpts.phi"], G["randpts.ct"], G["randpts.u"]), G["randpts.out"], "rand_points", 1e-6)
    same(orc.occ_grid_points(V, 1.0, TR), G["grid.points"], "grid points")
    same(orc.occ_grid_points(V, 1.0, TR, None, True), G["grid.points_jitter"], "grid points jitter")
    same(orc.occ_grid_points(V, 1.0, TR, G["grid.idx"], True), G["grid.points_subset_jitter"], "grid subset jitter")
    values, occ = orc.occ_update_with_sdf(G["grid.sdf"], None, 1.0, V, 512.0, 1e-4, 0, np.ones(V ** 3, np.float32), np.ones(V ** 3, np.uint8))
    occ = occ.astype(np.uint8)
    assert int((occ != G["grid.occ"]).sum()) <= 1, "occupancy bits from update_with_sdf (<=1 ulp-of-expf flip allowed)"
    occ = G["grid.occ"]                                  # continue from the reference's bits
    v2, o2 = orc.occ_update_with_sdf(G["grid.upd_rs.sdf"], G["grid.upd_rs.idx"], 1.0, V, 300.0, 1e-4, 1, values, occ)
    same(v2, G["grid.upd_rs.values"], "random-sample sdf update values")
    assert int((o2.astype(np.uint8) != G["grid.upd_rs.occ"]).sum()) <= 1
    v3, o3 = orc.occ_update_with_density(G["grid.upd_density.density"], None, 0.9, 1.2, np.ones(V ** 3, np.float32), np.ones(V ** 3, np.uint8))
    same(v3, G["grid.upd_density.values"], "density values"); same(o3.astype(np.uint8), G["grid.upd_density.occ"], "density occ")
    same(orc.occ_check_occupancy(V, 1.0, TR, occ, G["grid.check.pts"]).astype(np.uint8), G["grid.check.out"], "check_occupancy")
    for j in (0, 1):
        p = orc.occ_samples_in_occupied_regions(V, 1.0, TR, o, d, te, tx, occ, 2e-3, 24, bool(j))
        cmp_canon(G, "occ_samples.j%d" % j, p)
    p0 = orc.packed_compact(orc.occ_samples_in_occupied_regions(V, 1.0, TR, o, d, te, tx, occ, 2e-3, 24, False))
    cmp_canon(G, "first_sample", orc.occ_first_sample_start(V, 1.0, TR, o, d, te, tx, occ))
    npos, within = orc.occ_advance_to_next_occupied(V, 1.0, TR, d, G["advance.pos"], occ)
    same(npos, G["advance.out"], "advance pos"); same(within.astype(np.uint8), G["advance.within"], "advance within")
    for j in (0, 1):
        cmp_canon(G, "sampler_fg.j%d" % j, orc.sampler_fg(o, d, te, tx, 0.02, 20, bool(j)))
        b = orc.sampler_bg(o, d, tx, 8, 0.5, TR, bool(j), False)
        hitm = np.repeat(hit[:, 0], 8)                  # rays that miss have t_exit = 0 -> 0/0 paths, skip them
        for nm, v in (("z", b.z), ("dt", b.dt), ("pos", b.pos), ("pos4", b.pos4)):
            g = G["sampler_bg.j%d.%s" % (j, nm)]
            if j == 0 and nm != "pos4":
                same(v[hitm], g[hitm], "sampler_bg " + nm)          # bit exact without jitter
            else:
                # jitter: t = 1 - i*dt + mov cancels near t -> 0, so z = t_exit / t amplifies a 1-ulp difference
                assert np.allclose(v[hitm], g[hitm], rtol=1e-4, atol=1e-4 * float(np.abs(g[hitm]).max()) * 1e-2), "sampler_bg " + nm
    # volume rendering (p0 is compact and ray ordered == the canonical order of the fixture)
    alpha, w, rgb = G["vr.alpha"], G["vr.w"], G["vr.rgb"]
    T, bg = orc.vr_cumprod(p0, alpha)
    same(T, G["vr.T"], "T"); same(bg, G["vr.bg"], "bg_T")
    same(orc.vr_integrate(p0, rgb, w), G["vr.integrate"], "integrate")
    sr, ss = orc.vr_sum(p0, w)
    same(sr, G["vr.sum_ray"], "sum ray"); same(ss, G["vr.sum_sample"], "sum sample")
    same(orc.vr_sum(p0, rgb)[0], G["vr.sum_ray3"], "sum ray (3)")
    same(orc.vr_cumsum(p0, w, False), G["vr.cumsum_f"], "cumsum"); same(orc.vr_cumsum(p0, w, True), G["vr.cumsum_r"], "rev cumsum")
    same(orc.vr_cdf(p0, w), G["vr.cdf"], "cdf")
    same(orc.vr_compute_dt(p0, tx, True), G["vr.dt_exit"], "compute_dt")
    same(orc.vr_sdf2alpha(p0, G["vr.sdf"], 512.0, True, 1.0), G["vr.sdf2alpha"], "sdf2alpha dynamic", 5e-6)
    same(orc.vr_sdf2alpha(p0, G["vr.sdf"], 64.0, False, 2.0), G["vr.sdf2alpha_fixed"], "sdf2alpha fixed", 5e-6)
    cs = orc.vr_cumsum(p0, G["vr.gT"] * T, True)
    same(orc.vr_cumprod_backward(p0, G["vr.gbg"], alpha, bg, cs), G["vr.cumprod_bwd"], "cumprod backward", 1e-5)
    gv, gw = orc.vr_integrate_backward(p0, G["vr.gp"], rgb, w, True)       # reference behaviour incl. the blue-channel bug
    same(gv, G["vr.int_bwd_vals"], "integrate backward vals"); same(gw, G["vr.int_bwd_w"], "integrate backward w")
    pr, dp, nbg, nw = orc.vr_render_nerf(p0, rgb, G["vr.rad"])
    same(pr, G["vr.nerf.rgb"], "nerf rgb", 3e-5); same(dp, G["vr.nerf.depth"], "nerf depth", 3e-5)
    same(nbg, G["vr.nerf.bg"], "nerf bg", 3e-5); same(nw, G["vr.nerf.w"], "nerf w", 3e-5)
    cdf = G["imp.cdf"]
    for j in (0, 1):
        q = orc.vr_importance_sample(o, d, p0, cdf, 8, bool(j))
        same(q.z, G["imp.j%d.z" % j], "importance z"); same(q.pos, G["imp.j%d.pos" % j], "importance pos")
    q = orc.vr_importance_sample(o, d, p0, cdf, 8, False)
    p0.sdf, p0.has_sdf = G["vr.sdf"], True
    q.sdf, q.has_sdf = G["imp.sdf"], True
    cmp_canon(G, "combine", orc.vr_combine(o, d, tx, p0, q), ("z", "dt", "pos", "sdf"))
    for deg in (1, 3, 5, 7):
        same(orc.spherical_harmonics(G["sh.dirs"], deg), G["sh.deg%d" % deg], "SH degree %d" % deg, 2e-6)
    rgbr, mask, K, tf = scenes.synthetic_reel(nimg=3, H=20, W=24)
    ro, rd, rgt, rgm = orc.random_rays_from_reel(rgbr, mask, K, tf, G["reel.pix"], G["reel.img"])
    same(ro, G["reel.o"], "reel origins"); same(rd, G["reel.d"], "reel dirs", 1e-6); same(rgt, G["reel.gt"], "reel gt"); same(rgm, G["reel.gm"], "reel mask")
    return True


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import rayops as orc
    check(os.path.join(HERE, "ref_rayops.npz"), orc)
    print("oracle matches the reference-made fixtures")
