"""TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

numpy front-end of the C restatement in rayops_oracle.c (built by __graft_entry__.build_oracle() into
oracle/_build/librayops_oracle.so). Arrays are passed as contiguous numpy buffers; functions return numpy
arrays shaped like the reference's tensors. Only tests/, smoke() and bench.py --impl reference /
cpu_baseline may import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

PCG_STATE = 0x853C49E6748FEA9B
PCG_INC = 0xDA3E39CB94B95BDB


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "_build", "librayops_oracle.so")
        src = os.path.join(_HERE, "rayops_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            os.makedirs(os.path.dirname(so), exist_ok=True)
            subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-fvisibility=hidden", "-o", so, src, "-lm"])
        _LIB = ctypes.CDLL(so)
    return _LIB


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.c_void_p)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _c(v):
    if isinstance(v, np.ndarray) or v is None:
        return _p(v)
    if isinstance(v, float):
        return ctypes.c_float(v)
    if isinstance(v, (bool, np.bool_)):
        return ctypes.c_int(int(v))
    if isinstance(v, (int, np.integer)):
        v = int(v)
        return ctypes.c_uint64(v) if v > 0x7FFFFFFF else ctypes.c_int(v)
    raise TypeError(type(v))


def _call(name, *args, restype=None):
    fn = getattr(lib(), name)
    fn.restype = restype
    return fn(*[_c(a) for a in args])


def U64(v):
    return ctypes.c_uint64(int(v))


# ---------------------------------------------------------------------------------------------------------
def pcg32_draw(n, state=PCG_STATE, inc=PCG_INC, advance=0):
    u = np.zeros(n, np.uint32)
    f = np.zeros(n, np.float32)
    lib().oracle_pcg32_draw(U64(state), U64(inc), ctypes.c_int64(advance), ctypes.c_int(n), _p(u), _p(f), None)
    return u, f


def morton(xyz):
    xyz = np.ascontiguousarray(xyz, dtype=np.uint32)
    n = xyz.shape[0]
    code = np.zeros(n, np.uint32)
    back = np.zeros((n, 3), np.uint32)
    lib().oracle_morton(ctypes.c_int(n), _p(xyz), _p(code), _p(back))
    return code, back


def sphere_ray_intersection(radius, center, origins, dirs):
    o, d = f32(origins), f32(dirs)
    n = o.shape[0]
    pe, px = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
    te, tx = np.zeros((n, 1), np.float32), np.zeros((n, 1), np.float32)
    hit = np.zeros((n, 1), np.uint8)
    _call("oracle_sphere_ray_intersection", n, float(radius), f32(center), o, d, pe, te, px, tx, hit)
    return pe, te, px, tx, hit.astype(bool)


def sphere_rand_points_inside(radius, phi, costheta, u):
    n = phi.shape[0]
    pts = np.zeros((n, 3), np.float32)
    _call("oracle_sphere_rand_points_inside", n, float(radius), f32(phi), f32(costheta), f32(u), pts)
    return pts


def occ_grid_points(V, extent, trans, idx=None, randomize=False, state=PCG_STATE, inc=PCG_INC):
    n = V ** 3 if idx is None else idx.shape[0]
    out = np.zeros((n, 3), np.float32)
    lib().oracle_occ_grid_points(ctypes.c_int(n), ctypes.c_int(V), ctypes.c_float(extent), _p(f32(trans)),
                                 _p(None if idx is None else i32(idx)), U64(state), U64(inc), ctypes.c_int(int(randomize)), _p(out))
    return out


def occ_update_with_density(density, idx, decay, thresh, values, occ):
    values, occ = f32(values).copy(), u8(occ).copy()
    n = density.shape[0]
    _call("oracle_occ_update_with_density", n, f32(density), None if idx is None else i32(idx), float(decay), float(thresh), values, occ)
    return values, occ.astype(bool)


def occ_update_with_sdf(sdf, idx, extent, V, inv_s, thresh, random_variant, values, occ):
    values, occ = f32(values).copy(), u8(occ).copy()
    n = sdf.shape[0]
    _call("oracle_occ_update_with_sdf", n, f32(sdf), None if idx is None else i32(idx), float(extent), int(V), float(inv_s),
          float(thresh), int(random_variant), values, occ)
    return values, occ.astype(bool)


def occ_check_occupancy(V, extent, trans, occ, pts):
    n = pts.shape[0]
    out = np.zeros((n, 1), np.uint8)
    _call("oracle_occ_check_occupancy", n, int(V), float(extent), f32(trans), u8(occ), f32(pts), out)
    return out.astype(bool)


class Packed:
    """numpy mirror of RaySamplesPacked"""

    def __init__(self, nr_rays, M):
        self.pos = np.zeros((M, 3), np.float32)
        self.dirs = np.zeros((M, 3), np.float32)
        self.z = np.zeros((M, 1), np.float32)
        self.dt = np.zeros((M, 1), np.float32)
        self.sdf = np.zeros((M, 1), np.float32)
        self.pos4 = None
        self.fixed_dt = np.zeros((nr_rays, 1), np.float32)
        self.start_end = np.zeros((nr_rays, 2), np.int32)
        self.max_nr_samples = M
        self.cur = 0
        self.equal = False
        self.fixed_n = 0
        self.has_sdf = False

    def rsp(self):
        return (self.start_end.shape[0], self.max_nr_samples, self.start_end, int(self.equal), int(self.fixed_n))


def occ_samples_in_occupied_regions(V, extent, trans, origins, dirs, t_entry, t_exit, occ, min_dist, max_per_ray, jitter=False,
                                    state=PCG_STATE, inc=PCG_INC, strided_slots=True):
    o, d = f32(origins), f32(dirs)
    R = o.shape[0]
    M = max(R * max_per_ray, 1)
    p = Packed(R, M)
    forced = i32(np.arange(R) * max_per_ray) if strided_slots else None
    fn = lib().oracle_occ_compute_samples_in_occupied_regions
    fn.restype = ctypes.c_int
    p.cur = fn(ctypes.c_int(R), ctypes.c_int(V), ctypes.c_float(extent), _p(f32(trans)), _p(o), _p(d), _p(f32(t_entry)),
               _p(f32(t_exit)), _p(u8(occ)), ctypes.c_float(min_dist), ctypes.c_int(max_per_ray), ctypes.c_int(M), U64(state),
               U64(inc), ctypes.c_int(int(jitter)), _p(p.pos), _p(p.dirs), _p(p.z), _p(p.dt), _p(p.fixed_dt), _p(p.start_end),
               _p(forced))
    return p


def occ_first_sample_start(V, extent, trans, origins, dirs, t_entry, t_exit, occ):
    o, d = f32(origins), f32(dirs)
    R = o.shape[0]
    p = Packed(R, max(R, 1))
    forced = i32(np.arange(R))
    fn = lib().oracle_occ_compute_first_sample_start
    fn.restype = ctypes.c_int
    p.cur = fn(ctypes.c_int(R), ctypes.c_int(V), ctypes.c_float(extent), _p(f32(trans)), _p(o), _p(d), _p(f32(t_entry)),
               _p(f32(t_exit)), _p(u8(occ)), ctypes.c_int(max(R, 1)), _p(p.pos), _p(p.dirs), _p(p.z), _p(p.dt), _p(p.fixed_dt),
               _p(p.start_end), _p(forced))
    return p


def occ_advance_to_next_occupied(V, extent, trans, dirs, pos, occ):
    pos = f32(pos).copy()
    n = pos.shape[0]
    within = np.zeros((n, 1), np.uint8)
    _call("oracle_occ_advance_sample_to_next_occupied_voxel", n, int(V), float(extent), f32(trans), f32(dirs), pos, u8(occ), within)
    return pos, within.astype(bool)


def packed_compact(p):
    R = p.start_end.shape[0]
    exact = int((p.start_end[:, 1] - p.start_end[:, 0]).sum())
    q = Packed(R, exact)
    q.equal, q.fixed_n, q.has_sdf = p.equal, p.fixed_n, p.has_sdf
    if p.pos4 is not None:
        q.pos4 = np.zeros((exact, 4), np.float32)
    fn = lib().oracle_packed_compact
    fn.restype = ctypes.c_int
    q.cur = fn(ctypes.c_int(R), _p(p.pos), _p(p.pos4), _p(p.dirs), _p(p.z), _p(p.dt), _p(p.sdf), _p(p.fixed_dt), _p(p.start_end),
               _p(q.pos), _p(q.pos4), _p(q.dirs), _p(q.z), _p(q.dt), _p(q.sdf), _p(q.fixed_dt), _p(q.start_end))
    return q


def packed_per_sample_ray_idx(start_end, nr_samples):
    out = np.zeros(nr_samples, np.int32)
    _call("oracle_packed_per_sample_ray_idx", start_end.shape[0], i32(start_end), out)
    return out


def sampler_fg(origins, dirs, t_entry, t_exit, min_dist, max_per_ray, jitter=False, state=PCG_STATE, inc=PCG_INC):
    o, d = f32(origins), f32(dirs)
    R = o.shape[0]
    M = max(R * max_per_ray, 1)
    p = Packed(R, M)
    forced = i32(np.arange(R) * max_per_ray)
    fn = lib().oracle_sampler_fg
    fn.restype = ctypes.c_int
    p.cur = fn(ctypes.c_int(R), _p(o), _p(d), _p(f32(t_entry)), _p(f32(t_exit)), ctypes.c_float(min_dist), ctypes.c_int(max_per_ray),
               ctypes.c_int(M), U64(state), U64(inc), ctypes.c_int(int(jitter)), _p(p.pos), _p(p.dirs), _p(p.z), _p(p.dt),
               _p(p.fixed_dt), _p(p.start_end), _p(forced))
    return p


def sampler_bg(origins, dirs, t_exit, n_per_ray, radius, center, randomize=False, contract=False, state=PCG_STATE, inc=PCG_INC):
    o, d = f32(origins), f32(dirs)
    R = o.shape[0]
    p = Packed(R, R * n_per_ray)
    p.pos4 = np.zeros((R * n_per_ray, 4), np.float32)
    p.equal, p.fixed_n = True, n_per_ray
    lib().oracle_sampler_bg(ctypes.c_int(R), ctypes.c_int(n_per_ray), _p(o), _p(d), _p(f32(t_exit)), ctypes.c_float(radius),
                            _p(f32(center)), U64(state), U64(inc), ctypes.c_int(int(randomize)), ctypes.c_int(int(contract)),
                            _p(p.pos), _p(p.pos4), _p(p.dirs), _p(p.z), _p(p.dt), _p(p.fixed_dt), _p(p.start_end))
    return p


# ---- volume rendering; `p` is a Packed, per-sample arrays are [N,c] float32 -----------------------------------------
def vr_cumprod(p, alpha):
    N, R = p.z.shape[0], p.start_end.shape[0]
    T = np.zeros((N, 1), np.float32)
    bg = np.ones((R, 1), np.float32)
    _call("oracle_vr_cumprod_alpha2transmittance", *p.rsp(), f32(alpha), T, bg)
    return T, bg


def vr_integrate(p, vals, w):
    out = np.zeros((p.start_end.shape[0], 3), np.float32)
    _call("oracle_vr_integrate_with_weights", *p.rsp(), f32(vals), f32(w), out)
    return out


def vr_sdf2alpha(p, sdf, inv_s, dynamic, mult):
    alpha = np.zeros((p.z.shape[0], 1), np.float32)
    R, M, se, eq, fn = p.rsp()
    _call("oracle_vr_sdf2alpha", R, M, se, p.fixed_dt, p.dt, eq, fn, f32(sdf), float(inv_s), int(dynamic), float(mult), alpha)
    return alpha


def vr_sum(p, vals):
    vals = f32(vals)
    D = vals.shape[1]
    s_ray = np.zeros((p.start_end.shape[0], D), np.float32)
    s_smp = np.zeros_like(vals)
    R, M, se, eq, fn = p.rsp()
    _call("oracle_vr_sum_over_each_ray", R, M, se, eq, fn, D, vals, s_ray, s_smp)
    return s_ray, s_smp


def vr_cumsum(p, vals, inverse):
    out = np.zeros((p.z.shape[0], 1), np.float32)
    R, M, se, eq, fn = p.rsp()
    _call("oracle_vr_cumsum_over_each_ray", R, M, se, eq, fn, f32(vals), int(inverse), out)
    return out


def vr_cdf(p, w):
    out = np.zeros((p.z.shape[0], 1), np.float32)
    _call("oracle_vr_compute_cdf", *p.rsp(), f32(w), out)
    return out


def vr_importance_sample(origins, dirs, p, cdf, nr_imp, jitter=False, state=PCG_STATE, inc=PCG_INC):
    R = p.start_end.shape[0]
    q = Packed(R, R * nr_imp)
    q.equal, q.fixed_n = True, nr_imp
    R_, M, se, eq, fn = p.rsp()
    lib().oracle_vr_importance_sample(ctypes.c_int(R), _p(f32(origins)), _p(f32(dirs)), ctypes.c_int(M), _p(se), _p(p.fixed_dt),
                                      ctypes.c_int(eq), ctypes.c_int(fn), _p(p.z), _p(f32(cdf)), ctypes.c_int(nr_imp), U64(state),
                                      U64(inc), ctypes.c_int(int(jitter)), _p(q.pos), _p(q.dirs), _p(q.z))
    q.start_end = np.stack([np.arange(R) * nr_imp, (np.arange(R) + 1) * nr_imp], 1).astype(np.int32)
    return q


def vr_combine(origins, dirs, t_exit, p, q):
    R = p.start_end.shape[0]
    c_max = max(p.z.shape[0] + R * q.fixed_n, 1)
    c = Packed(R, c_max)
    c.has_sdf = p.has_sdf
    n = p.start_end[:, 1] - p.start_end[:, 0]
    cnt = np.where(n > 1, n + q.fixed_n, 0)
    forced = i32(np.concatenate([[0], np.cumsum(cnt)[:-1]]))
    R_, M, se, eq, fn = p.rsp()
    f = lib().oracle_vr_combine_uniform_samples_with_imp
    f.restype = ctypes.c_int
    c.cur = f(ctypes.c_int(R), _p(f32(origins)), _p(f32(dirs)), _p(f32(t_exit)), ctypes.c_int(M), _p(se), _p(p.fixed_dt),
              ctypes.c_int(eq), ctypes.c_int(fn), _p(p.z), _p(p.sdf), ctypes.c_int(int(p.has_sdf)), ctypes.c_int(q.fixed_n), _p(q.z),
              _p(q.sdf), ctypes.c_int(int(q.has_sdf)), ctypes.c_int(c_max), _p(c.pos), _p(c.dirs), _p(c.z), _p(c.dt), _p(c.sdf),
              _p(c.fixed_dt), _p(c.start_end), _p(forced))
    return c


def vr_cumprod_backward(p, g_bg, alpha, bg_T, cumsumLV):
    g = np.zeros((p.z.shape[0], 1), np.float32)
    _call("oracle_vr_cumprod_backward", *p.rsp(), f32(g_bg), f32(alpha), f32(bg_T), f32(cumsumLV), g)
    return g


def vr_integrate_backward(p, g_pred, vals, w, reference_bug=False):
    N = p.z.shape[0]
    g_vals, g_w = np.zeros((N, 3), np.float32), np.zeros((N, 1), np.float32)
    _call("oracle_vr_integrate_backward", *p.rsp(), f32(g_pred), f32(vals), f32(w), int(reference_bug), g_vals, g_w)
    return g_vals, g_w


def vr_sum_backward(p, g_ray, g_sample):
    g_sample = f32(g_sample)
    D = g_sample.shape[1]
    g = np.zeros_like(g_sample)
    R, M, se, eq, fn = p.rsp()
    _call("oracle_vr_sum_backward", R, M, se, eq, fn, D, f32(g_ray), g_sample, g)
    return g


def vr_compute_dt(p, t_exit, use_t_exit):
    dt = np.zeros((p.z.shape[0], 1), np.float32)
    R, M, se, eq, fn = p.rsp()
    _call("oracle_vr_compute_dt", R, int(use_t_exit), f32(t_exit), M, p.z, se, eq, fn, dt)
    return dt


def vr_render_nerf(p, rgb, radiance):
    N, R = p.z.shape[0], p.start_end.shape[0]
    pred, depth, bg, w = np.zeros((R, 3), np.float32), np.zeros((R, 1), np.float32), np.zeros((R, 1), np.float32), np.zeros((N, 1), np.float32)
    R_, M, se, eq, fn = p.rsp()
    _call("oracle_vr_volume_render_nerf", R, f32(rgb), f32(radiance), M, p.z, p.dt, se, eq, fn, pred, depth, bg, w)
    return pred, depth, bg, w


def vr_render_nerf_backward(p, g_pred, g_bg, pred, bg, rgb, radiance):
    N = p.z.shape[0]
    g_rgb, g_rad = np.zeros((N, 3), np.float32), np.zeros((N, 1), np.float32)
    R, M, se, eq, fn = p.rsp()
    _call("oracle_vr_volume_render_nerf_backward", R, f32(g_pred), f32(g_bg), f32(pred), f32(bg), f32(rgb), f32(radiance), M, p.dt, se,
          eq, fn, g_rgb, g_rad)
    return g_rgb, g_rad


def spherical_harmonics(dirs, degree):
    d = f32(dirs)
    out = np.zeros((d.shape[0], degree * degree), np.float32)
    _call("oracle_spherical_harmonics", d.shape[0], int(degree), d, out)
    return out


def random_rays_from_reel(rgb_reel, mask_reel, K, tf, pix, img):
    R = pix.shape[0]
    nimg, _, H, W = rgb_reel.shape
    o, d, gt, gm = (np.zeros((R, 3), np.float32), np.zeros((R, 3), np.float32), np.zeros((R, 3), np.float32), np.zeros((R, 1), np.float32))
    _call("oracle_random_rays_from_reel", R, nimg, H, W, f32(rgb_reel), None if mask_reel is None else f32(mask_reel), f32(K), f32(tf),
          i32(pix), i32(img), int(mask_reel is not None), o, d, gt, gm)
    return o, d, gt, gm
