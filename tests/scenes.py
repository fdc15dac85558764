"""Deterministic synthetic inputs shared by the CPU and GPU tests (numpy only, no GPU, no reference)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SPHERE_RADIUS = 0.5           # bounding primitive, permuto_sdf_py/utils/common_utils.py:519
OBJECT_RADIUS = 0.3           # analytic SDF ||x|| - 0.3


def make_rays(R, seed=0, miss_fraction=0.1, axis_aligned=2):
    """R rays from cameras on a radius-1.2 sphere looking (roughly) at the origin. A fraction is aimed
    away so that they miss the bounding sphere; a few are axis aligned (zero direction components)."""
    rng = np.random.RandomState(seed)
    cam = rng.randn(R, 3)
    cam /= np.linalg.norm(cam, axis=1, keepdims=True)
    cam *= 1.2
    target = rng.uniform(-0.35, 0.35, size=(R, 3))
    nmiss = int(R * miss_fraction)
    if nmiss:
        target[:nmiss] = cam[:nmiss] + rng.randn(nmiss, 3)
    d = target - cam
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = cam.astype(np.float32)
    d = d.astype(np.float32)
    for k in range(min(axis_aligned, R)):
        o[R - 1 - k] = np.array([0.05 * k, -0.02, 1.2], np.float32)
        d[R - 1 - k] = np.array([0.0, 0.0, -1.0], np.float32)
    # renormalise in float32 like a float pipeline would
    d = (d / np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)).astype(np.float32)
    return o, d


def analytic_sdf(p):
    return (np.linalg.norm(p.astype(np.float64), axis=1, keepdims=True) - OBJECT_RADIUS).astype(np.float32)


def analytic_occupancy(V, inv_s=512.0, thresh=1e-4):
    """occupancy of the analytic sphere through the oracle's update_with_sdf (OccupancyGridGPU.cuh:387-445)"""
    from oracle import rayops as orc
    pts = orc.occ_grid_points(V, 1.0, [0, 0, 0])
    sdf = analytic_sdf(pts)
    values = np.ones(V ** 3, np.float32)
    occ = np.ones(V ** 3, np.uint8)
    values, occ = orc.occ_update_with_sdf(sdf, None, 1.0, V, inv_s, thresh, 0, values, occ)
    return values, occ


def synthetic_reel(nimg=4, H=60, W=80, seed=1):
    """TensorReel-shaped tensors: rgb [I,3,H,W], mask [I,1,H,W], K [I,3,3], tf_world_cam [I,4,4]"""
    rng = np.random.RandomState(seed)
    rgb = rng.rand(nimg, 3, H, W).astype(np.float32)
    mask = (rng.rand(nimg, 1, H, W) > 0.3).astype(np.float32)
    K = np.zeros((nimg, 3, 3), np.float32)
    tf = np.zeros((nimg, 4, 4), np.float32)
    for i in range(nimg):
        f = 70.0 + 5 * i
        K[i] = np.array([[f, 0, W / 2.0], [0, f, H / 2.0], [0, 0, 1]], np.float32)
        c = rng.randn(3); c = 1.2 * c / np.linalg.norm(c)
        zaxis = -c / np.linalg.norm(c)
        up = np.array([0, 1.0, 0])
        xaxis = np.cross(up, zaxis); xaxis /= np.linalg.norm(xaxis)
        yaxis = np.cross(zaxis, xaxis)
        tf[i, :3, 0], tf[i, :3, 1], tf[i, :3, 2], tf[i, :3, 3] = xaxis, yaxis, zaxis, c
        tf[i, 3, 3] = 1
    return rgb, mask, K, tf
