"""Reference-SHAPED model classes for tests that run where /root/reference does not exist (the GPU box): the attribute layout,
constructor arguments, parameter names and forward formulas of permuto_sdf_py/models/models.py:131-259 (SDF) and :309-391 (RGB),
written against the public `permutohedral_encoding` / `permuto_sdf` names only -- i.e. what the reference's file looks like to
`permuto_sdf_b200.patch_reference_models`. Test infrastructure; the product classes are permuto_sdf_b200/models.py."""
import sys

import numpy as np
import torch
import torch.nn.functional as F

import permutohedral_encoding as permuto_enc
from permuto_sdf import PermutoSDF
from permuto_sdf_b200.models import LipshitzMLP, leaky_relu_init, map_range_val
from permuto_sdf_b200.volume_rendering import VolumeRenderingNeus


def _enc(pos_dim, scaling, nr_levels, capacity):
    scale_list = np.geomspace(1.0, 0.0001, num=nr_levels)
    return permuto_enc.PermutoEncoding(pos_dim, capacity, nr_levels, 2, scale_list, appply_random_shift_per_level=True, concat_points=True,
                                       concat_points_scaling=scaling)


class SDF(torch.nn.Module):
    def __init__(self, in_channels, boundary_primitive, geom_feat_size_out, nr_iters_for_c2f, nr_levels=24, capacity=2 ** 18):
        super().__init__()
        self.in_channels, self.boundary_primitive, self.geom_feat_size_out = in_channels, boundary_primitive, geom_feat_size_out
        self.encoding = _enc(in_channels, 1e-3, nr_levels, capacity)
        self.sdf_shift = 1e-2
        self.mlp_sdf = torch.nn.Sequential(torch.nn.Linear(self.encoding.output_dims(), 32), torch.nn.GELU(), torch.nn.Linear(32, 32), torch.nn.GELU(),
                                           torch.nn.Linear(32, 32), torch.nn.GELU(), torch.nn.Linear(32, 1 + geom_feat_size_out))
        for m in self.mlp_sdf:
            leaky_relu_init(m, negative_slope=0.0)
        leaky_relu_init(self.mlp_sdf[-1], negative_slope=1.0)
        with torch.no_grad():
            self.mlp_sdf[-1].bias += self.sdf_shift
        self.c2f = permuto_enc.Coarse2Fine(nr_levels)
        self.nr_iters_for_c2f = nr_iters_for_c2f
        self.last_iter_nr = sys.maxsize

    def forward(self, points, iter_nr):
        self.last_iter_nr = iter_nr
        window = self.c2f(map_range_val(iter_nr, 0.0, self.nr_iters_for_c2f, 0.3, 1.0))
        y = self.mlp_sdf(self.encoding(points, window.view(-1)))
        return y[:, 0:1], y[:, -self.geom_feat_size_out:]

    def get_sdf_and_gradient(self, points, iter_nr, method="autograd"):
        with torch.set_grad_enabled(True):
            points.requires_grad_(True)
            sdf, geom = self.forward(points, iter_nr)
            g = torch.autograd.grad(outputs=sdf, inputs=points, grad_outputs=torch.ones_like(sdf), create_graph=True, retain_graph=True,
                                    only_inputs=True)[0]
        return sdf, g, geom


class RGB(torch.nn.Module):
    def __init__(self, in_channels, boundary_primitive, geom_feat_size_in, nr_iters_for_c2f, nr_levels=24, capacity=2 ** 18):
        super().__init__()
        self.in_channels, self.boundary_primitive, self.geom_feat_size_in = in_channels, boundary_primitive, geom_feat_size_in
        self.volume_renderer_neus = VolumeRenderingNeus()
        self.encoding = _enc(in_channels, 1.0, nr_levels, capacity)
        self.mlp = LipshitzMLP(self.encoding.output_dims() + 25 + 3 + geom_feat_size_in, [128, 128, 64, 3], last_layer_linear=True)
        self.c2f = permuto_enc.Coarse2Fine(nr_levels)
        self.nr_iters_for_c2f = nr_iters_for_c2f
        self.last_iter_nr = sys.maxsize

    def forward(self, points, samples_dirs, sdf_gradients, geom_feat, iter_nr, model_colorcal=None, img_indices=None, ray_start_end_idx=None):
        self.last_iter_nr = iter_nr
        window = self.c2f(map_range_val(iter_nr, 0.0, self.nr_iters_for_c2f, 0.3, 1.0))
        feat = self.encoding(points, window.view(-1))
        with torch.no_grad():
            sh = PermutoSDF.spherical_harmonics(samples_dirs, 5)
        x = torch.cat([feat, sh, F.normalize(sdf_gradients.view(-1, 3), dim=1), geom_feat], 1)
        x = self.mlp(x)
        if model_colorcal is not None:
            x = model_colorcal.calib_RGB_samples_packed(x, img_indices, ray_start_end_idx)
        return torch.sigmoid(x)
