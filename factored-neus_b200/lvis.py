"""Stage-2 light-visibility / traced-radiance ground truth (reference: models/calLvis.py:9-204,302-409).

Per surface point: 4 secondary rays about the normal; per secondary ray 512 coarse SDF evaluations (no grad),
32 importance depths from the inverse CDF at the learned inv_s, visibility = 1 - sum of inside-sphere weights of
those 32 samples, traced radiance = colour network at the first sign change (secant root).  Every SDF / colour
evaluation, the up-sampling step and the alpha/weight computation run in libfneus_b200.so; only tiny per-ray
index arithmetic is torch.  Shapes are fixed (non-hit rays are masked, no nonzero()/boolean indexing).
"""
from __future__ import annotations

import math

import torch

from . import ops


def sample_dirs(normals, r_theta, r_phi):
    """calLvis.py:302-320.  normals [m,1,3]; r_theta, r_phi [m,k] -> [m,k,3]."""
    tiny = 1e-6

    def nrm(v):
        return v / (torch.norm(v, dim=-1, keepdim=True) + tiny)

    n = nrm(normals)
    ex = torch.zeros_like(n)
    ex[..., 0] = 1
    u = nrm(torch.cross(ex, n, dim=-1))
    v = nrm(torch.cross(n, u, dim=-1))
    th, ph = r_theta.unsqueeze(-1), r_phi.unsqueeze(-1)
    return u * torch.cos(th) * torch.sin(ph) + v * torch.sin(th) * torch.sin(ph) + n * torch.cos(ph)


def query_indir_illum(lgtSGs, dirs):
    """calLvis.py:323-336: spherical-Gaussian radiance [n,24,7] x [n,k,3] -> [n,k,3] (tiny; torch)."""
    k, nl = dirs.shape[1], lgtSGs.shape[1]
    sg = lgtSGs.unsqueeze(-3).expand(-1, k, -1, -1)
    d = dirs.unsqueeze(-2).expand(-1, -1, nl, -1)
    lobes = sg[..., :3] / torch.norm(sg[..., :3], dim=-1, keepdim=True)
    return (sg[..., -3:] * torch.exp(sg[..., 3:4] * (torch.sum(d * lobes, dim=-1, keepdim=True) - 1.0))).sum(dim=2)


@torch.no_grad()
def trace_visibility(surf, normal, sdf_network, deviation_network, color_network, r_theta, rand_z,
                     n_coarse=512, n_imp=32, chunk_points=2048):
    """Ground-truth part of cal_indiLgt (calLvis.py:351-397).  surf, normal [m,3]; r_theta, rand_z [m,k].
    Returns gt_lvis [m,k], gt_trace_radiance [m,k,3], dirs [m,k,3].  No host read-back anywhere: the learned inv_s
    stays a device scalar."""
    dev = surf.device
    m, k = r_theta.shape
    dirs = sample_dirs(normal[:, None, :], r_theta, torch.asin(rand_z))
    inv_s = deviation_network(torch.zeros([1, 3], device=dev))[:, :1].clip(1e-6, 1e6)
    zc_row = torch.linspace(0.0, 1.0, n_coarse, device=dev)
    u = torch.linspace(0.5 / n_imp, 1.0 - 0.5 / n_imp, n_imp, device=dev)
    sample_dist = (1 - 0.1) / 32.0                              # calLvis.py:95,155
    net = sdf_network
    w_sdf = net.flat_weights().detach()
    lvis_out = torch.empty(m, k, device=dev)
    rad_out = torch.empty(m, k, 3, device=dev)
    for p0 in range(0, m, chunk_points):
        p1 = min(m, p0 + chunk_points)
        o = surf[p0:p1, None, :].expand(-1, k, -1).reshape(-1, 3).contiguous()
        d = dirs[p0:p1].reshape(-1, 3).contiguous()
        R = o.shape[0]
        zc = zc_row[None, :].expand(R, -1).contiguous()
        sdf_c = ops.sdf_forward_nograd(net.cfg, w_sdf, ops.ray_points(o, d, zc), want_feat=False)[0].reshape(R, n_coarse)
        z_fine = ops.upsample_step_dev(o, d, zc, sdf_c, n_imp, inv_s, u)
        # shared geometry of the 32 importance sections
        dists, mid_z, pts, dd = ops.core_geometry(o, d, z_fine, sample_dist)
        sdf_f, _, nrm_f = net.value_feature_normal(pts, want_normal=True, w=w_sdf)
        # (a) occlusion: alpha / weights of compute_weight (cos_anneal_ratio = 0)
        zeros_rgb = torch.zeros(R * n_imp, 3, device=dev)
        _, weights, _, _, _, _, _, _, _, _ = ops.Composite.apply(
            sdf_f, nrm_f, zeros_rgb, inv_s, None, None, dists, pts, d, None, n_imp, 0, 0.0)
        # (b) visibility = 1 - sum of inside-sphere weights; first sign change and its secant root (one launch)
        hit_idx, _, p_s, lvis, _ = ops.first_hit_secant(sdf_f, mid_z, pts, o, d, weights=weights)
        lvis_out[p0:p1] = lvis.reshape(-1, k)
        # (c) first-hit radiance: colour network at the root
        _, f_s, n_s = net.value_feature_normal(p_s, want_normal=True, w=w_sdf)
        rgb = color_network(p_s, n_s, d, f_s)
        rad_out[p0:p1] = torch.where((hit_idx >= 0)[:, None], rgb, torch.zeros_like(rgb)).reshape(-1, k, 3)
    return lvis_out, rad_out, dirs


def cal_indiLgt(surf, normal, sdf_network, deviation_network, color_network, lvis_network, indiLgt_network,
                r_theta=None, rand_z=None):
    """Drop-in for calLvis.cal_indiLgt (calLvis.py:339-409); the random angles may be supplied for reproducibility."""
    nsamp = 4
    dev = surf.device
    if r_theta is None:
        r_theta = torch.rand(surf.shape[0], nsamp, device=dev) * 2 * math.pi
    if rand_z is None:
        rand_z = torch.rand(surf.shape[0], nsamp, device=dev) * 0.95
    gt_lvis, gt_rad, dirs = trace_visibility(surf.detach(), normal.detach(), sdf_network, deviation_network,
                                             color_network, r_theta, rand_z)
    o = surf[:, None, :].repeat(1, nsamp, 1).reshape(-1, 3)
    d = dirs.reshape(-1, 3)
    pre_lvis = lvis_network(o, d).reshape(surf.shape[0], nsamp)
    pre_rad = query_indir_illum(indiLgt_network(surf), dirs)
    return {"gt_lvis": gt_lvis, "pre_lvis": pre_lvis, "gt_trace_radiance": gt_rad, "pre_trace_radiance": pre_rad}
