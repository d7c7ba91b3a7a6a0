"""Stage-2 light-visibility / traced-radiance ground truth (reference: models/calLvis.py:9-204,302-409).

Per surface point: 4 secondary rays about the normal; per secondary ray 512 coarse SDF evaluations (no grad),
32 importance depths from the inverse CDF at the learned inv_s, visibility = 1 - sum of inside-sphere weights of
those 32 samples, traced radiance = colour network at the first sign change (secant root).  The whole trace is ONE
call into libfneus_b200.so (``fneus_lvis_trace``, csrc/lvis.cu); only the direction sampling (a dozen element-wise ops on
[m,4,3]) is torch.  Shapes are fixed (non-hit rays are masked, no nonzero()/boolean indexing).
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
                     n_coarse=512, n_imp=32, rays_per_chunk=8192):
    """Ground-truth part of cal_indiLgt (calLvis.py:351-397).  surf, normal [m,3]; r_theta, rand_z [m,k].
    Returns gt_lvis [m,k], gt_trace_radiance [m,k,3], dirs [m,k,3].  One library call (``fneus_lvis_trace``): no host
    read-back, no per-chunk temporaries; the learned inv_s stays a device scalar."""
    dev = surf.device
    dirs = sample_dirs(normal[:, None, :], r_theta, torch.asin(rand_z)).contiguous()
    inv_s = deviation_network(torch.zeros([1, 3], device=dev))[:, :1].clip(1e-6, 1e6)
    # constant tables from torch.linspace on the device under test
    z_table = torch.linspace(0.0, 1.0, n_coarse, device=dev)
    u_table = torch.linspace(0.5 / n_imp, 1.0 - 0.5 / n_imp, n_imp, device=dev)
    lvis, rad, _ = ops.lvis_trace(sdf_network.cfg, sdf_network.flat_weights().detach(), color_network.cfg,
                                  color_network.flat_weights().detach(), surf, dirs, inv_s, z_table, u_table, rays_per_chunk)
    return lvis, rad, dirs


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
