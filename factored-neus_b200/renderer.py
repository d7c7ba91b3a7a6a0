"""Drop-in ``NeuSRenderer`` (reference: models/renderer.py:80-500) on the CUDA library.

Same constructor kwargs, attribute names, method names and output dictionaries as the reference;
the arithmetic runs in libfneus_b200.so through ``ops``.  There is no CPU path.
"""
from __future__ import annotations

import torch

from . import ops


class NeuSRenderer:
    # route the sparse (RefColor) gradient of `feature` into the colour network's dense one in place (ops.FanOut)
    fuse_feature_fanout = True
    # hand `feature_vector` from the SDF network to the colour network as an operand image whenever both networks can
    # (tensor-core precision, 256 features); False keeps the FP32 [N,256] tensor of the reference
    feature_image = True

    def __init__(self, n_samples, n_importance, n_outside, up_sample_steps, perturb, nerf=None, sdf_network=None,
                 deviation_network=None, color_network=None, refColor_network=None, lvis_network=None,
                 indiLgt_network=None, mateIllu_network=None):
        self.nerf = nerf
        self.sdf_network = sdf_network
        self.deviation_network = deviation_network
        self.color_network = color_network
        self.refColor_network = refColor_network
        self.lvis_network = lvis_network
        self.indiLgt_network = indiLgt_network
        self.mateIllu_network = mateIllu_network
        self.n_samples = n_samples
        self.n_importance = n_importance
        self.n_outside = n_outside
        self.up_sample_steps = up_sample_steps
        self.perturb = perturb
        self._tables = {}

    def set_precision(self, mode):
        """Precision of THIS renderer's networks ('fp32' | 'bf16'): a per-network configuration field, not a process
        switch -- another renderer of the same process keeps its own."""
        nets = [self.nerf, self.sdf_network, self.color_network, self.refColor_network, self.lvis_network,
                self.indiLgt_network]
        ops.set_precision(mode, [n for n in nets if n is not None and hasattr(n, "cfg")])

    # ------------------------------------------------------------------ helpers
    def _linspace(self, lo, hi, n, device):
        """Small constant tables come from torch.linspace on the device under test."""
        key = (float(lo), float(hi), int(n), str(device))
        t = self._tables.get(key)
        if t is None:
            t = torch.linspace(lo, hi, n, device=device, dtype=torch.float32).contiguous()
            self._tables[key] = t
        return t

    def _sdf_nograd(self, pts):
        net = self.sdf_network
        w = getattr(self, "_w_sdf", None)                    # packed once per render() (weight-norm + flat pack)
        if w is None:
            w = net.flat_weights()
        return ops.sdf_forward_nograd(net.cfg, w.detach(), pts, want_feat=False)[0]

    # ------------------------------------------------------------------ sampling
    def up_sample(self, rays_o, rays_d, z_vals, sdf, n_importance, inv_s):
        """renderer.py:152-189."""
        B, n = z_vals.shape
        u = self._linspace(0.5 / n_importance, 1.0 - 0.5 / n_importance, n_importance, z_vals.device)
        with torch.no_grad():
            return ops.upsample_step(rays_o.contiguous(), rays_d.contiguous(), z_vals.contiguous(),
                                     sdf.reshape(B, n).contiguous(), n_importance, float(inv_s), u)

    def cat_z_vals(self, rays_o, rays_d, z_vals, new_z_vals, sdf, last=False):
        """renderer.py:191-205 (merge of two sorted lists; the new SDF values ride along)."""
        B, n = z_vals.shape
        with torch.no_grad():
            if last:
                z, _ = ops.merge_sorted(z_vals.contiguous(), new_z_vals.contiguous())
                return z, sdf
            pts = ops.ray_points(rays_o.contiguous(), rays_d.contiguous(), new_z_vals.contiguous())
            new_sdf = self._sdf_nograd(pts).reshape(B, -1)
            return ops.merge_sorted(z_vals.contiguous(), new_z_vals.contiguous(), sdf.reshape(B, n).contiguous(),
                                    new_sdf.contiguous())

    def _hierarchical(self, rays_o, rays_d, z_vals):
        """renderer.py:166-176: up_sample / cat_z_vals `up_sample_steps` times.  One launch per iteration between the SDF
        passes (ops.upsample_iter: the previous iteration's merge, this iteration's inverse-CDF depths and their positions);
        `up_sample` / `cat_z_vals` above stay as the reference's own entry points."""
        B = z_vals.shape[0]
        k = self.n_importance // self.up_sample_steps
        with torch.no_grad():
            u = self._linspace(0.5 / k, 1.0 - 0.5 / k, k, z_vals.device)
            pts = ops.ray_points(rays_o, rays_d, z_vals)
            sdf = self._sdf_nograd(pts).reshape(B, self.n_samples)
            new_z = new_sdf = None
            for i in range(self.up_sample_steps):
                last = i + 1 == self.up_sample_steps
                z_vals, sdf, new_z, pts = ops.upsample_iter(rays_o, rays_d, z_vals, sdf, new_z, new_sdf, k,
                                                            float(64 * 2 ** i), u, want_pts=not last)
                if not last:
                    new_sdf = self._sdf_nograd(pts).reshape(B, k)
            z_vals, _ = ops.merge_sorted(z_vals, new_z)           # the last new depths need no sdf (renderer.py:174-176)
        return z_vals

    # ------------------------------------------------------------------ core
    def render_core(self, rays_o, rays_d, z_vals, sample_dist, sdf_network, deviation_network, color_network,
                    refColor_network, background_alpha=None, background_sampled_color=None, background_rgb=None,
                    cos_anneal_ratio=0.0):
        """renderer.py:208-389."""
        B, n = z_vals.shape
        dev = z_vals.device
        rays_o, rays_d, z_vals = rays_o.contiguous(), rays_d.contiguous(), z_vals.contiguous()
        dists, mid_z, pts, dirs = ops.core_geometry(rays_o, rays_d, z_vals, sample_dist)

        w_sdf = getattr(self, "_w_sdf", None) if sdf_network is self.sdf_network else None
        # tensor-core path: `feature_vector` goes from the SDF chain to the colour chain as an operand image (never as an
        # FP32 [B*n,256] tensor in HBM); RefColor's two rows per ray are read out of the image
        img = bool(self.feature_image and getattr(sdf_network, "supports_feature_image", lambda: False)() and
                   getattr(color_network, "supports_feature_image", lambda: False)())
        if img:
            sdf, feat, normals = sdf_network.value_feature_normal(pts, want_normal=True, w=w_sdf, feat_image=True)
        else:
            sdf, feat, normals = sdf_network.value_feature_normal(pts, want_normal=True, w=w_sdf)
        if hasattr(deviation_network, "variance"):
            # SingleVarianceNetwork.forward (fields.py:267-268) on one row: ones * exp(10 variance), without the ones
            inv_s = ops.InvS.apply(deviation_network.variance)
        else:
            inv_s = deviation_network(torch.zeros([1, 3], device=dev))[:, :1].clip(1e-6, 1e6)   # [1,1]
        # `feat` feeds the colour network (all rows) and RefColor (2 rows per ray): see ops.FanOut
        stash = {"feat_image": img}
        fan = (self.fuse_feature_fanout or img) and feat.requires_grad
        feat_dense, feat_sparse = ops.FanOut.apply(feat, stash) if fan else (feat, feat)
        rgb = (color_network(pts, normals, dirs, feat_dense, feat_image=True) if img else
               color_network(pts, normals, dirs, feat_dense))                                     # [B*n,3]

        n_out = 0 if background_alpha is None else background_alpha.shape[1] - n
        color, weights, wsum, wmax, cdf, inside, eik_num, eik_den, hit_idx, w_pair, grad_err, hit = ops.Composite.apply(
            sdf, normals, rgb, inv_s, background_alpha, background_sampled_color, dists, pts, rays_d,
            background_rgb, n, n_out, cos_anneal_ratio if torch.is_tensor(cos_anneal_ratio) else float(cos_anneal_ratio))
        self.last_eikonal_parts = (eik_num, eik_den)     # for exact loss normalisation across ray shards

        # surface term (renderer.py:284-343) with fixed shapes: every ray evaluates RefColor at the two
        # bracketing samples, rays without a sign change are masked to the reference's default of ones.
        rows = ops.hit_rows(hit_idx, n)                                                          # [2B]
        r_rgb, r_spec, r_diff = self._ref_rows(refColor_network, pts, feat_sparse, dirs, normals, rows,
                                               stash if fan else None, img)
        surf_rgb, surf_spec, surf_diff = ops.SurfaceBlend.apply(r_rgb, r_spec, r_diff, w_pair, hit_idx)
        self.last_hit_idx = hit_idx
        self.last_weight_sum = wsum

        return {
            "color": color,
            "surface_color": surf_rgb,
            "sdf_mask": hit,
            "sdf": sdf,
            "dists": dists,
            "gradients": normals.reshape(B, n, 3),
            "s_val": (1.0 / inv_s).expand(B * n, 1),
            "mid_z_vals": mid_z,
            "weights": weights,
            "cdf": cdf,
            "gradient_error": grad_err,
            "inside_sphere": inside,
            "weight_sum": wsum,
            "weight_max": wmax,
            "specular_color": surf_spec,
            "diffuse_color": surf_diff,
        }

    @staticmethod
    def _ref_rows(net, pts, feat, dirs, normals, rows, stash=None, feat_image=False):
        if stash is not None:
            f_rows = ops.GatherRows.apply(feat, rows, stash)
        else:
            f_rows = ops.image_gather_rows(feat, rows) if feat_image else feat.index_select(0, rows)
        p_rows, d_rows, n_rows = ops.GatherRows3.apply(pts, dirs, normals, rows)
        d = net(p_rows, f_rows, d_rows, n_rows)
        return d["rgb"], d["specular_rgb"], d["diffuse_rgb"]

    # ------------------------------------------------------------------ render
    def render(self, rays_o, rays_d, near, far, perturb_overwrite=-1, background_rgb=None, cos_anneal_ratio=0.0):
        """renderer.py:391-500."""
        if not rays_o.is_cuda:
            raise RuntimeError("factored-neus_b200.NeuSRenderer.render needs CUDA tensors (no CPU fallback)")
        dev = rays_o.device
        rays_o, rays_d = rays_o.float().contiguous(), rays_d.float().contiguous()
        batch_size = len(rays_o)
        sample_dist = 2.0 / self.n_samples
        lin = self._linspace(0.0, 1.0, self.n_samples, dev)

        z_vals_outside = None
        if self.n_outside > 0:
            z_vals_outside = self._linspace(1e-3, 1.0 - 1.0 / (self.n_outside + 1.0), self.n_outside, dev)

        n_samples = self.n_samples
        perturb = self.perturb
        if perturb_overwrite >= 0:
            perturb = perturb_overwrite
        # z = near + (far - near) * lin (+ (rand - 0.5) * 2 / n_samples), one launch (renderer.py:395-408)
        rnd = torch.rand([batch_size, 1], device=dev) if perturb > 0 else None
        z_vals = ops.coarse_z(near, far, lin, rnd, self.n_samples)
        if perturb > 0:
            if self.n_outside > 0:
                mids = 0.5 * (z_vals_outside[..., 1:] + z_vals_outside[..., :-1])
                upper = torch.cat([mids, z_vals_outside[..., -1:]], -1)
                lower = torch.cat([z_vals_outside[..., :1], mids], -1)
                t_rand = torch.rand([batch_size, z_vals_outside.shape[-1]], device=dev)
                z_vals_outside = lower[None, :] + (upper - lower)[None, :] * t_rand
        if self.n_outside > 0:
            z_vals_outside = far / torch.flip(z_vals_outside, dims=[-1]) + 1.0 / self.n_samples

        background_alpha = None
        background_sampled_color = None
        z_vals = z_vals.contiguous()
        # the SDF weights (weight-norm + flat pack, differentiable) are packed ONCE for the five SDF passes of a render
        self._w_sdf = self.sdf_network.flat_weights()
        if self.n_importance > 0:
            z_vals = self._hierarchical(rays_o, rays_d, z_vals)
            n_samples = self.n_samples + self.n_importance

        if self.n_outside > 0:
            zo = z_vals_outside.expand(batch_size, self.n_outside).contiguous()
            z_vals_feed, _ = ops.merge_sorted(z_vals, zo)
            ret_outside = self.render_core_outside(rays_o, rays_d, z_vals_feed, sample_dist, self.nerf)
            background_sampled_color = ret_outside["sampled_color"]
            background_alpha = ret_outside["alpha"]

        ret_fine = self.render_core(rays_o, rays_d, z_vals, sample_dist, self.sdf_network, self.deviation_network,
                                    self.color_network, self.refColor_network, background_rgb=background_rgb,
                                    background_alpha=background_alpha,
                                    background_sampled_color=background_sampled_color,
                                    cos_anneal_ratio=cos_anneal_ratio)
        self._w_sdf = None
        weights = ret_fine["weights"]
        # (the mean over a ray of a constant: renderer.py:474 averages the [B*n,1] expansion of 1 / inv_s)
        s_val = ret_fine["s_val"][:1].expand(batch_size, 1)
        return {
            "color_fine": ret_fine["color"],
            "surface_color": ret_fine["surface_color"],
            "sdf_mask": ret_fine["sdf_mask"],
            "s_val": s_val,
            "cdf_fine": ret_fine["cdf"],
            "weight_sum": ret_fine["weight_sum"],
            "weight_max": ret_fine["weight_max"],
            "gradients": ret_fine["gradients"],
            "weights": weights,
            "gradient_error": ret_fine["gradient_error"],
            "inside_sphere": ret_fine["inside_sphere"],
            "specular_color": ret_fine["specular_color"],
            "diffuse_color": ret_fine["diffuse_color"],
        }

    # ------------------------------------------------------------------ stage 2: light visibility
    def lvis_mateIllu_render_util(self, rays_o, rays_d, near, far):
        """renderer.py:503-564: unperturbed up-sampling, SDF at the 128 section mid-points."""
        dev = rays_o.device
        rays_o, rays_d = rays_o.float().contiguous(), rays_d.float().contiguous()
        B = len(rays_o)
        sample_dist = 2.0 / self.n_samples
        z_vals = (near + (far - near) * self._linspace(0.0, 1.0, self.n_samples, dev)[None, :]).contiguous()
        n_samples = self.n_samples
        if self.n_importance > 0:
            z_vals = self._hierarchical(rays_o, rays_d, z_vals)
            n_samples = self.n_samples + self.n_importance
        dists, mid_z, pts, _ = ops.core_geometry(rays_o, rays_d, z_vals, sample_dist)
        with torch.no_grad():
            sdf = self._sdf_nograd(pts)
            # first sign change, inside-sphere test and secant root in one launch (renderer.py:553-556,588-602)
            hit_idx, z_surf, pts_surf, _, any_in = ops.first_hit_secant(sdf, mid_z, pts, rays_o, rays_d)
        return {"n_samples": n_samples, "mid_z_vals": mid_z, "sdf": sdf, "inside_sphere_mask": any_in,
                "hit_idx": hit_idx, "z_surf": z_surf, "pts_surf": pts_surf}

    def lvis_render(self, rays_o, rays_d, near, far, r_theta=None, rand_z=None):
        """renderer.py:567-627 (fixed shapes: rays without a surface hit are masked to the default of ones).
        ``r_theta`` / ``rand_z`` [B,4]: the two random draws of calLvis.py:351-352 (2 pi rand, 0.95 rand); drawn here when
        omitted."""
        from .lvis import cal_indiLgt
        B = len(rays_o)
        dev = rays_o.device
        util = self.lvis_mateIllu_render_util(rays_o, rays_d, near, far)
        sdf_mask = util["hit_idx"] >= 0
        pts_surf = util["pts_surf"]
        n_surf = self.sdf_network.gradient(pts_surf).reshape(-1, 3)
        res = cal_indiLgt(pts_surf, n_surf, self.sdf_network, self.deviation_network, self.color_network,
                          self.lvis_network, self.indiLgt_network, r_theta=r_theta, rand_z=rand_z)
        m1, m3 = sdf_mask[:, None], sdf_mask[:, None, None]
        ones1, ones3 = torch.ones(B, 4, device=dev), torch.ones(B, 4, 3, device=dev)
        return {"gt_lvis": torch.where(m1, res["gt_lvis"], ones1),
                "pre_lvis": torch.where(m1, res["pre_lvis"], ones1),
                "gt_trace_radiance": torch.where(m3, res["gt_trace_radiance"], ones3),
                "pre_trace_radiance": torch.where(m3, res["pre_trace_radiance"], ones3),
                "sdf_mask": sdf_mask}

    # ------------------------------------------------------------------ grid query / mesh
    def extract_fields(self, bound_min, bound_max, resolution, ix0=0, ix1=None):
        """renderer.py:14-29: u = -sdf on the regular grid, as a device tensor [ix1-ix0, R, R] (x-slab)."""
        net = self.sdf_network
        dev = next(net.parameters()).device
        axes = [torch.linspace(float(bound_min[a]), float(bound_max[a]), resolution, device=dev) for a in range(3)]
        with torch.no_grad():
            return ops.sdf_grid(net.cfg, net.flat_weights().detach(), axes[0], axes[1], axes[2], ix0, ix1)

    def extract_geometry(self, bound_min, bound_max, resolution, threshold=0.0):
        """renderer.py:32-40,729-734: SDF grid AND marching cubes on the GPU (the reference copies the 512^3 grid to the
        host and runs PyMCubes there); only the mesh comes back.  Returns (vertices [V,3] float, triangles [T,3] int) as
        numpy arrays like the reference, vertices rescaled to world coordinates with the reference's expression."""
        from .mcubes import marching_cubes
        u = self.extract_fields(bound_min, bound_max, resolution)
        vertices, triangles = marching_cubes(u, threshold)
        b_max = torch.as_tensor(bound_max).detach().to(vertices.device, torch.float32)
        b_min = torch.as_tensor(bound_min).detach().to(vertices.device, torch.float32)
        vertices = vertices / (resolution - 1.0) * (b_max - b_min)[None, :] + b_min[None, :]
        return vertices.cpu().numpy(), triangles.cpu().numpy()

    def render_core_outside(self, rays_o, rays_d, z_vals, sample_dist, nerf, background_rgb=None):
        """renderer.py:112-149."""
        B, n = z_vals.shape
        dists, pts4, dirs = ops.outside_geometry(rays_o.contiguous(), rays_d.contiguous(), z_vals.contiguous(),
                                                 sample_dist)
        if self.n_outside <= 0:
            pts4 = pts4[:, :3].contiguous()
        density, rgb_raw = nerf(pts4, dirs)
        alpha, color = ops.OutsideAlpha.apply(density.reshape(-1), rgb_raw, dists.reshape(-1))
        alpha = alpha.reshape(B, n)
        sampled_color = color.reshape(B, n, 3)
        weights = alpha * torch.cumprod(torch.cat([torch.ones([B, 1], device=alpha.device), 1.0 - alpha + 1e-7], -1),
                                        -1)[:, :-1]
        color_out = (weights[:, :, None] * sampled_color).sum(dim=1)
        if background_rgb is not None:
            color_out = color_out + background_rgb * (1.0 - weights.sum(dim=-1, keepdim=True))
        return {"color": color_out, "sampled_color": sampled_color, "alpha": alpha, "weights": weights}
