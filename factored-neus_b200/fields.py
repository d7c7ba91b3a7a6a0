"""Drop-in ``nn.Module`` fields with the reference's constructor signatures and ``state_dict`` keys
(models/fields.py), whose forward/backward run in libfneus_b200.so.

Parameters keep the reference layout (``lin{l}.weight_g [out,1]``, ``lin{l}.weight_v [out,in]``,
``lin{l}.bias``; ``variance``; ``net_cd.{0,2,4,6,8}.*`` ...), so reference checkpoints load unchanged
and Adam sees the same parameter list.  Kernels consume flat packs of EFFECTIVE weights
``W = g * v / ||v||_row`` assembled by differentiable torch ops (tiny).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import ops


class WNLinear(nn.Module):
    """Parameters of ``nn.utils.weight_norm(nn.Linear)`` (fields.py:67-68,143-144): weight_g, weight_v, bias."""

    def __init__(self, linear: nn.Linear):
        super().__init__()
        w = linear.weight.detach()
        self.weight_g = nn.Parameter(w.norm(2, dim=1, keepdim=True).clone())
        self.weight_v = nn.Parameter(w.clone())
        self.bias = nn.Parameter(linear.bias.detach().clone())
        self.in_features, self.out_features = linear.in_features, linear.out_features

    def effective(self):
        v = self.weight_v
        return v * (self.weight_g / v.norm(2, dim=1, keepdim=True))


class PlainLinear(nn.Module):
    def __init__(self, linear: nn.Linear):
        super().__init__()
        self.weight = nn.Parameter(linear.weight.detach().clone())
        self.bias = nn.Parameter(linear.bias.detach().clone())
        self.in_features, self.out_features = linear.in_features, linear.out_features

    def effective(self):
        return self.weight


def _flat_pack(layers):
    """Flat pack [W0, b0, W1, b1, ...] of effective weights: one fused CUDA launch (weight-norm included) on the
    GPU; plain torch ops for parameters that still live on the CPU (no kernels are involved there)."""
    first = layers[0].bias
    if first.is_cuda:
        return ops.pack_weights(layers)
    parts = []
    for lin in layers:
        w = lin.effective() if hasattr(lin, "effective") else lin.weight
        parts.append(w.reshape(-1))
        parts.append(lin.bias.reshape(-1))
    return torch.cat(parts)


class SDFNetwork(nn.Module):
    """models/fields.py:9-111."""

    def __init__(self, d_in, d_out, d_hidden, n_layers, skip_in=(4,), multires=0, bias=0.5, scale=1,
                 geometric_init=True, weight_norm=True, inside_outside=False):
        super().__init__()
        if d_in != 3:
            raise ValueError("fneus SDFNetwork: d_in must be 3")
        skip_in = tuple(skip_in)
        if len(skip_in) > 1:
            raise ValueError("fneus SDFNetwork: at most one skip connection is supported")
        dims = [d_in] + [d_hidden for _ in range(n_layers)] + [d_out]
        if multires > 0:
            dims[0] = d_in * (1 + 2 * multires)
        self.num_layers = len(dims)
        self.skip_in = skip_in
        self.scale = scale
        self.multires = multires
        self.weight_norm = weight_norm
        for l in range(self.num_layers - 1):
            out_dim = dims[l + 1] - dims[0] if (l + 1) in skip_in else dims[l + 1]
            lin = nn.Linear(dims[l], out_dim)
            if geometric_init:                                    # fields.py:47-65 (same RNG call order)
                if l == self.num_layers - 2:
                    sign = -1.0 if inside_outside else 1.0
                    nn.init.normal_(lin.weight, mean=sign * np.sqrt(np.pi) / np.sqrt(dims[l]), std=0.0001)
                    nn.init.constant_(lin.bias, -sign * bias)
                elif multires > 0 and l == 0:
                    nn.init.constant_(lin.bias, 0.0)
                    nn.init.constant_(lin.weight[:, 3:], 0.0)
                    nn.init.normal_(lin.weight[:, :3], 0.0, np.sqrt(2) / np.sqrt(out_dim))
                elif multires > 0 and l in skip_in:
                    nn.init.constant_(lin.bias, 0.0)
                    nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(out_dim))
                    nn.init.constant_(lin.weight[:, -(dims[0] - 3):], 0.0)
                else:
                    nn.init.constant_(lin.bias, 0.0)
                    nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(out_dim))
            setattr(self, "lin" + str(l), WNLinear(lin) if weight_norm else PlainLinear(lin))
        self.cfg = L.SdfCfg(d_in=d_in, d_hidden=d_hidden, n_layers=n_layers, d_out=d_out, multires=multires,
                            skip_layer=(skip_in[0] if skip_in else -1), scale=float(scale), beta=100.0)
        ops.register_module(self)
        self.d_out = d_out

    def flat_weights(self):
        return _flat_pack([getattr(self, "lin" + str(l)) for l in range(self.num_layers - 1)])

    def value_feature_normal(self, x, want_normal=True, w=None, feat_image=False):
        """One fused pass: (sdf [N,1], feature [N,d_out-1], d sdf/dx [N,3]); all differentiable w.r.t. weights.
        ``w``: an already packed ``flat_weights()`` tensor of THIS network (the renderer packs once per render).
        ``feat_image``: hand the features over as an operand image (an opaque tensor only RenderingNetwork's
        ``feat_image=True`` and ops.GatherRows can read; see supports_feature_image)."""
        return ops.SdfValueGrad.apply(self.flat_weights() if w is None else w, x, self.cfg, want_normal, feat_image)

    def supports_feature_image(self):
        return bool(L.lib().fneus_sdf_feat_image_ok(self.cfg))

    def forward(self, inputs, iter_step=0):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            sdf, feat, _ = self.value_feature_normal(inputs, want_normal=False)
        else:
            sdf, feat = ops.sdf_forward_nograd(self.cfg, self.flat_weights(), inputs, want_feat=True)
        return torch.cat([sdf, feat], dim=-1)

    def sdf(self, x):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return self.value_feature_normal(x, want_normal=False)[0]
        return ops.sdf_forward_nograd(self.cfg, self.flat_weights(), x, want_feat=False)[0]

    def sdf_hidden_appearance(self, x):
        return self.forward(x)

    def gradient(self, x):
        """[N,1,3] analytic d sdf/dx (fields.py:100-111), differentiable w.r.t. the weights."""
        with torch.enable_grad():
            return self.value_feature_normal(x, want_normal=True)[2].unsqueeze(1)


class RenderingNetwork(nn.Module):
    """models/fields.py:114-175 (mode 'idr', squeeze_out)."""

    def __init__(self, d_feature, mode, d_in, d_out, d_hidden, n_layers, weight_norm=True, multires_view=0,
                 squeeze_out=True):
        super().__init__()
        if mode != "idr" or not squeeze_out or d_in != 9:
            raise ValueError("fneus RenderingNetwork: only mode='idr', d_in=9, squeeze_out=True is implemented")
        self.mode, self.squeeze_out = mode, squeeze_out
        dims = [d_in + d_feature] + [d_hidden for _ in range(n_layers)] + [d_out]
        if multires_view > 0:
            dims[0] += 3 * (1 + 2 * multires_view) - 3
        self.num_layers = len(dims)
        for l in range(self.num_layers - 1):
            lin = nn.Linear(dims[l], dims[l + 1])
            setattr(self, "lin" + str(l), WNLinear(lin) if weight_norm else PlainLinear(lin))
        self.cfg = L.ColorCfg(d_feature=d_feature, d_hidden=d_hidden, n_layers=n_layers, d_out=d_out,
                              multires_view=multires_view)
        ops.register_module(self)

    def flat_weights(self):
        return _flat_pack([getattr(self, "lin" + str(l)) for l in range(self.num_layers - 1)])

    def forward(self, points, normals, view_dirs, feature_vectors, feat_image=False):
        return ops.ColorMLP.apply(self.flat_weights(), points, normals, view_dirs, feature_vectors, self.cfg,
                                  feat_image)

    def supports_feature_image(self):
        return bool(L.lib().fneus_color_feat_image_ok(self.cfg))


class NeRF(nn.Module):
    """models/fields.py:178-259 (use_viewdirs=True)."""

    def __init__(self, D=8, W=256, d_in=3, d_in_view=3, multires=0, multires_view=0, output_ch=4, skips=[4],
                 use_viewdirs=False):
        super().__init__()
        if not use_viewdirs:
            raise ValueError("fneus NeRF: only use_viewdirs=True is implemented (the reference asserts it too)")
        if len(skips) > 1:
            raise ValueError("fneus NeRF: at most one skip")
        self.D, self.W, self.d_in, self.d_in_view = D, W, d_in, d_in_view
        self.input_ch = d_in * (1 + 2 * multires) if multires > 0 else d_in
        self.input_ch_view = d_in_view * (1 + 2 * multires_view) if multires_view > 0 else d_in_view
        self.skips, self.use_viewdirs = skips, use_viewdirs
        self.pts_linears = nn.ModuleList(
            [nn.Linear(self.input_ch, W)] +
            [nn.Linear(W, W) if i not in skips else nn.Linear(W + self.input_ch, W) for i in range(D - 1)])
        self.views_linears = nn.ModuleList([nn.Linear(self.input_ch_view + W, W // 2)])
        self.feature_linear = nn.Linear(W, W)
        self.alpha_linear = nn.Linear(W, 1)
        self.rgb_linear = nn.Linear(W // 2, 3)
        self.cfg = L.NerfCfg(D=D, W=W, d_in=d_in, d_in_view=d_in_view, multires=multires,
                             multires_view=multires_view, skip=(skips[0] if skips else -1))
        ops.register_module(self)

    def flat_weights(self):
        if self.rgb_linear.bias.is_cuda:
            params, layout = [], []

            def add(t):
                r, c = (t.shape if t.dim() == 2 else (1, t.numel()))
                layout.append(("copy", -1, len(params), r, c))
                params.append(t)

            for lin in self.pts_linears:
                add(lin.weight); add(lin.bias)
            add(self.alpha_linear.weight); add(self.feature_linear.weight)
            add(self.alpha_linear.bias); add(self.feature_linear.bias)
            for lin in (self.views_linears[0], self.rgb_linear):
                add(lin.weight); add(lin.bias)
            return ops.PackWeights.apply(layout, *params)
        parts = []
        for lin in self.pts_linears:
            parts += [lin.weight.reshape(-1), lin.bias.reshape(-1)]
        parts += [self.alpha_linear.weight.reshape(-1), self.feature_linear.weight.reshape(-1),
                  self.alpha_linear.bias.reshape(-1), self.feature_linear.bias.reshape(-1)]
        for lin in (self.views_linears[0], self.rgb_linear):
            parts += [lin.weight.reshape(-1), lin.bias.reshape(-1)]
        return torch.cat(parts)

    def forward(self, input_pts, input_views):
        return ops.NerfMLP.apply(self.flat_weights(), input_pts, input_views, self.cfg)


class SingleVarianceNetwork(nn.Module):
    """models/fields.py:262-268."""

    def __init__(self, init_val):
        super().__init__()
        self.register_parameter("variance", nn.Parameter(torch.tensor(init_val)))

    def forward(self, x):
        return torch.ones([len(x), 1], device=self.variance.device) * torch.exp(self.variance * 10.0)


class RefColor(nn.Module):
    """models/fields.py:271-335.  The reference's Lazy layers are materialised at construction
    (in-features 30+F, 33+F, H), so checkpoints load without a dummy forward."""

    def __init__(self, d_feature=256, d_hidden=256):
        super().__init__()
        H, Fd = d_hidden, d_feature
        self.net_cd = nn.Sequential(nn.Linear(30 + Fd, H), nn.ReLU(), nn.Linear(H, H), nn.ReLU(), nn.Linear(H, H),
                                    nn.ReLU(), nn.Linear(H, H), nn.ReLU(), nn.Linear(H, 3), nn.Sigmoid())
        self.viewdir_mlp = nn.ModuleList([nn.Linear(33 + Fd if i == 0 else H, H) for i in range(4)])
        self.net_cs = nn.Sequential(nn.Linear(H, 1), nn.Sigmoid())
        self.cfg = L.RefCfg(d_feature=Fd, d_hidden=H)
        ops.register_module(self)

    def flat_weights(self):
        layers = [self.net_cd[i] for i in (0, 2, 4, 6, 8)] + list(self.viewdir_mlp) + [self.net_cs[0]]
        return _flat_pack(layers)

    def forward(self, pts, x, dirs, n):
        rgb, spec, diff = ops.RefColorMLP.apply(self.flat_weights(), pts, x, dirs, n, self.cfg)
        return {"rgb": rgb, "specular_rgb": spec, "diffuse_rgb": diff}


class Lvis(nn.Module):
    """models/fields.py:338-369: predicted light visibility, [PE10(pts), PE4(view)] -> 256x4 ReLU -> 1 sigmoid."""

    def __init__(self):
        super().__init__()
        self.lvis = nn.Sequential(nn.Linear(90, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU(), nn.Linear(256, 256),
                                  nn.ReLU(), nn.Linear(256, 256), nn.ReLU(), nn.Linear(256, 1), nn.Sigmoid())
        self.cfg = L.MlpCfg(n_inputs=2, in_dim=(3, 3), in_multires=(10, 4), d_hidden=256, n_layers=4, d_out=1,
                            last_act=1)
        ops.register_module(self)

    def flat_weights(self):
        return _flat_pack([self.lvis[i] for i in (0, 2, 4, 6, 8)])

    def forward(self, pts, view):
        return ops.PlainMLP.apply(self.flat_weights(), pts, view, self.cfg)


class IndirectLight(nn.Module):
    """models/fields.py:372-413: PE10(pts) -> 512x4 ReLU -> 144 = 24 spherical Gaussians x 6, then the lobe /
    sharpness / amplitude parametrisation (tiny, torch)."""

    def __init__(self, num_lgt_sgs=24):
        super().__init__()
        self.num_lgt_sgs = num_lgt_sgs
        self.indi = nn.Sequential(nn.Linear(63, 512), nn.ReLU(), nn.Linear(512, 512), nn.ReLU(), nn.Linear(512, 512),
                                  nn.ReLU(), nn.Linear(512, 512), nn.ReLU(), nn.Linear(512, num_lgt_sgs * 6))
        self.cfg = L.MlpCfg(n_inputs=1, in_dim=(3, 0), in_multires=(10, 0), d_hidden=512, n_layers=4,
                            d_out=num_lgt_sgs * 6, last_act=0)
        ops.register_module(self)

    def flat_weights(self):
        return _flat_pack([self.indi[i] for i in (0, 2, 4, 6, 8)])

    def forward(self, pts):
        output = ops.PlainMLP.apply(self.flat_weights(), pts, None, self.cfg).reshape(-1, self.num_lgt_sgs, 6)
        lobes = torch.sigmoid(output[..., :2])
        theta, phi = lobes[..., :1] * 2 * np.pi, lobes[..., 1:2] * 2 * np.pi
        lgt_lobes = torch.cat([torch.cos(theta) * torch.sin(phi), torch.sin(theta) * torch.sin(phi), torch.cos(phi)],
                              dim=-1)
        sharp = torch.sigmoid(output[..., 2:3]) * 30 + 0.1
        amp = torch.relu(output[..., 3:])
        return torch.cat([lgt_lobes, sharp, amp], dim=-1)
