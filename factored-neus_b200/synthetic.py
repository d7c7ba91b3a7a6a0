"""Deterministic synthetic weights, rays and targets.

The reference ships no data and no checkpoints, so benchmarks and parity tests use
the "sphere-SDF scene": every network at (geometric) random init, rays aimed at
the unit sphere from distance 2.5.  Everything is drawn from ``numpy.random.
RandomState`` (bit-stable across numpy versions and independent of torch's RNG
stream), so the golden fixtures under ``tests/golden`` can be regenerated
anywhere.  Tensors use the reference's ``state_dict`` key names
(fields.py:67-70,139-146,214-231,280-300) and load into either implementation.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

SDF_CONF = dict(d_in=3, d_out=257, d_hidden=256, n_layers=8, skip_in=(4,), multires=6,
                bias=0.5, scale=1.0, geometric_init=True, weight_norm=True)
COLOR_CONF = dict(d_feature=256, mode="idr", d_in=9, d_out=3, d_hidden=256, n_layers=4,
                  weight_norm=True, multires_view=4, squeeze_out=True)
NERF_CONF = dict(D=8, W=256, d_in=4, d_in_view=3, multires=10, multires_view=4, output_ch=4,
                 skips=[4], use_viewdirs=True)
RENDER_CONF_WMASK = dict(n_samples=64, n_importance=64, n_outside=0, up_sample_steps=4, perturb=1.0)
RENDER_CONF_WOMASK = dict(n_samples=64, n_importance=64, n_outside=32, up_sample_steps=4, perturb=1.0)


def _t(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


def _default_linear(rs, out_d, in_d):
    """nn.Linear-style U(-1/sqrt(in), 1/sqrt(in)) for weight and bias."""
    k = 1.0 / math.sqrt(in_d)
    return rs.uniform(-k, k, (out_d, in_d)), rs.uniform(-k, k, (out_d,))


def _wn(sd, name, w, b, rs, jitter, dtype):
    """Store (w, b) as weight-norm parameters; g = ||v|| * (1 + jitter*N(0,1))."""
    g = np.linalg.norm(w, axis=1, keepdims=True)
    if jitter > 0:
        g = g * (1.0 + jitter * rs.standard_normal(g.shape))
    sd[name + ".weight_g"] = _t(g, dtype)
    sd[name + ".weight_v"] = _t(w, dtype)
    sd[name + ".bias"] = _t(b, dtype)


def sdf_state(seed=0, conf=SDF_CONF, jitter=0.0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Geometric init of fields.py:47-65; ``jitter`` adds N(0, jitter*sigma) to every
    weight/bias (including the structurally-zero PE columns) so parity tests exercise
    every term, like a partly trained network."""
    rs = np.random.RandomState(seed)
    d_in, d_h, nl, d_out = conf["d_in"], conf["d_hidden"], conf["n_layers"], conf["d_out"]
    e = d_in * (1 + 2 * conf["multires"]) if conf["multires"] > 0 else d_in
    dims = [e] + [d_h] * nl + [d_out]
    sd = {}
    for l in range(len(dims) - 1):
        out_d = dims[l + 1] - dims[0] if (l + 1) in conf["skip_in"] else dims[l + 1]
        in_d = dims[l]
        sig = math.sqrt(2.0) / math.sqrt(out_d)
        if l == len(dims) - 2:
            w = math.sqrt(math.pi) / math.sqrt(in_d) + 1e-4 * rs.standard_normal((out_d, in_d))
            b = np.full((out_d,), -conf["bias"])
            sig = 1e-4 if jitter == 0 else 0.02
        elif conf["multires"] > 0 and l == 0:
            w = np.zeros((out_d, in_d))
            w[:, :3] = sig * rs.standard_normal((out_d, 3))
            b = np.zeros((out_d,))
        elif conf["multires"] > 0 and l in conf["skip_in"]:
            w = sig * rs.standard_normal((out_d, in_d))
            w[:, -(dims[0] - 3):] = 0.0
            b = np.zeros((out_d,))
        else:
            w = sig * rs.standard_normal((out_d, in_d))
            b = np.zeros((out_d,))
        if jitter > 0:
            w = w + jitter * sig * rs.standard_normal(w.shape)
            b = b + jitter * sig * rs.standard_normal(b.shape)
        _wn(sd, "lin%d" % l, w, b, rs, jitter, dtype)
    return sd


def color_state(seed=1, conf=COLOR_CONF, jitter=0.0, dtype=torch.float32):
    rs = np.random.RandomState(seed)
    e_v = 3 * (1 + 2 * conf["multires_view"]) if conf["multires_view"] > 0 else 3
    dims = [conf["d_in"] + conf["d_feature"] + (e_v - 3)] + [conf["d_hidden"]] * conf["n_layers"] + [conf["d_out"]]
    sd = {}
    for l in range(len(dims) - 1):
        w, b = _default_linear(rs, dims[l + 1], dims[l])
        _wn(sd, "lin%d" % l, w, b, rs, jitter, dtype)
    return sd


def variance_state(init_val=0.3, dtype=torch.float32):
    return {"variance": torch.tensor(init_val, dtype=dtype)}


def nerf_state(seed=2, conf=NERF_CONF, dtype=torch.float32):
    rs = np.random.RandomState(seed)
    W, D = conf["W"], conf["D"]
    e_p = conf["d_in"] * (1 + 2 * conf["multires"])
    e_v = conf["d_in_view"] * (1 + 2 * conf["multires_view"])
    sd = {}

    def put(name, out_d, in_d):
        w, b = _default_linear(rs, out_d, in_d)
        sd[name + ".weight"], sd[name + ".bias"] = _t(w, dtype), _t(b, dtype)

    put("pts_linears.0", W, e_p)
    for i in range(D - 1):
        put("pts_linears.%d" % (i + 1), W, W + e_p if i in conf["skips"] else W)
    put("views_linears.0", W // 2, e_v + W)
    put("feature_linear", W, W)
    put("alpha_linear", 1, W)
    put("rgb_linear", 3, W // 2)
    return sd


def refcolor_state(seed=3, d_feature=256, d_hidden=256, dtype=torch.float32):
    """RefColor (fields.py:280-300) with its Lazy layers materialised: net_cd in = 3+27+F,
    viewdir_mlp.0 in = 3+3+27+F, net_cs.0 in = d_hidden."""
    rs = np.random.RandomState(seed)
    sd = {}

    def put(name, out_d, in_d):
        w, b = _default_linear(rs, out_d, in_d)
        sd[name + ".weight"], sd[name + ".bias"] = _t(w, dtype), _t(b, dtype)

    put("net_cd.0", d_hidden, 30 + d_feature)
    for i in (2, 4, 6):
        put("net_cd.%d" % i, d_hidden, d_hidden)
    put("net_cd.8", 3, d_hidden)
    put("viewdir_mlp.0", d_hidden, 33 + d_feature)
    for i in (1, 2, 3):
        put("viewdir_mlp.%d" % i, d_hidden, d_hidden)
    put("net_cs.0", 1, d_hidden)
    return sd


def _relu_stack_state(prefix, seed, d_in, d_hidden, d_out, dtype):
    rs = np.random.RandomState(seed)
    sd = {}
    dims = [d_in, d_hidden, d_hidden, d_hidden, d_hidden, d_out]
    for i in range(5):
        w, b = _default_linear(rs, dims[i + 1], dims[i])
        sd["%s.%d.weight" % (prefix, 2 * i)], sd["%s.%d.bias" % (prefix, 2 * i)] = _t(w, dtype), _t(b, dtype)
    return sd


def lvis_state(seed=5, dtype=torch.float32):
    """Lvis (fields.py:338-369), Lazy first layer materialised: in = 63 + 27."""
    return _relu_stack_state("lvis", seed, 90, 256, 1, dtype)


def indirect_light_state(seed=6, num_lgt_sgs=24, dtype=torch.float32):
    """IndirectLight (fields.py:372-413), Lazy first layer materialised: in = 63."""
    return _relu_stack_state("indi", seed, 63, 512, num_lgt_sgs * 6, dtype)


def scene_states(seed=4, jitter=0.0, dtype=torch.float32, sdf_conf=SDF_CONF, color_conf=COLOR_CONF,
                 nerf_conf=NERF_CONF):
    """All networks of one synthetic scene: {'sdf','var','color','ref','nerf'} -> state dict."""
    return {
        "sdf": sdf_state(seed, sdf_conf, jitter, dtype),
        "color": color_state(seed + 1, color_conf, jitter, dtype),
        "var": variance_state(0.3, dtype),
        "nerf": nerf_state(seed + 2, nerf_conf, dtype),
        "ref": refcolor_state(seed + 3, color_conf["d_feature"], color_conf["d_hidden"], dtype),
    }


def make_rays(B, seed=1, dtype=torch.float32):
    """o = 2.5*normalize(randn); d = normalize((rand-0.5)*0.8 - o); near/far = mid -/+ 1
    (dataset.py:186-192)."""
    rs = np.random.RandomState(seed)
    o = rs.standard_normal((B, 3))
    o = 2.5 * o / np.linalg.norm(o, axis=1, keepdims=True)
    tgt = (rs.uniform(0, 1, (B, 3)) - 0.5) * 0.8
    d = tgt - o
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    o32, d32 = _t(o, dtype), _t(d, dtype)
    a = (d32 * d32).sum(-1, keepdim=True)
    b = 2.0 * (o32 * d32).sum(-1, keepdim=True)
    mid = 0.5 * (-b) / a
    return o32, d32, mid - 1.0, mid + 1.0


def make_targets(B, seed=2, dtype=torch.float32):
    rs = np.random.RandomState(seed)
    return _t(rs.uniform(0, 1, (B, 3)), dtype), torch.ones(B, 1, dtype=dtype)


def pinhole_camera(H=1200, W=1600, focal=2900.0, distance=2.5, dtype=torch.float32):
    """Synthetic DTU-like camera: principal point at the image centre, pose at ``distance`` on the +z
    axis looking at the origin.  Returns (intrinsics_inv [4,4], pose [4,4]) in the layout of Dataset.intrinsics_all_inv /
    pose_all (dataset.py:60-100)."""
    K = np.eye(4)
    K[0, 0] = K[1, 1] = focal
    K[0, 2], K[1, 2] = (W - 1) * 0.5, (H - 1) * 0.5
    pose = np.eye(4)
    pose[:3, :3] = np.diag([1.0, -1.0, -1.0])          # camera looks down -z of the world, y flipped (image rows go down)
    pose[:3, 3] = [0.0, 0.0, distance]
    return _t(np.linalg.inv(K), dtype), _t(pose, dtype)


def image_pixels(H, W, resolution_level=1, device="cpu"):
    """Pixel coordinates of Dataset.gen_rays_at (dataset.py:115-131) in its OUTPUT order ([H//l, W//l] row-major):
    px, py [H//l * W//l]."""
    l = resolution_level
    tx = torch.linspace(0, W - 1, W // l, device=device)
    ty = torch.linspace(0, H - 1, H // l, device=device)
    py, px = torch.meshgrid(ty, tx, indexing="ij")
    return px.reshape(-1).contiguous(), py.reshape(-1).contiguous()
