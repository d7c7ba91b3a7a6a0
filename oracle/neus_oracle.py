"""CPU oracle for the Factored-NeuS per-ray volume-rendering hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``factored-neus_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker or
the timed CPU baseline -- never as the product path.

This is a from-scratch *restatement* (functional style, plain torch CPU ops, any
float dtype) of the algorithms in the reference, written from the maths in
SURVEY.md Appendix A and the reference's behaviour.  Each function cites the
reference ``file:line`` it follows (paths relative to /root/reference).

Parity pinning: the reference has no tests/golden vectors of its own
(SURVEY.md section 4/8c), so the oracle is pinned against outputs of the
*imported reference itself*, generated in the build container by
``tools/make_golden.py`` and committed under ``tests/golden/``
(``tests/test_oracle_golden.py`` checks the oracle against them).

Parameters are plain dicts ``name -> tensor`` using the reference's
``state_dict`` key names (``lin0.weight_g``, ``lin0.weight_v``, ``lin0.bias``,
``variance``, ``pts_linears.0.weight`` ...), so reference checkpoints feed it.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

# ---------------------------------------------------------------------------
# configuration (confs/wmask.conf:49-97, confs/womask.conf)
# ---------------------------------------------------------------------------
SDF_CONF = dict(d_in=3, d_out=257, d_hidden=256, n_layers=8, skip_in=(4,), multires=6,
                bias=0.5, scale=1.0)
COLOR_CONF = dict(d_feature=256, d_in=9, d_out=3, d_hidden=256, n_layers=4, multires_view=4)
NERF_CONF = dict(D=8, W=256, d_in=4, d_in_view=3, multires=10, multires_view=4, skips=(4,))
RENDER_CONF_WMASK = dict(n_samples=64, n_importance=64, n_outside=0, up_sample_steps=4, perturb=1.0)
RENDER_CONF_WOMASK = dict(n_samples=64, n_importance=64, n_outside=32, up_sample_steps=4, perturb=1.0)


# ---------------------------------------------------------------------------
# positional encoding -- models/embedder.py:11-36
# ---------------------------------------------------------------------------
def embed(x: torch.Tensor, multires: int) -> torch.Tensor:
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)] (embedder.py:15-36)."""
    if multires <= 0:
        return x
    pieces = [x]
    for k in range(multires):
        f = float(2.0 ** k)  # freq_bands = 2**linspace(0, L-1, L) are exact powers of two
        pieces.append(torch.sin(x * f))
        pieces.append(torch.cos(x * f))
    return torch.cat(pieces, dim=-1)


def embed_dim(d: int, multires: int) -> int:
    return d * (1 + 2 * multires) if multires > 0 else d


# ---------------------------------------------------------------------------
# weight-norm'd linear layers -- fields.py:67-68,143-144 (torch.nn.utils.weight_norm)
# ---------------------------------------------------------------------------
def wn_weight(p: Params, name: str) -> torch.Tensor:
    """Effective weight g * v / ||v||_row of a weight-normalised Linear."""
    g, v = p[name + ".weight_g"], p[name + ".weight_v"]
    return v * (g / v.norm(2, dim=1, keepdim=True))


def _lin(p: Params, name: str, x: torch.Tensor, weight_norm: bool) -> torch.Tensor:
    w = wn_weight(p, name) if weight_norm else p[name + ".weight"]
    return F.linear(x, w, p[name + ".bias"])


# ---------------------------------------------------------------------------
# SDF network -- fields.py:74-111
# ---------------------------------------------------------------------------
def sdf_forward(p: Params, x: torch.Tensor, conf=SDF_CONF) -> torch.Tensor:
    """[N,3] -> [N,d_out]; col 0 = sdf, the rest = feature (fields.py:74-91)."""
    n_lin = conf["n_layers"] + 1
    scale = conf["scale"]
    h0 = embed(x * scale, conf["multires"])
    h = h0
    for l in range(n_lin):
        if l in conf["skip_in"]:
            h = torch.cat([h, h0], dim=1) / math.sqrt(2.0)
        h = _lin(p, "lin%d" % l, h, True)
        if l < n_lin - 1:
            h = F.softplus(h, beta=100)
    return torch.cat([h[:, :1] / scale, h[:, 1:]], dim=-1)


def sdf_value(p: Params, x: torch.Tensor, conf=SDF_CONF) -> torch.Tensor:
    """fields.py:93-95."""
    return sdf_forward(p, x, conf)[:, :1]


def sdf_gradient(p: Params, x: torch.Tensor, conf=SDF_CONF, create_graph=True) -> torch.Tensor:
    """d sdf / d x by autograd with create_graph (fields.py:100-111). Returns [N,3]."""
    with torch.enable_grad():
        xr = x.detach().requires_grad_(True)
        y = sdf_value(p, xr, conf)
        (g,) = torch.autograd.grad(y, xr, torch.ones_like(y), create_graph=create_graph,
                                   retain_graph=True)
    return g


def sdf_gradient_analytic(p: Params, x: torch.Tensor, conf=SDF_CONF):
    """Closed-form value + feature + normal (SURVEY.md A.1 reverse sweep); used to
    cross-check the autograd form and as the blueprint of the CUDA kernels."""
    n_lin = conf["n_layers"] + 1
    beta = 100.0
    scale = conf["scale"]
    xs = x * scale
    h0 = embed(xs, conf["multires"])
    h, sig = h0, []
    ws = [wn_weight(p, "lin%d" % l) for l in range(n_lin)]
    for l in range(n_lin):
        if l in conf["skip_in"]:
            h = torch.cat([h, h0], dim=1) / math.sqrt(2.0)
        a = F.linear(h, ws[l], p["lin%d.bias" % l])
        if l < n_lin - 1:
            sig.append(torch.where(a * beta > 20, torch.ones_like(a), torch.sigmoid(a * beta)))
            h = F.softplus(a, beta=beta)
        else:
            h = a
    out = torch.cat([h[:, :1] / scale, h[:, 1:]], dim=-1)
    g = ws[n_lin - 1][0:1, :].expand(x.shape[0], -1)
    g0_extra = None
    for l in range(n_lin - 2, -1, -1):
        q = sig[l] * g
        g = q @ ws[l]
        if l in conf["skip_in"]:
            k = ws[l].shape[1] - h0.shape[1]
            g0_extra = g[:, k:] / math.sqrt(2.0)
            g = g[:, :k] / math.sqrt(2.0)
    if g0_extra is not None:
        g = g + g0_extra
    d = x.shape[1]
    n = g[:, :d].clone()
    for k in range(conf["multires"]):
        f = float(2.0 ** k)
        gs = g[:, d * (1 + 2 * k): d * (2 + 2 * k)]
        gc = g[:, d * (2 + 2 * k): d * (3 + 2 * k)]
        n = n + f * torch.cos(xs * f) * gs - f * torch.sin(xs * f) * gc
    return out, n


# ---------------------------------------------------------------------------
# colour network -- fields.py:150-175 (mode 'idr')
# ---------------------------------------------------------------------------
def color_forward(p: Params, points, normals, view_dirs, feats, conf=COLOR_CONF) -> torch.Tensor:
    h = torch.cat([points, embed(view_dirs, conf["multires_view"]), normals, feats], dim=-1)
    n_lin = conf["n_layers"] + 1
    for l in range(n_lin):
        h = _lin(p, "lin%d" % l, h, True)
        if l < n_lin - 1:
            h = torch.relu(h)
    return torch.sigmoid(h)


# ---------------------------------------------------------------------------
# outside NeRF -- fields.py:233-259
# ---------------------------------------------------------------------------
def nerf_forward(p: Params, pts, views, conf=NERF_CONF):
    e_pts = embed(pts, conf["multires"])
    e_views = embed(views, conf["multires_view"])
    h = e_pts
    for i in range(conf["D"]):
        h = torch.relu(_lin(p, "pts_linears.%d" % i, h, False))
        if i in conf["skips"]:
            h = torch.cat([e_pts, h], dim=-1)
    density = _lin(p, "alpha_linear", h, False)
    feat = _lin(p, "feature_linear", h, False)
    h = torch.relu(_lin(p, "views_linears.0", torch.cat([feat, e_views], dim=-1), False))
    return density, _lin(p, "rgb_linear", h, False)


# ---------------------------------------------------------------------------
# variance -- fields.py:262-268 (+ clip at renderer.py:245)
# ---------------------------------------------------------------------------
def inv_s_of(variance: torch.Tensor) -> torch.Tensor:
    return torch.exp(variance * 10.0).clip(1e-6, 1e6)


# ---------------------------------------------------------------------------
# RefColor -- fields.py:303-335, math_utils.py:12-22,138-144
# ---------------------------------------------------------------------------
_EPS32 = float(torch.finfo(torch.float32).eps)


def linear_to_srgb(c: torch.Tensor) -> torch.Tensor:
    lo = 323.0 / 25.0 * c
    hi = (211.0 * torch.clamp_min(c, _EPS32) ** (5.0 / 12.0) - 11.0) / 200.0
    return torch.where(c <= 0.0031308, lo, hi)


def refcolor_forward(p: Params, pts, feats, dirs, n):
    """Returns (rgb, specular_rgb, diffuse_rgb), each [M,3]."""
    n_hat = n / torch.sqrt(torch.clamp_min((n * n).sum(-1, keepdim=True), _EPS32))
    wo = -dirs
    refl = 2.0 * (wo * n_hat).sum(-1, keepdim=True) * n_hat - wo
    h = torch.cat([pts, embed(n, 4), feats], dim=-1)
    for i in range(4):
        h = torch.relu(_lin(p, "net_cd.%d" % (2 * i), h, False))
    diffuse = torch.sigmoid(_lin(p, "net_cd.8", h, False))
    h = torch.cat([n, pts, embed(refl, 4), feats], dim=-1)
    for i in range(4):  # the i%4==0 and i>0 skip never fires for i<4 (fields.py:317)
        h = torch.relu(_lin(p, "viewdir_mlp.%d" % i, h, False))
    spec = torch.sigmoid(_lin(p, "net_cs.0", h, False)).repeat(1, 3)
    rgb = torch.clip(linear_to_srgb(spec + diffuse), 0.0, 1.0)
    return rgb, torch.clip(linear_to_srgb(spec), 0.0, 1.0), torch.clip(linear_to_srgb(diffuse), 0.0, 1.0)


# ---------------------------------------------------------------------------
# sampling -- renderer.py:43-77 (== calLvis.py:25-52), :152-189, :191-205
# ---------------------------------------------------------------------------
def pdf_to_cdf(weights: torch.Tensor) -> torch.Tensor:
    """renderer.py:52-55: cdf = [0, cumsum((w+1e-5)/sum(w+1e-5))]."""
    w = weights + 1e-5
    pdf = w / w.sum(-1, keepdim=True)
    c = torch.cumsum(pdf, -1)
    return torch.cat([torch.zeros_like(c[..., :1]), c], -1)


def invert_cdf(bins, cdf, u):
    """renderer.py:64-77. bins,cdf [B,n]; u [B,k]. Returns (samples [B,k], inds int64 [B,k])."""
    inds = torch.searchsorted(cdf, u.contiguous(), right=True)
    lo = (inds - 1).clamp_min(0)
    hi = inds.clamp_max(cdf.shape[-1] - 1)
    c_lo, c_hi = cdf.gather(-1, lo), cdf.gather(-1, hi)
    b_lo, b_hi = bins.gather(-1, lo), bins.gather(-1, hi)
    den = c_hi - c_lo
    den = torch.where(den < 1e-5, torch.ones_like(den), den)
    t = (u - c_lo) / den
    return b_lo + t * (b_hi - b_lo), inds


def sample_pdf_det(bins, weights, k):
    u = torch.linspace(0.5 / k, 1.0 - 0.5 / k, steps=k, dtype=bins.dtype)
    u = u.expand(bins.shape[0], k)
    return invert_cdf(bins, pdf_to_cdf(weights), u)[0]


def upsample_weights(rays_o, rays_d, z, sdf, inv_s):
    """renderer.py:158-186 -> per-interval weights [B,n-1]."""
    B, n = z.shape
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z[..., None]
    r = torch.linalg.norm(pts, ord=2, dim=-1)
    inside = (r[:, :-1] < 1.0) | (r[:, 1:] < 1.0)
    sdf = sdf.reshape(B, n)
    f0, f1 = sdf[:, :-1], sdf[:, 1:]
    z0, z1 = z[:, :-1], z[:, 1:]
    mid = (f0 + f1) * 0.5
    cos = (f1 - f0) / (z1 - z0 + 1e-5)
    prev = torch.cat([torch.zeros(B, 1, dtype=z.dtype), cos[:, :-1]], dim=-1)
    cos = torch.minimum(prev, cos).clip(-1e3, 0.0) * inside
    dz = z1 - z0
    c_prev = torch.sigmoid((mid - cos * dz * 0.5) * inv_s)
    c_next = torch.sigmoid((mid + cos * dz * 0.5) * inv_s)
    alpha = (c_prev - c_next + 1e-5) / (c_prev + 1e-5)
    trans = torch.cumprod(torch.cat([torch.ones(B, 1, dtype=z.dtype), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
    return alpha * trans


def up_sample(rays_o, rays_d, z, sdf, k, inv_s):
    """renderer.py:152-189 (and calLvis.py:55-90): k new depths per ray, no grad."""
    with torch.no_grad():
        return sample_pdf_det(z, upsample_weights(rays_o, rays_d, z, sdf, inv_s), k)


def cat_z_vals(sdf_fn, rays_o, rays_d, z, new_z, sdf, last=False):
    """renderer.py:191-205."""
    B, n = z.shape
    k = new_z.shape[1]
    zz, index = torch.sort(torch.cat([z, new_z], dim=-1), dim=-1)
    if not last:
        pts = rays_o[:, None, :] + rays_d[:, None, :] * new_z[..., None]
        new_sdf = sdf_fn(pts.reshape(-1, 3)).reshape(B, k)
        sdf = torch.cat([sdf, new_sdf], dim=-1).gather(-1, index)
    return zz, sdf


# ---------------------------------------------------------------------------
# render_core_outside -- renderer.py:112-149
# ---------------------------------------------------------------------------
def render_core_outside(nerf_p: Params, rays_o, rays_d, z, sample_dist, conf=NERF_CONF):
    B, n = z.shape
    dists = torch.cat([z[:, 1:] - z[:, :-1], torch.full((B, 1), sample_dist, dtype=z.dtype)], -1)
    mid = z + dists * 0.5
    pts = rays_o[:, None, :] + rays_d[:, None, :] * mid[..., None]
    r = torch.linalg.norm(pts, ord=2, dim=-1, keepdim=True).clip(1.0, 1e10)
    pts4 = torch.cat([pts / r, 1.0 / r], dim=-1).reshape(-1, 4)
    dirs = rays_d[:, None, :].expand(B, n, 3).reshape(-1, 3)
    density, rgb = nerf_forward(nerf_p, pts4, dirs, conf)
    alpha = 1.0 - torch.exp(-F.softplus(density.reshape(B, n)) * dists)
    return dict(alpha=alpha, sampled_color=torch.sigmoid(rgb).reshape(B, n, 3))


# ---------------------------------------------------------------------------
# render_core -- renderer.py:208-389
# ---------------------------------------------------------------------------
def first_hit(sdf_bn, inside):
    """renderer.py:290-292: first index with sdf<0 (needs idx>=1 and any sample inside)."""
    n = sdf_bn.shape[1]
    ramp = torch.arange(n, 0, -1, dtype=sdf_bn.dtype).reshape(1, n)
    val, idx = torch.min(torch.sign(sdf_bn) * ramp, dim=-1)
    mask = (val < 0.0) & (idx >= 1) & (inside.sum(-1) > 0.0)
    return mask, idx


def neus_alpha(sdf, grads, dirs, dists, inv_s, cos_anneal_ratio):
    """renderer.py:248-268: sdf [N,1], grads/dirs [N,3], dists [B,n] -> (alpha [B,n], prev_cdf [N,1])."""
    B, n = dists.shape
    true_cos = (dirs * grads).sum(-1, keepdim=True)
    iter_cos = -(torch.relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_anneal_ratio)
                 + torch.relu(-true_cos) * cos_anneal_ratio)
    d1 = dists.reshape(-1, 1)
    c_prev = torch.sigmoid((sdf - iter_cos * d1 * 0.5) * inv_s)
    c_next = torch.sigmoid((sdf + iter_cos * d1 * 0.5) * inv_s)
    alpha = ((c_prev - c_next + 1e-5) / (c_prev + 1e-5)).reshape(B, n).clip(0.0, 1.0)
    return alpha, c_prev


def transmittance_weights(alpha):
    """renderer.py:360: w = alpha * exclusive_cumprod(1 - alpha + 1e-7)."""
    B = alpha.shape[0]
    return alpha * torch.cumprod(torch.cat([torch.ones(B, 1, dtype=alpha.dtype), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]


def render_core(P: Dict[str, Params], rays_o, rays_d, z, sample_dist, background_alpha=None,
                background_sampled_color=None, background_rgb=None, cos_anneal_ratio=0.0,
                sdf_conf=SDF_CONF, color_conf=COLOR_CONF):
    """P = {'sdf':..., 'var':..., 'color':..., 'ref':...}. Follows renderer.py:208-389."""
    B, n = z.shape
    dt = z.dtype
    dists = torch.cat([z[:, 1:] - z[:, :-1], torch.full((B, 1), sample_dist, dtype=dt)], -1)
    mid_z = z + dists * 0.5
    pts = (rays_o[:, None, :] + rays_d[:, None, :] * mid_z[..., None]).reshape(-1, 3)
    dirs = rays_d[:, None, :].expand(B, n, 3).reshape(-1, 3)

    out = sdf_forward(P["sdf"], pts, sdf_conf)
    sdf, feat = out[:, :1], out[:, 1:]
    grads = sdf_gradient(P["sdf"], pts, sdf_conf)
    inv_s = inv_s_of(P["var"]["variance"]).reshape(1, 1)

    alpha, c_prev = neus_alpha(sdf, grads, dirs, dists, inv_s, cos_anneal_ratio)

    r = torch.linalg.norm(pts, ord=2, dim=-1).reshape(B, n)
    inside = (r < 1.0).to(dt)
    relax = (r < 1.2).to(dt)

    rgb = color_forward(P["color"], pts, grads, dirs, feat, color_conf).reshape(B, n, 3)

    spec_c = torch.ones(B, 3, dtype=dt)
    diff_c = torch.ones(B, 3, dtype=dt)
    surf_c = torch.ones(B, 3, dtype=dt)
    sdf_bn = sdf.reshape(B, n)
    hit, idx = first_hit(sdf_bn, inside)
    if int(hit.sum()) > 0:
        ii = idx[hit]
        pair = torch.stack([ii - 1, ii], dim=1)                      # [m,2]

        def take(t):                                                  # t [B,n,C] -> [2m,C]
            th = t[hit]
            return th.gather(1, pair[..., None].expand(-1, -1, th.shape[-1])).reshape(-1, th.shape[-1])

        r_rgb, r_spec, r_diff = refcolor_forward(
            P["ref"], take(pts.reshape(B, n, 3)), take(feat.reshape(B, n, -1)),
            take(dirs.reshape(B, n, 3)), take(grads.reshape(B, n, 3)))
        a_in = alpha * inside
        w_in = a_in * torch.cumprod(torch.cat([torch.ones(B, 1, dtype=dt), 1.0 - a_in + 1e-7], -1), -1)[:, :-1]
        w2 = w_in[hit].gather(1, pair) + 1e-5                         # [m,2]
        den = w2.sum(1, keepdim=True)

        def blend(c):
            c = c.reshape(-1, 2, 3)
            return (c[:, 0] * w2[:, :1] + c[:, 1] * w2[:, 1:]) / den

        spec_c[hit], diff_c[hit], surf_c[hit] = blend(r_spec), blend(r_diff), blend(r_rgb)

    if background_alpha is not None:
        alpha = alpha * inside + background_alpha[:, :n] * (1.0 - inside)
        alpha = torch.cat([alpha, background_alpha[:, n:]], dim=-1)
        rgb = rgb * inside[:, :, None] + background_sampled_color[:, :n] * (1.0 - inside)[:, :, None]
        rgb = torch.cat([rgb, background_sampled_color[:, n:]], dim=1)

    weights = alpha * torch.cumprod(torch.cat([torch.ones(B, 1, dtype=dt), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
    wsum = weights.sum(-1, keepdim=True)
    color = (rgb * weights[:, :, None]).sum(1)
    if background_rgb is not None:
        color = color + background_rgb * (1.0 - wsum)
    g3 = grads.reshape(B, n, 3)
    eik = (torch.linalg.norm(g3, ord=2, dim=-1) - 1.0) ** 2
    eik_num, eik_den = (relax * eik).sum(), relax.sum()
    eik = eik_num / (eik_den + 1e-5)
    return dict(color=color, surface_color=surf_c, sdf_mask=hit, sdf=sdf, dists=dists, gradients=g3,
                s_val=(1.0 / inv_s).expand(B * n, 1), mid_z_vals=mid_z, weights=weights,
                cdf=c_prev.reshape(B, n), gradient_error=eik, inside_sphere=inside,
                specular_color=spec_c, diffuse_color=diff_c, eik_num=eik_num, eik_den=eik_den)


# ---------------------------------------------------------------------------
# render -- renderer.py:391-500
# ---------------------------------------------------------------------------
def coarse_z(near, far, n_samples, t_rand=None):
    z = near + (far - near) * torch.linspace(0.0, 1.0, n_samples, dtype=near.dtype)[None, :]
    if t_rand is not None:
        z = z + t_rand * 2.0 / n_samples
    return z


def outside_z(far, n_outside, n_samples, t_rand=None):
    zo = torch.linspace(1e-3, 1.0 - 1.0 / (n_outside + 1.0), n_outside, dtype=far.dtype)
    if t_rand is not None:
        mids = 0.5 * (zo[1:] + zo[:-1])
        upper = torch.cat([mids, zo[-1:]], -1)
        lower = torch.cat([zo[:1], mids], -1)
        zo = lower[None, :] + (upper - lower)[None, :] * t_rand
    return far / torch.flip(zo, dims=[-1]) + 1.0 / n_samples


def hierarchical_z(sdf_p, rays_o, rays_d, z, n_importance, up_sample_steps, sdf_conf=SDF_CONF):
    """renderer.py:425-447; returns (z [B,n+n_importance], list of per-step new z)."""
    B, n = z.shape
    fn = lambda q: sdf_value(sdf_p, q, sdf_conf)
    steps = []
    with torch.no_grad():
        pts = rays_o[:, None, :] + rays_d[:, None, :] * z[..., None]
        sdf = fn(pts.reshape(-1, 3)).reshape(B, n)
        for i in range(up_sample_steps):
            nz = up_sample(rays_o, rays_d, z, sdf, n_importance // up_sample_steps, 64 * 2 ** i)
            steps.append(nz)
            z, sdf = cat_z_vals(fn, rays_o, rays_d, z, nz, sdf, last=(i + 1 == up_sample_steps))
    return z, steps


def render(P, rays_o, rays_d, near, far, conf=RENDER_CONF_WMASK, perturb_overwrite=-1,
           background_rgb=None, cos_anneal_ratio=0.0, t_rand=None, t_rand_outside=None,
           sdf_conf=SDF_CONF, color_conf=COLOR_CONF, nerf_conf=NERF_CONF, z_override=None):
    """renderer.py:391-500.  RNG is externalised: t_rand [B,1] in [-0.5,0.5) and
    t_rand_outside [B,n_outside] in [0,1) are drawn by the caller when perturbing."""
    B = rays_o.shape[0]
    ns, ni, no = conf["n_samples"], conf["n_importance"], conf["n_outside"]
    sample_dist = 2.0 / ns
    perturb = conf["perturb"] if perturb_overwrite < 0 else perturb_overwrite
    if perturb > 0:
        if t_rand is None:
            t_rand = torch.rand(B, 1, dtype=near.dtype) - 0.5
        if no > 0 and t_rand_outside is None:
            t_rand_outside = torch.rand(B, no, dtype=near.dtype)
    else:
        t_rand, t_rand_outside = None, None
    z = coarse_z(near, far, ns, t_rand)
    z_out = outside_z(far, no, ns, t_rand_outside) if no > 0 else None
    n = ns
    if ni > 0:
        if z_override is None:
            z, _ = hierarchical_z(P["sdf"], rays_o, rays_d, z, ni, conf["up_sample_steps"], sdf_conf)
        else:  # test hook: inverse-CDF sampling is ill-conditioned in empty bins, so per-sample
            z = z_override  # comparisons are made on identical depths
        n = ns + ni
    bg_a = bg_c = None
    if no > 0:
        z_feed, _ = torch.sort(torch.cat([z, z_out], dim=-1), dim=-1)
        o = render_core_outside(P["nerf"], rays_o, rays_d, z_feed, sample_dist, nerf_conf)
        bg_a, bg_c = o["alpha"], o["sampled_color"]
    r = render_core(P, rays_o, rays_d, z, sample_dist, bg_a, bg_c, background_rgb, cos_anneal_ratio,
                    sdf_conf, color_conf)
    w = r["weights"]
    return dict(color_fine=r["color"], surface_color=r["surface_color"], sdf_mask=r["sdf_mask"],
                s_val=r["s_val"].reshape(B, n).mean(-1, keepdim=True), cdf_fine=r["cdf"],
                weight_sum=w.sum(-1, keepdim=True), weight_max=w.max(-1, keepdim=True)[0],
                gradients=r["gradients"], weights=w, gradient_error=r["gradient_error"],
                inside_sphere=r["inside_sphere"], specular_color=r["specular_color"],
                diffuse_color=r["diffuse_color"], z_vals=z, eik_num=r["eik_num"], eik_den=r["eik_den"])


# ---------------------------------------------------------------------------
# stage-1 loss -- exp_runner.py:134-177
# ---------------------------------------------------------------------------
def stage1_loss(out, true_rgb, mask, surface_weight=0.1, igr_weight=0.1, mask_weight=0.1):
    mask = (mask > 0.5).to(true_rgb.dtype) if mask_weight > 0.0 else torch.ones_like(mask)
    mask_sum = mask.sum() + 1e-5
    hit = out["sdf_mask"]
    mask_sdf_sum = mask[hit].sum() + 1e-5
    color_loss = ((out["color_fine"] - true_rgb) * mask).abs().sum() / mask_sum
    surf_err = surface_weight * (out["surface_color"][hit] - true_rgb[hit]) * mask[hit]
    surf_loss = surf_err.abs().sum() / mask_sdf_sum
    mask_loss = F.binary_cross_entropy(out["weight_sum"].clip(1e-3, 1.0 - 1e-3), mask)
    loss = color_loss + surf_loss + out["gradient_error"] * igr_weight + mask_loss * mask_weight
    return loss, dict(color_loss=color_loss, surface_loss=surf_loss, mask_loss=mask_loss)


# ---------------------------------------------------------------------------
# grid query -- renderer.py:14-29
# ---------------------------------------------------------------------------
def extract_fields(sdf_p, bound_min, bound_max, resolution, sdf_conf=SDF_CONF, chunk=64):
    """u[x,y,z] = -sdf(grid point); grid = ij-meshgrid of per-axis linspace (renderer.py:16-28)."""
    dt = bound_min.dtype
    axes = [torch.linspace(float(bound_min[a]), float(bound_max[a]), resolution, dtype=dt) for a in range(3)]
    u = torch.zeros(resolution, resolution, resolution, dtype=dt)
    with torch.no_grad():
        for x0 in range(0, resolution, chunk):
            for y0 in range(0, resolution, chunk):
                for z0 in range(0, resolution, chunk):
                    xs, ys, zs = axes[0][x0:x0 + chunk], axes[1][y0:y0 + chunk], axes[2][z0:z0 + chunk]
                    g = torch.stack(torch.meshgrid(xs, ys, zs, indexing="ij"), dim=-1).reshape(-1, 3)
                    u[x0:x0 + len(xs), y0:y0 + len(ys), z0:z0 + len(zs)] = \
                        -sdf_value(sdf_p, g, sdf_conf).reshape(len(xs), len(ys), len(zs))
    return u


# ---------------------------------------------------------------------------
# stage-2 light visibility -- calLvis.py:9-204,302-409, renderer.py:503-627
# ---------------------------------------------------------------------------
def sample_dirs(normals, r_theta, r_phi):
    """calLvis.py:302-320.  normals [m,3]; r_theta, r_phi [m,k] -> [m,k,3]."""
    tiny = 1e-6
    nrm = lambda v: v / (torch.norm(v, dim=-1, keepdim=True) + tiny)
    n = nrm(normals)[:, None, :]
    ex = torch.zeros_like(n)
    ex[..., 0] = 1
    u = nrm(torch.cross(ex, n, dim=-1))
    v = nrm(torch.cross(n, u, dim=-1))
    th, ph = r_theta[..., None], r_phi[..., None]
    return u * torch.cos(th) * torch.sin(ph) + v * torch.sin(th) * torch.sin(ph) + n * torch.cos(ph)


def _secondary_geometry(rays_o, rays_d, z):
    B, n = z.shape
    sd = (1 - 0.1) / 32.0                                              # calLvis.py:95,155
    dists = torch.cat([z[:, 1:] - z[:, :-1], torch.full((B, 1), sd, dtype=z.dtype)], -1)
    mid = z + dists * 0.5
    pts = (rays_o[:, None, :] + rays_d[:, None, :] * mid[..., None]).reshape(-1, 3)
    return dists, mid, pts


def first_hit_rgb(P, rays_o, rays_d, z, sdf_conf=SDF_CONF, color_conf=COLOR_CONF):
    """calLvis.py:153-204."""
    B, n = z.shape
    dists, mid, pts = _secondary_geometry(rays_o, rays_d, z)
    sdf = sdf_forward(P["sdf"], pts, sdf_conf)[:, :1].reshape(B, n)
    inside = (torch.linalg.norm(pts, dim=-1).reshape(B, n) < 1.0).to(z.dtype)
    hit, idx = first_hit(sdf, inside)
    rgb = torch.zeros(B, 3, dtype=z.dtype)
    if int(hit.sum()) > 0:
        ii = idx[hit].reshape(-1, 1)
        z_lo, z_hi = mid[hit].gather(1, ii - 1), mid[hit].gather(1, ii)
        s_lo, s_hi = sdf[hit].gather(1, ii - 1), sdf[hit].gather(1, ii)
        z_s = (s_lo * z_hi - s_hi * z_lo) / (s_lo - s_hi + 1e-10)
        p_s = rays_o[hit] + rays_d[hit] * z_s
        n_s = sdf_gradient(P["sdf"], p_s, sdf_conf)
        f_s = sdf_forward(P["sdf"], p_s, sdf_conf)[:, 1:]
        rgb[hit] = color_forward(P["color"], p_s, n_s, rays_d[hit], f_s, color_conf)
    return rgb, hit


def occlusion_weights(P, rays_o, rays_d, z, sdf_conf=SDF_CONF):
    """calLvis.py:93-150 (cos_anneal_ratio fixed at 0, everything detached)."""
    B, n = z.shape
    dists, mid, pts = _secondary_geometry(rays_o, rays_d, z)
    dirs = rays_d[:, None, :].expand(B, n, 3).reshape(-1, 3)
    sdf = sdf_forward(P["sdf"], pts, sdf_conf)[:, :1].detach()
    inv_s = inv_s_of(P["var"]["variance"]).detach().reshape(1, 1)
    g = sdf_gradient(P["sdf"], pts, sdf_conf).detach()
    tc = (dirs * g).sum(-1, keepdim=True)
    ic = -torch.relu(-tc * 0.5 + 0.5)
    d1 = dists.reshape(-1, 1)
    c_prev = torch.sigmoid((sdf - ic * d1 * 0.5) * inv_s)
    c_next = torch.sigmoid((sdf + ic * d1 * 0.5) * inv_s)
    alpha = ((c_prev - c_next + 1e-5) / (c_prev + 1e-5)).reshape(B, n).clip(0.0, 1.0)
    inside = (torch.linalg.norm(pts, dim=-1).reshape(B, n) < 1.0).to(z.dtype)
    w = alpha * torch.cumprod(torch.cat([torch.ones(B, 1, dtype=z.dtype), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
    return w, w * inside


def trace_visibility(P, surf, normal, r_theta, rand_z, n_coarse=512, n_imp=32,
                     sdf_conf=SDF_CONF, color_conf=COLOR_CONF):
    """gt part of calLvis.py:339-397 with RNG externalised (r_theta = 2*pi*rand, rand_z = 0.95*rand).
    Returns gt_lvis [m,k], gt_trace_radiance [m,k,3], dirs [m,k,3], z_fine [m*k,n_imp]."""
    m, k = r_theta.shape
    dirs = sample_dirs(normal, r_theta, torch.asin(rand_z))
    o = surf[:, None, :].repeat(1, k, 1).reshape(-1, 3)
    d = dirs.reshape(-1, 3)
    with torch.no_grad():
        zc = torch.linspace(0.0, 1.0, n_coarse, dtype=surf.dtype)[None, :].expand(o.shape[0], -1)
        pc = (o[:, None, :] + d[:, None, :] * zc[:, :, None]).reshape(-1, 3)
        sdf_c = sdf_forward(P["sdf"], pc, sdf_conf)[:, :1]
    inv_s = inv_s_of(P["var"]["variance"]).detach().reshape(())
    z_fine = up_sample(o, d, zc, sdf_c, n_imp, inv_s)
    rgb, _ = first_hit_rgb(P, o, d, z_fine, sdf_conf, color_conf)
    _, w_in = occlusion_weights(P, o, d, z_fine, sdf_conf)
    lvis = 1 - w_in.detach().sum(-1)
    return lvis.reshape(m, k).detach(), rgb.reshape(m, k, 3).detach(), dirs, z_fine


def surface_points(P, rays_o, rays_d, near, far, conf=RENDER_CONF_WMASK, sdf_conf=SDF_CONF):
    """renderer.py:503-605: unperturbed up-sampling, sign-change, secant root, normal."""
    B = rays_o.shape[0]
    ns, ni = conf["n_samples"], conf["n_importance"]
    z = coarse_z(near, far, ns)
    z, _ = hierarchical_z(P["sdf"], rays_o, rays_d, z, ni, conf["up_sample_steps"], sdf_conf)
    n = ns + ni
    dists = torch.cat([z[:, 1:] - z[:, :-1], torch.full((B, 1), 2.0 / ns, dtype=z.dtype)], -1)
    mid = z + dists * 0.5
    pts = (rays_o[:, None, :] + rays_d[:, None, :] * mid[..., None]).reshape(-1, 3)
    sdf = sdf_forward(P["sdf"], pts, sdf_conf)[:, :1].reshape(B, n)
    inside = (torch.linalg.norm(pts, dim=-1).reshape(B, n) < 1.0).to(z.dtype)
    hit, idx = first_hit(sdf, inside)
    ii = idx.clamp_min(1).reshape(-1, 1)
    z_lo, z_hi = mid.gather(1, ii - 1), mid.gather(1, ii)
    s_lo, s_hi = sdf.gather(1, ii - 1), sdf.gather(1, ii)
    z_s = (s_lo * z_hi - s_hi * z_lo) / (s_lo - s_hi + 1e-10)
    p_s = rays_o + rays_d * z_s
    n_s = sdf_gradient(P["sdf"], p_s, sdf_conf)
    return hit, p_s, n_s


def query_indir_illum(lgtSGs, dirs):
    """calLvis.py:323-336: 24 spherical Gaussians [n,24,7] evaluated along [n,k,3] directions -> [n,k,3]."""
    k, nl = dirs.shape[1], lgtSGs.shape[1]
    sg = lgtSGs.unsqueeze(-3).expand(-1, k, -1, -1)
    d = dirs.unsqueeze(-2).expand(-1, -1, nl, -1)
    lobes = sg[..., :3] / torch.norm(sg[..., :3], dim=-1, keepdim=True)
    return (sg[..., -3:] * torch.exp(sg[..., 3:4] * (torch.sum(d * lobes, dim=-1, keepdim=True) - 1.0))).sum(dim=2)


def lvis_render(P, p_lvis: Params, p_indi: Params, rays_o, rays_d, near, far, r_theta, rand_z,
                conf=RENDER_CONF_WMASK, sdf_conf=SDF_CONF, color_conf=COLOR_CONF):
    """renderer.py:567-627 with fixed shapes: every ray runs the trace; rays without a surface hit keep the
    reference's defaults of ones.  r_theta / rand_z [B,4]: the two draws of calLvis.py:351-352 (rows of rays without a
    hit are ignored)."""
    B = rays_o.shape[0]
    hit, p_s, n_s = surface_points(P, rays_o, rays_d, near, far, conf, sdf_conf)
    n_s = n_s.detach()
    p_s = p_s.detach()
    gt_lvis, gt_rad, dirs, _ = trace_visibility(P, p_s, n_s, r_theta, rand_z, sdf_conf=sdf_conf, color_conf=color_conf)
    k = r_theta.shape[1]
    o = p_s[:, None, :].repeat(1, k, 1).reshape(-1, 3)
    pre_lvis = lvis_forward(p_lvis, o, dirs.reshape(-1, 3)).reshape(B, k)
    pre_rad = query_indir_illum(indirect_light_forward(p_indi, p_s), dirs)
    m1, m3 = hit[:, None], hit[:, None, None]
    one1, one3 = torch.ones(B, k, dtype=rays_o.dtype), torch.ones(B, k, 3, dtype=rays_o.dtype)
    return dict(gt_lvis=torch.where(m1, gt_lvis, one1), pre_lvis=torch.where(m1, pre_lvis, one1),
                gt_trace_radiance=torch.where(m3, gt_rad, one3), pre_trace_radiance=torch.where(m3, pre_rad, one3),
                sdf_mask=hit)


def stage2_loss(out):
    """lvis.py:163-170: lvis_loss = sum |gt_lvis - pre_lvis| / (4 n_hit + 1e-6) (unmasked difference: rows without a
    hit are ones on both sides), radiance_loss = sum |(gt - pre) * mask| / (12 n_hit + 1e-6)."""
    hit = out["sdf_mask"]
    dt = out["gt_lvis"].dtype
    lvis_err = out["gt_lvis"] - out["pre_lvis"]
    lvis_loss = lvis_err.abs().sum() / (hit[..., None].expand(out["gt_lvis"].shape).sum().to(dt) + 1e-6)
    rad_err = (out["gt_trace_radiance"] - out["pre_trace_radiance"]) * hit[..., None, None].to(dt)
    rad_loss = rad_err.abs().sum() / (hit[..., None, None].expand(out["gt_trace_radiance"].shape).sum().to(dt) + 1e-6)
    return lvis_loss + rad_loss, dict(lvis_loss=lvis_loss, radiance_loss=rad_loss)


# ---------------------------------------------------------------------------
# stage-2 prediction networks -- fields.py:338-413
# ---------------------------------------------------------------------------
def lvis_forward(p: Params, pts, view):
    h = torch.cat([embed(pts, 10), embed(view, 4)], dim=-1)
    for i in range(4):
        h = torch.relu(_lin(p, "lvis.%d" % (2 * i), h, False))
    return torch.sigmoid(_lin(p, "lvis.8", h, False))


def indirect_light_forward(p: Params, pts, num_lgt_sgs=24):
    h = embed(pts, 10)
    for i in range(4):
        h = torch.relu(_lin(p, "indi.%d" % (2 * i), h, False))
    out = _lin(p, "indi.8", h, False).reshape(-1, num_lgt_sgs, 6)
    lobes = torch.sigmoid(out[..., :2])
    theta, phi = lobes[..., :1] * 2 * math.pi, lobes[..., 1:2] * 2 * math.pi
    axis = torch.cat([torch.cos(theta) * torch.sin(phi), torch.sin(theta) * torch.sin(phi), torch.cos(phi)], dim=-1)
    sharp = torch.sigmoid(out[..., 2:3]) * 30 + 0.1
    amp = torch.relu(out[..., 3:])
    return torch.cat([axis, sharp, amp], dim=-1)
