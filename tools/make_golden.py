#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference with the four shims of SURVEY.md section 8c) on the deterministic
synthetic weights/rays of ``factored-neus_b200/synthetic.py``.

Only runs in the build container (the GPU box has no /root/reference); the
fixtures it writes are committed.  Usage:  python tools/make_golden.py
"""
import math
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("FNEUS_REFERENCE", "/root/reference")


def import_reference():
    for name in ("mcubes", "icecream", "imageio"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.ic = lambda *a, **k: None
            sys.modules[name] = m
    if not hasattr(np, "math"):
        np.math = math
    torch.Tensor.cuda = lambda self, *a, **k: self          # calLvis.py:305,351-352 hard-code .cuda()
    sys.path.insert(0, REF)
    from models import fields, renderer, calLvis            # noqa
    return fields, renderer, calLvis


def build_reference(fields, renderer, states, sdf_conf, color_conf, nerf_conf, render_conf):
    sdf = fields.SDFNetwork(**sdf_conf)
    col = fields.RenderingNetwork(**color_conf)
    var = fields.SingleVarianceNetwork(0.3)
    nerf = fields.NeRF(**nerf_conf)
    ref = fields.RefColor()
    F_ = color_conf["d_feature"]
    ref(torch.zeros(2, 3), torch.zeros(2, F_), torch.ones(2, 3), torch.ones(2, 3))   # materialise Lazy layers
    sdf.load_state_dict(states["sdf"]); col.load_state_dict(states["color"])
    var.load_state_dict(states["var"]); nerf.load_state_dict(states["nerf"]); ref.load_state_dict(states["ref"])
    r = renderer.NeuSRenderer(**render_conf, nerf=nerf, sdf_network=sdf, deviation_network=var,
                              color_network=col, refColor_network=ref)
    return dict(sdf=sdf, color=col, var=var, nerf=nerf, ref=ref, renderer=r)


def grad_digest(named_params, n_probe=48):
    """Per-parameter digest of .grad: sum, abs-sum, l2 and n_probe strided entries."""
    out = {}
    for name, p in named_params:
        g = p.grad
        g = torch.zeros_like(p) if g is None else g
        flat = g.detach().reshape(-1).double()
        idx = torch.linspace(0, flat.numel() - 1, min(n_probe, flat.numel())).long()
        out[name] = np.concatenate([[flat.sum().item(), flat.abs().sum().item(), flat.norm().item()],
                                    flat[idx].numpy()])
    return out


def stage2_networks(fields, syn, gdir):
    """Lvis / IndirectLight (fields.py:338-413): forward values and weight-gradient digests of a probe loss."""
    np_ = lambda t: t.detach().cpu().numpy()
    rs = np.random.RandomState(13)
    pts = torch.from_numpy(rs.uniform(-1, 1, (80, 3)).astype(np.float32))
    view = torch.from_numpy(rs.standard_normal((80, 3)).astype(np.float32))
    view = view / view.norm(dim=-1, keepdim=True)
    probe_v = torch.from_numpy(rs.standard_normal((80, 1)).astype(np.float32))
    probe_s = torch.from_numpy(rs.standard_normal((80, 24, 7)).astype(np.float32))
    lv, il = fields.Lvis(), fields.IndirectLight()
    lv(pts, view); il(pts)                                       # materialise the Lazy layers
    lv.load_state_dict(syn.lvis_state()); il.load_state_dict(syn.indirect_light_state())
    vis = lv(pts, view)
    (vis * probe_v).sum().backward()
    sgs = il(pts)
    (sgs * probe_s).sum().backward()
    save = dict(pts=np_(pts), view=np_(view), probe_v=np_(probe_v), probe_s=np_(probe_s), vis=np_(vis), sgs=np_(sgs))
    for k, val in grad_digest(lv.named_parameters()).items():
        save["g.lvis." + k] = val
    for k, val in grad_digest(il.named_parameters()).items():
        save["g.indi." + k] = val
    np.savez_compressed(os.path.join(gdir, "stage2_nets.npz"), **save)


def main():
    syn = __import__("factored_neus_b200").synthetic
    fields, renderer, calLvis = import_reference()
    gdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gdir, exist_ok=True)
    np_ = lambda t: t.detach().cpu().numpy()
    stage2_networks(fields, syn, gdir)
    if "--only-stage2" in sys.argv:
        return
    if "--only-lvis-render" in sys.argv:
        lvis_render_golden(fields, renderer, syn, syn.scene_states(seed=4, jitter=0.03), gdir)
        return

    # ---------------- networks, per-function (full-size wmask shapes, jittered weights) -------------
    states = syn.scene_states(seed=4, jitter=0.03)
    mods = build_reference(fields, renderer, states, syn.SDF_CONF, syn.COLOR_CONF, syn.NERF_CONF,
                           syn.RENDER_CONF_WMASK)
    rs = np.random.RandomState(7)
    x = torch.from_numpy(rs.uniform(-1, 1, (96, 3)).astype(np.float32))
    v = torch.from_numpy(rs.standard_normal((96, 3)).astype(np.float32))
    v = v / v.norm(dim=-1, keepdim=True)
    out = mods["sdf"](x)
    grad = mods["sdf"].gradient(x.clone()).squeeze(1)
    rgb = mods["color"](x, grad, v, out[:, 1:])
    x4 = torch.cat([x, torch.from_numpy(rs.uniform(0.1, 1, (96, 1)).astype(np.float32))], -1)
    dens, nrgb = mods["nerf"](x4, v)
    rd = mods["ref"](x, out[:, 1:], v, grad)
    np.savez_compressed(os.path.join(gdir, "fields.npz"), x=np_(x), v=np_(v), sdf_out=np_(out), grad=np_(grad),
                        rgb=np_(rgb), x4=np_(x4), nerf_density=np_(dens), nerf_rgb=np_(nrgb),
                        ref_rgb=np_(rd["rgb"]), ref_spec=np_(rd["specular_rgb"]), ref_diff=np_(rd["diffuse_rgb"]))

    # ---------------- sampling chain ---------------------------------------------------------------
    B = 24
    o, d, near, far = syn.make_rays(B, seed=1)
    R = mods["renderer"]
    with torch.no_grad():
        z = near + (far - near) * torch.linspace(0.0, 1.0, 64)[None, :]
        pts = o[:, None, :] + d[:, None, :] * z[..., None]
        sdf = mods["sdf"].sdf(pts.reshape(-1, 3)).reshape(B, 64)
        samp = dict(z0=np_(z), sdf0=np_(sdf))
        for i in range(4):
            # intermediate cdf/inds: same code path as sample_pdf, captured for the bit-exact index test
            w = _upsample_weights_via_reference(R, o, d, z, sdf, 64 * 2 ** i)
            wp = w + 1e-5
            cdf = torch.cumsum(wp / wp.sum(-1, keepdim=True), -1)
            cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
            u = torch.linspace(0.5 / 16, 1 - 0.5 / 16, 16).expand(B, 16).contiguous()
            inds = torch.searchsorted(cdf, u, right=True)
            new_z = R.up_sample(o, d, z, sdf, 16, 64 * 2 ** i)
            samp.update({"w%d" % i: np_(w), "cdf%d" % i: np_(cdf), "inds%d" % i: np_(inds), "u": np_(u[0]),
                         "newz%d" % i: np_(new_z)})
            z, sdf = R.cat_z_vals(o, d, z, new_z, sdf, last=(i == 3))
            samp.update({"z%d" % (i + 1): np_(z), "sdf%d" % (i + 1): np_(sdf)})
    np.savez_compressed(os.path.join(gdir, "sampling.npz"), o=np_(o), d=np_(d), near=np_(near), far=np_(far), **samp)

    # ---------------- full render fwd+bwd, wmask and womask ------------------------------------------
    for tag, rconf, car, mw in (("wmask", syn.RENDER_CONF_WMASK, 1.0, 0.1), ("womask", syn.RENDER_CONF_WOMASK, 0.3, 0.0)):
        mods = build_reference(fields, renderer, states, syn.SDF_CONF, syn.COLOR_CONF, syn.NERF_CONF, rconf)
        true_rgb, mask = syn.make_targets(B, seed=2)
        zbox = {}
        core = mods["renderer"].render_core

        def core_spy(ro_, rd_, z_, *a, **k):
            zbox["z"] = z_.detach().clone()
            return core(ro_, rd_, z_, *a, **k)

        mods["renderer"].render_core = core_spy
        ro = mods["renderer"].render(o, d, near, far, perturb_overwrite=0, cos_anneal_ratio=car)
        # stage-1 loss, exp_runner.py:134-177
        m = (mask > 0.5).float() if mw > 0 else torch.ones_like(mask)
        msum = m.sum() + 1e-5
        hit = ro["sdf_mask"]
        msdf = m[hit].sum() + 1e-5
        cl = ((ro["color_fine"] - true_rgb) * m).abs().sum() / msum
        se = 0.1 * (ro["surface_color"][hit] - true_rgb[hit]) * m[hit]
        sl = se.abs().sum() / msdf
        ml = torch.nn.functional.binary_cross_entropy(ro["weight_sum"].clip(1e-3, 1 - 1e-3), m)
        loss = cl + sl + ro["gradient_error"] * 0.1 + ml * mw
        loss.backward()
        save = {k: np_(t.float() if t.dtype == torch.bool else t) for k, t in ro.items()}
        save["loss"] = np.array(loss.item())
        save["z_vals"] = np_(zbox["z"])
        for net in ("sdf", "color", "var", "ref", "nerf"):
            for k, val in grad_digest(mods[net].named_parameters()).items():
                save["grad.%s.%s" % (net, k)] = val
        np.savez_compressed(os.path.join(gdir, "render_%s.npz" % tag), **save)
        print(tag, "loss", loss.item(), "hits", int(hit.sum()))

    # ---------------- grid query (renderer.py:14-29) ------------------------------------------------
    mods = build_reference(fields, renderer, states, syn.SDF_CONF, syn.COLOR_CONF, syn.NERF_CONF, syn.RENDER_CONF_WMASK)
    bmin, bmax = torch.tensor([-1.01] * 3), torch.tensor([1.01] * 3)
    u = renderer.extract_fields(bmin, bmax, 20, lambda p: -mods["sdf"].sdf(p))
    np.savez_compressed(os.path.join(gdir, "grid.npz"), u=u, bmin=np_(bmin), bmax=np_(bmax))

    # ---------------- stage-2 trace (calLvis.py:339-397), RNG captured ------------------------------
    m_pts = 12
    rs = np.random.RandomState(11)
    surf = rs.standard_normal((m_pts, 3)); surf = 0.5 * surf / np.linalg.norm(surf, axis=1, keepdims=True)
    surf = torch.from_numpy(surf.astype(np.float32))
    normal = mods["sdf"].gradient(surf.clone()).squeeze(1).detach()
    torch.manual_seed(5)
    r1, r2 = torch.rand(m_pts, 4), torch.rand(m_pts, 4)
    torch.manual_seed(5)

    class _Zero(torch.nn.Module):
        def forward(self, *a):
            return torch.zeros(a[0].shape[0], 1)

    class _ZeroSG(torch.nn.Module):
        def forward(self, p):
            return torch.ones(p.shape[0], 24, 7)

    res = calLvis.cal_indiLgt(surf, normal, mods["sdf"], mods["var"], mods["color"], _Zero(), _ZeroSG())
    np.savez_compressed(os.path.join(gdir, "lvis.npz"), surf=np_(surf), normal=np_(normal),
                        r_theta=np_(r1 * 2 * np.pi), rand_z=np_(r2 * 0.95),
                        gt_lvis=np_(res["gt_lvis"]), gt_trace_radiance=np_(res["gt_trace_radiance"]))
    lvis_render_golden(fields, renderer, syn, states, gdir)
    print("golden fixtures written to", gdir)


def lvis_render_golden(fields, renderer, syn, states, gdir):
    """NeuSRenderer.lvis_render itself (renderer.py:567-627) with the real Lvis / IndirectLight networks.  Its only
    random numbers are the two torch.rand([m,4]) draws of calLvis.py:351-352 over the m hit rays: they are captured by
    re-seeding, and scattered to per-ray rows for the fixed-shape implementation."""
    np_ = lambda t: t.detach().cpu().numpy()
    mods = build_reference(fields, renderer, states, syn.SDF_CONF, syn.COLOR_CONF, syn.NERF_CONF, syn.RENDER_CONF_WMASK)
    lv, il = fields.Lvis(), fields.IndirectLight()
    z3 = torch.zeros(2, 3)
    lv(z3, torch.ones(2, 3)); il(z3)
    lv.load_state_dict(syn.lvis_state()); il.load_state_dict(syn.indirect_light_state())
    R = mods["renderer"]
    R.lvis_network, R.indiLgt_network = lv, il
    B = 24
    o, d, near, far = syn.make_rays(B, seed=1)
    # a few rays that miss the surface / the unit sphere, so that the default-ones rows are exercised
    o[-2:] = o[-2:] * 1.0 + torch.tensor([[0.0, 3.0, 0.0]])
    torch.manual_seed(9)
    out = R.lvis_render(o, d, near, far)
    mask = out["sdf_mask"]
    m = int(mask.sum())
    torch.manual_seed(9)
    r1, r2 = torch.rand(m, 4), torch.rand(m, 4)
    r_theta, rand_z = torch.zeros(B, 4), torch.zeros(B, 4)
    r_theta[mask] = r1 * 2 * np.pi
    rand_z[mask] = r2 * 0.95
    util = R.lvis_mateIllu_render_util(o, d, near, far)
    save = dict(o=np_(o), d=np_(d), near=np_(near), far=np_(far), r_theta=np_(r_theta), rand_z=np_(rand_z),
                sdf_mask=np_(mask.float()), inside_sphere_mask=np_(util["inside_sphere_mask"].float()),
                mid_z_vals=np_(util["mid_z_vals"]), sdf=np_(util["sdf"]))
    for k in ("gt_lvis", "pre_lvis", "gt_trace_radiance", "pre_trace_radiance"):
        save[k] = np_(out[k])
    np.savez_compressed(os.path.join(gdir, "lvis_render.npz"), **save)
    print("lvis_render: %d of %d rays hit" % (m, B))


def _upsample_weights_via_reference(R, o, d, z, sdf, inv_s):
    """Run reference up_sample with sample_pdf intercepted to capture the weights it is given."""
    import models.renderer as rr
    box = {}
    orig = rr.sample_pdf

    def spy(bins, weights, n, det=False):
        box["w"] = weights.clone()
        return orig(bins, weights, n, det=det)

    rr.sample_pdf = spy
    try:
        R.up_sample(o, d, z, sdf, 16, inv_s)
    finally:
        rr.sample_pdf = orig
    return box["w"]


if __name__ == "__main__":
    torch.set_grad_enabled(True)
    main()
