"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the committed golden
fixtures.  Tolerances follow BASELINE.json north_star: FP32 path <= 1e-4 max-abs on colours, SDF gradients
and weight gradients; searchsorted indices bit-exact given identical CDFs."""
import os

import numpy as np
import pytest
import torch

import factored_neus_b200 as fn
from factored_neus_b200 import ops
from oracle import neus_oracle as O
from util import assert_close, build_modules, compare_param_grads, grad_params, max_err, syn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FP32_TOL = 1e-4


@pytest.fixture(scope="module")
def states():
    return syn.scene_states(seed=4, jitter=0.03)


def _golden(golden_dir, name):
    return {k: v for k, v in np.load(os.path.join(golden_dir, name)).items()}


def _cu(a):
    return torch.as_tensor(np.ascontiguousarray(a)).to(DEV)


# ------------------------------------------------------------------------------------------ sampling
def test_searchsorted_indices_bit_exact(golden_dir):
    g = _golden(golden_dir, "sampling.npz")
    u = _cu(g["u"])
    for i in range(4):
        bins, cdf = _cu(g["z%d" % i]), _cu(g["cdf%d" % i])
        samples, inds = ops.inverse_cdf(bins, cdf, u)
        assert np.array_equal(inds.cpu().numpy(), g["inds%d" % i]), "indices differ at step %d" % i
        ref, _ = O.invert_cdf(torch.from_numpy(g["z%d" % i]), torch.from_numpy(g["cdf%d" % i]),
                              torch.from_numpy(g["u"]).expand(bins.shape[0], -1))
        assert_close(samples, ref, 1e-6, "inverse-cdf samples %d" % i)


def test_inverse_cdf_random_bit_exact():
    """Larger randomised check incl. flat (repeated) CDF entries and the n=512,k=32 lvis shape."""
    gen = torch.Generator().manual_seed(3)
    for B, n, k in ((257, 64, 16), (64, 112, 16), (33, 512, 32), (5, 2, 4)):
        w = torch.rand(B, n - 1, generator=gen)
        w[w < 0.3] = 0.0
        cdf = torch.cat([torch.zeros(B, 1), torch.cumsum(w / w.sum(-1, keepdim=True).clamp_min(1e-9), -1)], -1)
        bins = torch.sort(torch.rand(B, n, generator=gen), -1)[0]
        u = torch.linspace(0.5 / k, 1 - 0.5 / k, k)
        ref_s, ref_i = O.invert_cdf(bins, cdf, u.expand(B, k))
        s, i = ops.inverse_cdf(bins.to(DEV), cdf.to(DEV), u.to(DEV))
        assert torch.equal(i.cpu(), ref_i), "indices differ for shape %s" % ((B, n, k),)
        assert_close(s, ref_s, 1e-6, "samples %s" % ((B, n, k),))


def test_upsample_and_merge_chain(golden_dir):
    g = _golden(golden_dir, "sampling.npz")
    o, d = _cu(g["o"]), _cu(g["d"])
    u = _cu(g["u"])
    for i in range(4):
        z, sdf = _cu(g["z%d" % i]), _cu(g["sdf%d" % i])
        new_z, cdf, inds = ops.upsample_step(o, d, z, sdf, 16, 64 * 2 ** i, u, debug=True)
        assert_close(cdf, g["cdf%d" % i], 2e-6, "cdf %d" % i)
        # interpolation in near-empty bins amplifies cdf rounding: compare on the bins' scale
        assert_close(new_z, g["newz%d" % i], 2e-4, "new z %d" % i)
        frac_same = (inds.cpu().numpy() == g["inds%d" % i]).mean()
        assert frac_same > 0.99, "only %.3f of indices agree" % frac_same
        nz = _cu(g["newz%d" % i])
        if i < 3:
            pts = ops.ray_points(o, d, nz)
            assert_close(pts.reshape(-1, 16, 3), (g["o"][:, None, :] + g["d"][:, None, :] * g["newz%d" % i][..., None]),
                         0.0, "ray points")
            # carry the reference's new sdf values through the merge
            ref_new_sdf = O.sdf_value(syn.scene_states(seed=4, jitter=0.03)["sdf"],
                                      torch.from_numpy(pts.cpu().numpy())).reshape(-1, 16)
            zz, ss = ops.merge_sorted(z, nz, sdf, ref_new_sdf.to(DEV))
            assert_close(ss, g["sdf%d" % (i + 1)], 2e-6, "merged sdf %d" % i)
        else:
            zz, _ = ops.merge_sorted(z, nz)
        assert np.array_equal(zz.cpu().numpy(), g["z%d" % (i + 1)]), "merged z %d" % i


def test_merge_sorted_properties():
    gen = torch.Generator().manual_seed(0)
    for B, n, k in ((1000, 64, 16), (7, 128, 32), (3, 1, 1), (50, 512, 32), (5, 2000, 100), (9, 33, 70), (4, 31, 1)):
        a = torch.sort(torch.rand(B, n, generator=gen), -1)[0]
        b = torch.sort(torch.rand(B, k, generator=gen), -1)[0]
        b[:, 0] = a[:, 0]                                           # ties
        if k > 2 and n > 5:
            b[:, 1] = a[:, 5]
            b[:, 2] = a[:, 5]                                       # a double tie
        b = torch.sort(b, -1)[0]
        sa, sb = torch.rand(B, n, generator=gen), torch.rand(B, k, generator=gen)
        z, s = ops.merge_sorted(a.to(DEV), b.to(DEV), sa.to(DEV), sb.to(DEV))
        ref, idx = torch.sort(torch.cat([a, b], -1), dim=-1, stable=True)
        assert torch.equal(z.cpu(), ref)
        assert torch.equal(s.cpu(), torch.cat([sa, sb], -1).gather(-1, idx))


# ------------------------------------------------------------------------------------------ fields
def test_fields_forward(golden_dir, states):
    g = _golden(golden_dir, "fields.npz")
    m = build_modules(states, DEV)
    x, v = _cu(g["x"]), _cu(g["v"])
    with torch.no_grad():
        out = m["sdf"](x)
        assert_close(out, g["sdf_out"], 1e-5, "sdf forward (no-grad path)")
        assert_close(m["sdf"].sdf(x), g["sdf_out"][:, :1], 1e-5, "sdf() (no-grad path)")
    sdf, feat, nrm = m["sdf"].value_feature_normal(x)
    assert_close(sdf, g["sdf_out"][:, :1], 1e-5, "sdf value")
    assert_close(feat, g["sdf_out"][:, 1:], 1e-5, "sdf feature")
    assert_close(nrm, g["grad"], 2e-5, "sdf normal")
    assert_close(m["sdf"].gradient(x), g["grad"][:, None, :], 2e-5, "sdf.gradient")
    rgb = m["color"](x, _cu(g["grad"]), v, _cu(g["sdf_out"][:, 1:]))
    assert_close(rgb, g["rgb"], 1e-5, "colour")
    rd = m["ref"](x, _cu(g["sdf_out"][:, 1:]), v, _cu(g["grad"]))
    assert_close(rd["rgb"], g["ref_rgb"], 1e-5, "refcolor rgb")
    assert_close(rd["specular_rgb"], g["ref_spec"], 1e-5, "refcolor specular")
    assert_close(rd["diffuse_rgb"], g["ref_diff"], 1e-5, "refcolor diffuse")


@pytest.mark.parametrize("N", [1, 130, 1000])
def test_sdf_backward_incl_double_backward(states, N):
    """d/dW of a random functional of (sdf, feature, normal) -- the normal term is the second-order path
    (SoftplusBackwardBackward in the reference)."""
    gen = torch.Generator().manual_seed(N)
    x = (torch.rand(N, 3, generator=gen) * 2 - 1)
    c_s, c_f, c_n = torch.randn(N, 1, generator=gen), torch.randn(N, 256, generator=gen) * 0.1, \
        torch.randn(N, 3, generator=gen)
    P = grad_params(states)
    out = O.sdf_forward(P["sdf"], x)
    nrm = O.sdf_gradient(P["sdf"], x)
    ((out[:, :1] * c_s).sum() + (out[:, 1:] * c_f).sum() + (nrm * c_n).sum()).backward()
    m = build_modules(states, DEV)
    sdf, feat, n2 = m["sdf"].value_feature_normal(x.to(DEV))
    ((sdf * c_s.to(DEV)).sum() + (feat * c_f.to(DEV)).sum() + (n2 * c_n.to(DEV)).sum()).backward()
    scale = max(1.0, max(float(t.grad.abs().max()) for t in P["sdf"].values()))
    compare_param_grads(m, P, ["sdf"], FP32_TOL * scale, 1e-4, "sdf N=%d" % N)


def test_sdf_backward_value_only(states):
    gen = torch.Generator().manual_seed(5)
    x = (torch.rand(300, 3, generator=gen) * 2 - 1)
    c = torch.randn(300, 257, generator=gen) * 0.1
    P = grad_params(states)
    (O.sdf_forward(P["sdf"], x) * c).sum().backward()
    m = build_modules(states, DEV)
    (m["sdf"](x.to(DEV)) * c.to(DEV)).sum().backward()
    scale = max(1.0, max(float(t.grad.abs().max()) for t in P["sdf"].values()))
    compare_param_grads(m, P, ["sdf"], FP32_TOL * scale, 1e-4, "sdf value-only")


def test_color_and_refcolor_backward(states):
    gen = torch.Generator().manual_seed(9)
    N = 300
    x = torch.rand(N, 3, generator=gen) * 2 - 1
    v = torch.nn.functional.normalize(torch.randn(N, 3, generator=gen), dim=-1)
    nrm = torch.randn(N, 3, generator=gen)
    feat = torch.randn(N, 256, generator=gen) * 0.3
    c = torch.randn(N, 3, generator=gen)
    P = grad_params(states)
    n_o, f_o = nrm.clone().requires_grad_(True), feat.clone().requires_grad_(True)
    (O.color_forward(P["color"], x, n_o, v, f_o) * c).sum().backward()
    m = build_modules(states, DEV)
    n_g, f_g = nrm.to(DEV).requires_grad_(True), feat.to(DEV).requires_grad_(True)
    (m["color"](x.to(DEV), n_g, v.to(DEV), f_g) * c.to(DEV)).sum().backward()
    assert_close(n_g.grad, n_o.grad, FP32_TOL, "colour d_normals")
    assert_close(f_g.grad, f_o.grad, FP32_TOL, "colour d_features")
    compare_param_grads(m, P, ["color"], FP32_TOL, 1e-4, "colour")

    n_o, f_o = nrm.clone().requires_grad_(True), feat.clone().requires_grad_(True)
    r, s, d = O.refcolor_forward(P["ref"], x, f_o, v, n_o)
    c2, c3 = torch.randn(N, 3, generator=gen), torch.randn(N, 3, generator=gen)
    ((r * c).sum() + (s * c2).sum() + (d * c3).sum()).backward()
    n_g, f_g = nrm.to(DEV).requires_grad_(True), feat.to(DEV).requires_grad_(True)
    rd = m["ref"](x.to(DEV), f_g, v.to(DEV), n_g)
    ((rd["rgb"] * c.to(DEV)).sum() + (rd["specular_rgb"] * c2.to(DEV)).sum()
     + (rd["diffuse_rgb"] * c3.to(DEV)).sum()).backward()
    assert_close(n_g.grad, n_o.grad, FP32_TOL, "refcolor d_normals", rtol=1e-4)
    assert_close(f_g.grad, f_o.grad, FP32_TOL, "refcolor d_features", rtol=1e-4)
    compare_param_grads(m, P, ["ref"], FP32_TOL, 1e-4, "refcolor")


# ------------------------------------------------------------------------------------------ composite
@pytest.mark.parametrize("n_out,car,use_bg_rgb", [(0, 1.0, False), (0, 0.3, True), (32, 0.3, False)])
def test_composite_fwd_bwd(n_out, car, use_bg_rgb):
    gen = torch.Generator().manual_seed(11 + n_out)
    B, n = 37, 128
    o, d, near, far = syn.make_rays(B, seed=3)
    z = torch.sort(near + (far - near) * torch.rand(B, n, generator=gen), -1)[0]
    dists = torch.cat([z[:, 1:] - z[:, :-1], torch.full((B, 1), 2.0 / 64)], -1)
    pts = (o[:, None, :] + d[:, None, :] * (z + dists * 0.5)[..., None]).reshape(-1, 3)
    sdf = (pts.norm(dim=-1, keepdim=True) - 0.5 + 0.02 * torch.randn(B * n, 1, generator=gen))
    nrm = torch.nn.functional.normalize(pts, dim=-1) * (1 + 0.2 * torch.randn(B * n, 1, generator=gen))
    rgb = torch.rand(B * n, 3, generator=gen)
    var = torch.tensor(0.3)
    bga = torch.rand(B, n + n_out, generator=gen) * 0.2 if n_out else None
    bgc = torch.rand(B, n + n_out, 3, generator=gen) if n_out else None
    bgr = torch.tensor([[0.7, 0.8, 0.9]]) if use_bg_rgb else None
    c_col, c_w = torch.randn(B, 3, generator=gen), torch.randn(B, n + n_out, generator=gen) * 0.1
    c_pair = torch.randn(B, 2, generator=gen)

    def run_oracle():
        leaves = [t.clone().requires_grad_(True) for t in (sdf, nrm, rgb, var)]
        bl = [t.clone().requires_grad_(True) for t in (bga, bgc)] if n_out else [None, None]
        s_, n_, c_, v_ = leaves
        inv_s = O.inv_s_of(v_).reshape(1, 1)
        dirs = d[:, None, :].expand(B, n, 3).reshape(-1, 3)
        alpha, c_prev = O.neus_alpha(s_, n_, dirs, dists, inv_s, car)
        r = pts.norm(dim=-1).reshape(B, n)
        inside, relax = (r < 1.0).float(), (r < 1.2).float()
        hit, idx = O.first_hit(s_.reshape(B, n).detach(), inside)
        w_in = O.transmittance_weights(alpha * inside)
        ii = idx.clamp_min(1)
        pair = torch.stack([w_in.gather(1, (ii - 1)[:, None]), w_in.gather(1, ii[:, None])], 1).reshape(B, 2) + 1e-5
        pair = torch.where(hit[:, None], pair, torch.ones_like(pair))
        col = c_.reshape(B, n, 3)
        if n_out:
            alpha = torch.cat([alpha * inside + bl[0][:, :n] * (1 - inside), bl[0][:, n:]], -1)
            col = torch.cat([col * inside[..., None] + bl[1][:, :n] * (1 - inside)[..., None], bl[1][:, n:]], 1)
        w = O.transmittance_weights(alpha)
        color = (col * w[..., None]).sum(1)
        if bgr is not None:
            color = color + bgr * (1 - w.sum(-1, keepdim=True))
        g3 = n_.reshape(B, n, 3)
        eik = (relax * (g3.norm(dim=-1) - 1) ** 2).sum() / (relax.sum() + 1e-5)
        loss = (color * c_col).sum() + (w * c_w).sum() + 0.7 * eik + (pair * c_pair).sum() + w.sum() * 0.3
        loss.backward()
        return dict(color=color, weights=w, cdf=c_prev.reshape(B, n), inside=inside, eik=eik, hit=hit, idx=idx,
                    pair=pair), [t.grad for t in leaves], [t.grad if t is not None else None for t in bl]

    ref, gl, gb = run_oracle()
    cu = lambda t: None if t is None else t.to(DEV)
    leaves = [t.to(DEV).requires_grad_(True) for t in (sdf, nrm, rgb, var)]
    bl = [t.to(DEV).requires_grad_(True) for t in (bga, bgc)] if n_out else [None, None]
    inv_s = ops.InvS.apply(leaves[3])               # clip(exp(10 variance)) as one launch (fields.py:267-268)
    assert float((inv_s.detach().cpu() - torch.exp(var * 10.0).clip(1e-6, 1e6)).abs().max()) <= 1e-6 * float(inv_s)
    color, w, wsum, wmax, cdf, inside, eik_num, eik_den, hit_idx, pair, eik, hit_mask = ops.Composite.apply(
        leaves[0], leaves[1], leaves[2], inv_s, bl[0], bl[1], cu(dists), cu(pts), cu(d), cu(bgr), n, n_out, car)
    assert torch.equal(hit_mask, hit_idx >= 0)
    assert abs(float(eik) - float(eik_num / (eik_den + 1e-5))) <= 1e-6 * abs(float(eik)) + 1e-12
    assert_close(color, ref["color"], 1e-5, "color")
    assert_close(w, ref["weights"], 1e-5, "weights")
    assert_close(wsum, ref["weights"].sum(-1, keepdim=True), 1e-5, "weight_sum")
    assert_close(wmax, ref["weights"].max(-1, keepdim=True)[0], 1e-5, "weight_max")
    assert_close(cdf, ref["cdf"], 1e-5, "cdf")
    assert torch.equal(inside.cpu(), ref["inside"])
    assert_close(eik, ref["eik"], 1e-5, "gradient_error")
    assert torch.equal((hit_idx >= 0).cpu(), ref["hit"])
    assert torch.equal(hit_idx.cpu()[ref["hit"]].long(), ref["idx"][ref["hit"]])
    assert_close(pair, ref["pair"], 1e-5, "w_pair")
    loss = (color * cu(c_col)).sum() + (w * cu(c_w)).sum() + 0.7 * eik + (pair * cu(c_pair)).sum() + wsum.sum() * 0.3
    loss.backward()
    for name, a, b in zip(("d_sdf", "d_normals", "d_rgb", "d_variance"), leaves, gl):
        assert_close(a.grad, b, FP32_TOL, name, rtol=1e-4)
    if n_out:
        assert_close(bl[0].grad, gb[0], FP32_TOL, "d_bg_alpha", rtol=1e-4)
        assert_close(bl[1].grad, gb[1], FP32_TOL, "d_bg_color", rtol=1e-4)


# ------------------------------------------------------------------------------------------ render
def _stage1(out, true_rgb, mask, mw):
    return O.stage1_loss(out, true_rgb, mask, 0.1, 0.1, mw)[0]


def test_render_core_fwd_bwd_wmask(golden_dir, states):
    """Full wmask step on the reference's own depths: outputs vs golden (reference) and vs oracle,
    every parameter gradient vs the oracle's autograd."""
    g = _golden(golden_dir, "render_wmask.npz")
    B = g["color_fine"].shape[0]
    o, d, near, far = syn.make_rays(B, seed=1)
    true_rgb, mask = syn.make_targets(B, seed=2)
    z = torch.from_numpy(g["z_vals"])
    P = grad_params(states)
    ref = O.render(P, o, d, near, far, conf=O.RENDER_CONF_WMASK, perturb_overwrite=0, cos_anneal_ratio=1.0,
                   z_override=z)
    _stage1(ref, true_rgb, mask, 0.1).backward()

    m = build_modules(states, DEV, syn.RENDER_CONF_WMASK)
    R = m["renderer"]
    core = R.render_core(o.to(DEV), d.to(DEV), z.to(DEV), 2.0 / 64, m["sdf"], m["var"], m["color"], m["ref"],
                         cos_anneal_ratio=1.0)
    w = core["weights"]
    out = dict(color_fine=core["color"], surface_color=core["surface_color"], sdf_mask=core["sdf_mask"],
               weight_sum=w.sum(-1, keepdim=True), gradient_error=core["gradient_error"])
    for k, gk in (("color", "color_fine"), ("surface_color", "surface_color"), ("cdf", "cdf_fine"),
                  ("gradients", "gradients"), ("weights", "weights"), ("gradient_error", "gradient_error"),
                  ("inside_sphere", "inside_sphere"), ("specular_color", "specular_color"),
                  ("diffuse_color", "diffuse_color")):
        assert_close(core[k], g[gk], FP32_TOL, "render_core %s vs reference golden" % k)
    assert np.array_equal(core["sdf_mask"].cpu().numpy().astype(np.float32), g["sdf_mask"])
    loss = _stage1(out, true_rgb.to(DEV), mask.to(DEV), 0.1)
    assert abs(loss.item() - float(g["loss"])) < 1e-4
    loss.backward()
    worst = compare_param_grads(m, P, ["sdf", "color", "var", "ref"], FP32_TOL, 1e-3, "wmask")
    print("wmask worst param-grad abs err %.3e" % worst)


def test_render_end_to_end_wmask(golden_dir, states):
    g = _golden(golden_dir, "render_wmask.npz")
    B = g["color_fine"].shape[0]
    o, d, near, far = syn.make_rays(B, seed=1)
    m = build_modules(states, DEV, syn.RENDER_CONF_WMASK)
    out = m["renderer"].render(o.to(DEV), d.to(DEV), near.to(DEV), far.to(DEV), perturb_overwrite=0,
                               cos_anneal_ratio=1.0)
    for k in ("color_fine", "surface_color", "s_val", "weight_sum", "weight_max", "gradient_error",
              "specular_color", "diffuse_color"):
        assert_close(out[k], g[k], FP32_TOL, "render %s vs reference golden" % k)
    assert set(out.keys()) == {"color_fine", "surface_color", "sdf_mask", "s_val", "cdf_fine", "weight_sum",
                               "weight_max", "gradients", "weights", "gradient_error", "inside_sphere",
                               "specular_color", "diffuse_color"}
    assert out["gradients"].shape == (B, 128, 3) and out["weights"].shape == (B, 128)
    assert out["sdf_mask"].dtype == torch.bool


def test_render_perturbed_runs_and_is_sane(states):
    m = build_modules(states, DEV, syn.RENDER_CONF_WMASK)
    o, d, near, far = syn.make_rays(512, seed=1)
    out = m["renderer"].render(o.to(DEV), d.to(DEV), near.to(DEV), far.to(DEV), cos_anneal_ratio=1.0)
    assert torch.isfinite(out["color_fine"]).all()
    ws = out["weight_sum"]
    assert float(ws.min()) >= -1e-5 and float(ws.max()) <= 1.0 + 1e-4
    assert int(out["sdf_mask"].sum()) > 256


# ------------------------------------------------------------------------------------------ womask
def test_nerf_forward_backward(golden_dir, states):
    g = _golden(golden_dir, "fields.npz")
    m = build_modules(states, DEV)
    x4, v = _cu(g["x4"]), _cu(g["v"])
    dens, rgb = m["nerf"](x4, v)
    assert_close(dens, g["nerf_density"], 1e-5, "nerf density vs reference golden")
    assert_close(rgb, g["nerf_rgb"], 1e-5, "nerf rgb vs reference golden")
    gen = torch.Generator().manual_seed(2)
    N = 700
    p4 = torch.rand(N, 4, generator=gen) * 2 - 1
    vv = torch.nn.functional.normalize(torch.randn(N, 3, generator=gen), dim=-1)
    c1, c3 = torch.randn(N, 1, generator=gen), torch.randn(N, 3, generator=gen)
    P = grad_params(states)
    d_o, r_o = O.nerf_forward(P["nerf"], p4, vv)
    ((d_o * c1).sum() + (r_o * c3).sum()).backward()
    d_g, r_g = m["nerf"](p4.to(DEV), vv.to(DEV))
    assert_close(d_g, d_o, 1e-5, "nerf density")
    assert_close(r_g, r_o, 1e-5, "nerf rgb")
    ((d_g * c1.to(DEV)).sum() + (r_g * c3.to(DEV)).sum()).backward()
    compare_param_grads(m, P, ["nerf"], FP32_TOL, 1e-4, "nerf")


def test_render_core_fwd_bwd_womask(golden_dir, states):
    """womask (n_outside=32, cos_anneal 0.3): outputs vs the reference golden, all grads vs the oracle."""
    g = _golden(golden_dir, "render_womask.npz")
    B = g["color_fine"].shape[0]
    o, d, near, far = syn.make_rays(B, seed=1)
    true_rgb, mask = syn.make_targets(B, seed=2)
    z = torch.from_numpy(g["z_vals"])
    P = grad_params(states)
    ref = O.render(P, o, d, near, far, conf=O.RENDER_CONF_WOMASK, perturb_overwrite=0, cos_anneal_ratio=0.3,
                   z_override=z)
    _stage1(ref, true_rgb, mask, 0.0).backward()

    m = build_modules(states, DEV, syn.RENDER_CONF_WOMASK)
    R = m["renderer"]
    od, dd, zd = o.to(DEV), d.to(DEV), z.to(DEV)
    z_out = O.outside_z(far, 32, 64).to(DEV).contiguous()
    z_feed, _ = ops.merge_sorted(zd, z_out)
    assert_close(z_feed, torch.sort(torch.cat([z, O.outside_z(far, 32, 64)], -1), -1)[0], 0.0, "z_feed")
    ro = R.render_core_outside(od, dd, z_feed, 2.0 / 64, m["nerf"])
    core = R.render_core(od, dd, zd, 2.0 / 64, m["sdf"], m["var"], m["color"], m["ref"],
                         background_alpha=ro["alpha"], background_sampled_color=ro["sampled_color"],
                         cos_anneal_ratio=0.3)
    w = core["weights"]
    assert w.shape == (B, 160)
    for k, gk in (("color", "color_fine"), ("surface_color", "surface_color"), ("cdf", "cdf_fine"),
                  ("gradients", "gradients"), ("weights", "weights"), ("gradient_error", "gradient_error"),
                  ("inside_sphere", "inside_sphere")):
        assert_close(core[k], g[gk], FP32_TOL, "womask render_core %s vs reference golden" % k)
    out = dict(color_fine=core["color"], surface_color=core["surface_color"], sdf_mask=core["sdf_mask"],
               weight_sum=w.sum(-1, keepdim=True), gradient_error=core["gradient_error"])
    loss = _stage1(out, true_rgb.to(DEV), mask.to(DEV), 0.0)
    assert abs(loss.item() - float(g["loss"])) < 1e-4
    loss.backward()
    compare_param_grads(m, P, ["sdf", "color", "var", "ref", "nerf"], FP32_TOL, 1e-3, "womask")


def test_render_end_to_end_womask(golden_dir, states):
    g = _golden(golden_dir, "render_womask.npz")
    B = g["color_fine"].shape[0]
    o, d, near, far = syn.make_rays(B, seed=1)
    m = build_modules(states, DEV, syn.RENDER_CONF_WOMASK)
    out = m["renderer"].render(o.to(DEV), d.to(DEV), near.to(DEV), far.to(DEV), perturb_overwrite=0,
                               cos_anneal_ratio=0.3)
    for k in ("color_fine", "surface_color", "s_val", "weight_sum", "weight_max", "gradient_error"):
        assert_close(out[k], g[k], FP32_TOL, "womask render %s vs reference golden" % k)
    assert out["weights"].shape == (B, 160) and out["gradients"].shape == (B, 128, 3)
    out2 = m["renderer"].render(o.to(DEV), d.to(DEV), near.to(DEV), far.to(DEV), cos_anneal_ratio=0.3)
    assert torch.isfinite(out2["color_fine"]).all()


def test_grid_query_matches_reference(golden_dir, states):
    g = _golden(golden_dir, "grid.npz")
    m = build_modules(states, DEV, syn.RENDER_CONF_WMASK)
    R = m["renderer"]
    bmin, bmax = torch.from_numpy(g["bmin"]), torch.from_numpy(g["bmax"])
    u = R.extract_fields(bmin, bmax, 20)
    assert_close(u, g["u"], 1e-5, "grid vs reference golden")
    slab = R.extract_fields(bmin, bmax, 20, ix0=5, ix1=9)               # x-slab sharding (multi-GPU unit)
    assert torch.equal(slab, u[5:9])


def test_lvis_trace_matches_reference(golden_dir, states):
    """Stage-2 ground truth (calLvis.cal_indiLgt gt_lvis / gt_trace_radiance) vs the reference golden."""
    from factored_neus_b200 import lvis as LV
    g = _golden(golden_dir, "lvis.npz")
    m = build_modules(states, DEV)
    lv, rad, dirs = LV.trace_visibility(_cu(g["surf"]), _cu(g["normal"]), m["sdf"], m["var"], m["color"],
                                        _cu(g["r_theta"]), _cu(g["rand_z"]))
    assert_close(lv, g["gt_lvis"], FP32_TOL, "gt_lvis vs reference golden")
    assert_close(rad, g["gt_trace_radiance"], FP32_TOL, "gt_trace_radiance vs reference golden")


def test_lvis_render_shapes(states):
    m = build_modules(states, DEV, syn.RENDER_CONF_WMASK)
    R = m["renderer"]

    class _L(torch.nn.Module):
        def forward(self, p, v):
            return torch.sigmoid(p.sum(-1, keepdim=True))

    class _I(torch.nn.Module):
        def forward(self, p):
            return torch.ones(p.shape[0], 24, 7, device=p.device)

    R.lvis_network, R.indiLgt_network = _L(), _I()
    o, d, near, far = syn.make_rays(64, seed=1)
    out = R.lvis_render(o.to(DEV), d.to(DEV), near.to(DEV), far.to(DEV))
    assert out["gt_lvis"].shape == (64, 4) and out["gt_trace_radiance"].shape == (64, 4, 3)
    assert out["sdf_mask"].dtype == torch.bool and int(out["sdf_mask"].sum()) > 32
    assert float(out["gt_lvis"].min()) >= -1e-4 and float(out["gt_lvis"].max()) <= 1.0 + 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("N", [80, 333])
def test_stage2_networks_fwd_bwd(golden_dir, N):
    """Lvis / IndirectLight through fneus_mlp_fwd/bwd vs the oracle (and the reference fixture at N=80)."""
    g = dict(np.load(os.path.join(golden_dir, "stage2_nets.npz")))
    rs = np.random.RandomState(13 if N == 80 else 14)
    pts = torch.from_numpy(rs.uniform(-1, 1, (N, 3)).astype(np.float32))
    view = torch.from_numpy(rs.standard_normal((N, 3)).astype(np.float32))
    view = view / view.norm(dim=-1, keepdim=True)
    pv = torch.from_numpy(rs.standard_normal((N, 1)).astype(np.float32))
    ps = torch.from_numpy(rs.standard_normal((N, 24, 7)).astype(np.float32))
    lv, il = fn.Lvis(), fn.IndirectLight()
    lv.load_state_dict(syn.lvis_state()); il.load_state_dict(syn.indirect_light_state())
    lv, il = lv.to(DEV), il.to(DEV)
    vis = lv(pts.to(DEV), view.to(DEV))
    sgs = il(pts.to(DEV))
    (vis * pv.to(DEV)).sum().backward()
    (sgs * ps.to(DEV)).sum().backward()
    Pl = {n: t.clone().requires_grad_(True) for n, t in syn.lvis_state().items()}
    Pi = {n: t.clone().requires_grad_(True) for n, t in syn.indirect_light_state().items()}
    vis_o, sgs_o = O.lvis_forward(Pl, pts, view), O.indirect_light_forward(Pi, pts)
    (vis_o * pv).sum().backward()
    (sgs_o * ps).sum().backward()
    assert_close(vis, vis_o, 1e-5, "lvis")
    assert_close(sgs, sgs_o, 1e-4, "indirect-light SGs")
    if N == 80:
        assert_close(vis, g["vis"], 1e-5, "lvis vs reference")
        assert_close(sgs, g["sgs"], 1e-4, "SGs vs reference")
    for mod, P, tag in ((lv, Pl, "lvis"), (il, Pi, "indi")):
        for name, p in mod.named_parameters():
            assert_close(p.grad, P[name].grad, 1e-4, "grad %s.%s" % (tag, name), rtol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("B,use_mask", [(512, True), (37, True), (300, False)])
def test_fused_loss_and_surface_blend(B, use_mask):
    """csrc/loss.cu vs the plain torch restatement (oracle stage1_loss, renderer.py:328-343 blend): values and
    gradients, FP32 <= 1e-5."""
    from factored_neus_b200.parallel import stage1_loss_sharded
    rs = np.random.RandomState(B)
    f = lambda *s: torch.from_numpy(rs.uniform(0.02, 0.98, s).astype(np.float32))
    c2 = [f(2 * B, 3) for _ in range(3)]
    w_pair = f(B, 2)
    hit_idx = torch.from_numpy(np.where(rs.uniform(size=B) < 0.7, rs.randint(1, 127, B), -1).astype(np.int32))
    color, wsum, true_rgb = f(B, 3), f(B, 1) * 1.05, f(B, 3)
    wsum[::7] = 1e-4                                     # outside the BCE clip
    mask = (torch.from_numpy(rs.uniform(size=(B, 1)).astype(np.float32)) > 0.3).float()
    eik_num, eik_den = torch.tensor(3.7), torch.tensor(41.0)
    mw = 0.1 if use_mask else 0.0

    def run(dev):
        leaves = [t.clone().to(dev).requires_grad_(True) for t in c2 + [w_pair, color, wsum, eik_num]]
        cr, cs, cd, wp, col, ws, en = leaves
        hit = (hit_idx >= 0).to(dev)
        if dev == "cpu":
            w0, w1 = wp[:, :1], wp[:, 1:]
            blend = lambda c: torch.where(hit[:, None], (c.reshape(B, 2, 3)[:, 0] * w0 + c.reshape(B, 2, 3)[:, 1] * w1)
                                          / (w0 + w1), torch.ones(B, 3))
            s_rgb, s_spec, s_diff = blend(cr), blend(cs), blend(cd)
        else:
            s_rgb, s_spec, s_diff = ops.SurfaceBlend.apply(cr, cs, cd, wp, hit_idx.to(dev))

        class R:
            pass
        R.last_eikonal_parts = (en, eik_den.to(dev))
        R.last_hit_idx = hit_idx.to(dev)
        out = dict(color_fine=col, surface_color=s_rgb, weight_sum=ws, sdf_mask=hit)
        loss, stats = stage1_loss_sharded(R, out, true_rgb.to(dev), mask.to(dev), 0.1, 0.1, mw)
        (loss + 0.3 * (s_spec * s_spec).sum() + 0.2 * s_diff.sum()).backward()
        return loss, stats, [s_rgb, s_spec, s_diff], [t.grad for t in leaves]

    l_ref, st_ref, s_ref, g_ref = run("cpu")
    l_got, st_got, s_got, g_got = run(DEV)
    assert abs(l_ref.item() - l_got.item()) <= 1e-5 * max(1.0, abs(l_ref.item()))
    for k in st_ref:
        assert abs(float(st_ref[k]) - float(st_got[k])) <= 1e-5 * max(1.0, abs(float(st_ref[k]))), k
    for a, b in zip(s_got, s_ref):
        assert_close(a, b.detach(), 1e-5, "surface blend")
    names = ["c_rgb", "c_spec", "c_diff", "w_pair", "color", "weight_sum", "eik_num"]
    for n_, a, b in zip(names, g_got, g_ref):
        assert_close(a, b, 1e-5, "grad " + n_, rtol=1e-4)


def test_ray_setup_glue_kernels_bit_exact():
    """fneus_near_far / fneus_coarse_z / fneus_hit_rows against the framework expressions they replace
    (dataset.py:186-192: <= 1 ulp; renderer.py:395-408 and renderer.py:296-303: bit-exact)."""
    B, n = 777, 64
    g = torch.Generator().manual_seed(3)
    o = (2.5 * torch.nn.functional.normalize(torch.randn(B, 3, generator=g), dim=-1)).to(DEV)
    d = torch.nn.functional.normalize(torch.randn(B, 3, generator=g), dim=-1).to(DEV)
    a = torch.sum(d ** 2, dim=-1, keepdim=True)
    b = 2.0 * torch.sum(o * d, dim=-1, keepdim=True)
    mid = 0.5 * (-b) / a
    near, far = ops.near_far_from_sphere(o, d)
    # torch's 3-element reduction order is its own: a last-bit difference in the two sums is allowed here
    assert float((near - (mid - 1.0)).abs().max()) <= 5e-7 and float((far - (mid + 1.0)).abs().max()) <= 5e-7
    lin = torch.linspace(0.0, 1.0, n, device=DEV)
    z_ref = near + (far - near) * lin[None, :]
    assert torch.equal(ops.coarse_z(near, far, lin, None, n), z_ref)
    rnd = torch.rand(B, 1, generator=g).to(DEV)
    z_ref = z_ref + (rnd - 0.5) * 2.0 / n
    assert torch.equal(ops.coarse_z(near, far, lin, rnd, n), z_ref)
    hit = torch.randint(-1, 128, (B,), generator=g, dtype=torch.int32).to(DEV)
    idx = hit.clamp(min=1).long()
    base = torch.arange(B, device=DEV) * 128
    rows_ref = torch.stack([base + idx - 1, base + idx], dim=1).reshape(-1)
    assert torch.equal(ops.hit_rows(hit, 128), rows_ref)


# ------------------------------------------------------------------------------------------ stage 2 surface / lvis_render
def _first_hit_reference(sdf, mid_z, pts, o, d):
    """renderer.py:588-602 restated with the reference's own expressions (sign * ramp, torch.min)."""
    B, n = sdf.shape
    ramp = torch.arange(n, 0, -1).float().reshape(1, n)
    val, idx = torch.min(torch.sign(sdf) * ramp, dim=-1)
    inside = (torch.linalg.norm(pts.reshape(B, n, 3), ord=2, dim=-1) < 1.0).float()
    mask = (val < 0.0) & (idx >= 1) & (inside.sum(-1) > 0.0)
    ii = idx.clamp(1, n - 1).reshape(-1, 1)
    z_lo, z_hi = mid_z.gather(1, ii - 1), mid_z.gather(1, ii)
    s_lo, s_hi = sdf.gather(1, ii - 1), sdf.gather(1, ii)
    zs = (s_lo * z_hi - s_hi * z_lo) / (s_lo - s_hi + 1e-10)
    return mask, idx, zs, o + d * zs, inside


def test_first_hit_secant_kernel_incl_exact_zeros():
    """fneus_first_hit_secant vs the reference expressions on crafted rows: exact zeros (sign(0) = 0 is not a hit and
    does not shadow a later negative sample), a negative FIRST sample (idx = 0: invalid), no negative sample, rays
    entirely outside the unit sphere, -0.0, and random rows at n = 32 / 128 / 512."""
    gen = torch.Generator().manual_seed(8)
    for n in (32, 128, 512):
        B = 64
        sdf = torch.randn(B, n, generator=gen) * 0.3 + 0.2
        sdf[0, :] = sdf[0, :].abs() + 0.01                 # never negative
        sdf[1, 0] = -0.5                                   # negative at index 0 -> invalid
        sdf[2, :] = sdf[2, :].abs() + 0.01
        sdf[2, 5] = 0.0; sdf[2, 9] = -0.25                 # zero before the first negative
        sdf[3, :] = 0.0                                    # all zeros
        sdf[4, :] = sdf[4, :].abs() + 0.01
        sdf[4, 7] = -0.0                                   # negative zero: sign(-0.0) = 0
        mid_z = torch.sort(torch.rand(B, n, generator=gen) * 2 + 1.5, -1)[0]
        o = 2.5 * torch.nn.functional.normalize(torch.randn(B, 3, generator=gen), dim=-1)
        d = torch.nn.functional.normalize(-o + 0.3 * torch.randn(B, 3, generator=gen), dim=-1)
        o[5] = torch.tensor([0.0, 4.0, 0.0]); d[5] = torch.tensor([1.0, 0.0, 0.0])      # misses the unit sphere
        pts = (o[:, None, :] + d[:, None, :] * mid_z[..., None]).reshape(-1, 3)
        w = torch.rand(B, n, generator=gen) / n
        mask, idx, zs, ps, inside = _first_hit_reference(sdf, mid_z, pts, o, d)
        hit, z_k, p_k, lv, any_in = ops.first_hit_secant(_cu(sdf), _cu(mid_z), _cu(pts), _cu(o), _cu(d), weights=_cu(w))
        assert torch.equal((hit >= 0).cpu(), mask), "hit mask differs at n=%d" % n
        assert torch.equal(hit.cpu()[mask].long(), idx[mask])
        assert torch.equal(any_in.cpu(), inside.sum(-1) > 0)
        assert not mask[0] and not mask[1] and not mask[3] and not mask[4] and not mask[5]
        assert bool(mask[2]) and int(hit[2]) == 9
        assert torch.equal(z_k.cpu()[mask], zs[mask]), "secant root is not bit-equal at n=%d" % n
        assert_close(p_k.cpu()[mask], ps[mask], 1e-6, "surface point")
        assert_close(lv, 1.0 - (w * inside).sum(-1), 1e-6, "visibility reduction")


def _stage2_modules(states):
    m = build_modules(states, DEV, syn.RENDER_CONF_WMASK)
    lv, il = fn.Lvis(), fn.IndirectLight()
    lv.load_state_dict(syn.lvis_state()); il.load_state_dict(syn.indirect_light_state())
    R = m["renderer"]
    R.lvis_network, R.indiLgt_network = lv.to(DEV), il.to(DEV)
    return m, R


def test_lvis_render_matches_reference(golden_dir, states):
    """NeuSRenderer.lvis_render (renderer.py:567-627) against the reference's own output (fixture generated by
    tools/make_golden.py from the imported reference): hit mask identical, all four outputs <= 1e-4, default rows = 1."""
    g = _golden(golden_dir, "lvis_render.npz")
    m, R = _stage2_modules(states)
    util = R.lvis_mateIllu_render_util(_cu(g["o"]), _cu(g["d"]), _cu(g["near"]), _cu(g["far"]))
    assert np.array_equal(util["inside_sphere_mask"].cpu().numpy().astype(np.float32), g["inside_sphere_mask"])
    # inverse-CDF depths are ill-conditioned in near-empty bins: a few samples move by ~1e-4
    dz = (util["mid_z_vals"].cpu() - torch.from_numpy(g["mid_z_vals"])).abs()
    assert float(dz.max()) < 1e-3 and float((dz > 2e-5).float().mean()) < 0.01, "util mid_z_vals"
    out = R.lvis_render(_cu(g["o"]), _cu(g["d"]), _cu(g["near"]), _cu(g["far"]), r_theta=_cu(g["r_theta"]),
                        rand_z=_cu(g["rand_z"]))
    assert np.array_equal(out["sdf_mask"].cpu().numpy().astype(np.float32), g["sdf_mask"])
    for k in ("gt_lvis", "pre_lvis", "gt_trace_radiance", "pre_trace_radiance"):
        print("lvis_render %s max err %.3e" % (k, max_err(out[k], g[k])))
        assert_close(out[k], g[k], FP32_TOL, "lvis_render %s vs reference golden" % k)
    miss = g["sdf_mask"] == 0
    assert miss.any() and float((out["gt_lvis"].cpu()[torch.from_numpy(miss)] - 1.0).abs().max()) == 0.0


# ------------------------------------------------------------------------------------------ BASELINE-size step vs oracle
def test_render_512_rays_fwd_bwd_vs_oracle(states):
    """The benchmarked shape (512 rays x 128 samples = 65 536 points = 512 tiles) against the CPU oracle: end-to-end
    render outputs, then render_core + stage-1 loss + backward on the SAME depths (inverse-CDF sampling is
    ill-conditioned in empty bins) with every parameter gradient compared."""
    B = 512
    o, d, near, far = syn.make_rays(B, seed=21)
    true_rgb, mask = syn.make_targets(B, seed=22)
    m = build_modules(states, DEV, syn.RENDER_CONF_WMASK)
    R = m["renderer"]
    od, dd = o.to(DEV), d.to(DEV)
    out = R.render(od, dd, near.to(DEV), far.to(DEV), perturb_overwrite=0, cos_anneal_ratio=1.0)
    ref_e2e = O.render({k: v for k, v in states.items()}, o, d, near, far, conf=O.RENDER_CONF_WMASK, perturb_overwrite=0,
                       cos_anneal_ratio=1.0)
    for k in ("color_fine", "surface_color", "weight_sum", "gradient_error"):
        assert_close(out[k], ref_e2e[k].detach(), FP32_TOL, "512-ray render %s" % k)
    lin = torch.linspace(0.0, 1.0, 64, device=DEV)
    z = R._hierarchical(od, dd, ops.coarse_z(near.to(DEV), far.to(DEV), lin, None, 64))
    P = grad_params(states)
    ref = O.render(P, o, d, near, far, conf=O.RENDER_CONF_WMASK, perturb_overwrite=0, cos_anneal_ratio=1.0,
                   z_override=z.cpu())
    _stage1(ref, true_rgb, mask, 0.1).backward()
    core = R.render_core(od, dd, z, 2.0 / 64, m["sdf"], m["var"], m["color"], m["ref"], cos_anneal_ratio=1.0)
    for k, rk in (("color", "color_fine"), ("surface_color", "surface_color"), ("gradients", "gradients"),
                  ("weights", "weights"), ("cdf", "cdf_fine"), ("gradient_error", "gradient_error")):
        assert_close(core[k], ref[rk].detach(), FP32_TOL, "512-ray render_core %s" % k)
    assert torch.equal(core["sdf_mask"].cpu(), ref["sdf_mask"])
    w = core["weights"]
    outd = dict(color_fine=core["color"], surface_color=core["surface_color"], sdf_mask=core["sdf_mask"],
                weight_sum=w.sum(-1, keepdim=True), gradient_error=core["gradient_error"])
    _stage1(outd, true_rgb.to(DEV), mask.to(DEV), 0.1).backward()
    worst = compare_param_grads(m, P, ["sdf", "color", "var", "ref"], FP32_TOL, 1e-3, "512 rays")
    print("512-ray step: worst param-grad abs err %.3e" % worst)


# ------------------------------------------------------------------------------------------ section 8f: rays on device, stage-2 loss
def test_gen_rays_matches_dataset_formulas():
    """fneus_gen_rays against the reference expressions of Dataset.gen_random_rays_at / gen_rays_at / near_far_from_sphere
    (dataset.py:115-151,186-192) restated in torch."""
    H, W, B = 48, 64, 333
    g = torch.Generator().manual_seed(4)
    kinv, pose = syn.pinhole_camera(H, W, focal=90.0)
    rot = torch.linalg.qr(torch.randn(3, 3, generator=g))[0]
    pose[:3, :3] = rot
    pose[:3, 3] = torch.tensor([0.3, -0.2, 2.4])
    image, mask = torch.rand(H, W, 3, generator=g), (torch.rand(H, W, 3, generator=g) > 0.4).float()
    px = torch.randint(0, W, (B,), generator=g)
    py = torch.randint(0, H, (B,), generator=g)
    p = torch.stack([px, py, torch.ones_like(py)], dim=-1).float()
    p = torch.matmul(kinv[None, :3, :3], p[:, :, None]).squeeze()
    v = p / torch.linalg.norm(p, ord=2, dim=-1, keepdim=True)
    v = torch.matmul(pose[None, :3, :3], v[:, :, None]).squeeze()
    o = pose[None, :3, 3].expand(v.shape)
    ref = torch.cat([o, v, image[(py, px)], mask[(py, px)][:, :1]], dim=-1)
    a = torch.sum(v ** 2, dim=-1, keepdim=True)
    b = 2.0 * torch.sum(o * v, dim=-1, keepdim=True)
    mid = 0.5 * (-b) / a
    out, near, far = ops.gen_rays(px.float().to(DEV), py.float().to(DEV), kinv.to(DEV), pose.to(DEV), image.to(DEV), mask.to(DEV))
    assert_close(out, ref, 2e-6, "gen_rays rows")
    assert torch.equal(out[:, 6:].cpu(), ref[:, 6:])                       # colour / mask look-ups are exact
    assert_close(near, mid - 1.0, 3e-6, "near")
    assert_close(far, mid + 1.0, 3e-6, "far")
    # whole-image order of gen_rays_at (transposed meshgrid -> [H, W] row-major)
    qx, qy = syn.image_pixels(H, W, 1, DEV)
    full, _, _ = ops.gen_rays(qx, qy, kinv.to(DEV), pose.to(DEV), with_near_far=False)
    tx, ty = torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H)
    gx, gy = torch.meshgrid(tx, ty, indexing="ij")
    pp = torch.stack([gx, gy, torch.ones_like(gy)], dim=-1)
    pp = torch.matmul(kinv[None, None, :3, :3], pp[:, :, :, None]).squeeze()
    rv = pp / torch.linalg.norm(pp, ord=2, dim=-1, keepdim=True)
    rv = torch.matmul(pose[None, None, :3, :3], rv[:, :, :, None]).squeeze().transpose(0, 1)
    assert_close(full[:, 3:6].reshape(H, W, 3), rv, 2e-6, "gen_rays_at directions")


def test_stage2_loss_matches_oracle():
    """fneus_stage2_loss (lvis.py:163-170) vs the oracle restatement: value, parts and gradients."""
    B, k = 77, 4
    g = torch.Generator().manual_seed(6)
    hit = torch.rand(B, generator=g) > 0.3
    gt_l, gt_r = torch.rand(B, k, generator=g), torch.rand(B, k, 3, generator=g)
    pre_l = torch.rand(B, k, generator=g).requires_grad_(True)
    pre_r = torch.rand(B, k, 3, generator=g).requires_grad_(True)
    m1, m3 = hit[:, None], hit[:, None, None]
    out_o = dict(gt_lvis=torch.where(m1, gt_l, torch.ones(B, k)), pre_lvis=torch.where(m1, pre_l, torch.ones(B, k)),
                 gt_trace_radiance=torch.where(m3, gt_r, torch.ones(B, k, 3)),
                 pre_trace_radiance=torch.where(m3, pre_r, torch.ones(B, k, 3)), sdf_mask=hit)
    loss_o, st_o = O.stage2_loss(out_o)
    loss_o.backward()
    pl, pr = pre_l.detach().to(DEV).requires_grad_(True), pre_r.detach().to(DEV).requires_grad_(True)
    hd = hit.to(DEV)
    out_g = dict(gt_lvis=torch.where(hd[:, None], gt_l.to(DEV), torch.ones(B, k, device=DEV)),
                 pre_lvis=torch.where(hd[:, None], pl, torch.ones(B, k, device=DEV)),
                 gt_trace_radiance=torch.where(hd[:, None, None], gt_r.to(DEV), torch.ones(B, k, 3, device=DEV)),
                 pre_trace_radiance=torch.where(hd[:, None, None], pr, torch.ones(B, k, 3, device=DEV)), sdf_mask=hd)
    loss_g, st_g = ops.stage2_loss(out_g)
    loss_g.backward()
    assert abs(float(loss_g) - float(loss_o)) < 1e-6
    assert abs(float(st_g["lvis_loss"]) - float(st_o["lvis_loss"])) < 1e-6
    assert_close(pl.grad, pre_l.grad, 1e-7, "d pre_lvis")
    assert_close(pr.grad, pre_r.grad, 1e-7, "d pre_trace_radiance")


@pytest.mark.parametrize("B,n,kp,k", [(37, 64, 0, 16), (37, 64, 16, 16), (5, 112, 16, 16), (3, 40, 7, 9)])
def test_upsample_iter_is_the_three_kernels_in_one(B, n, kp, k):
    """fneus_upsample_iter (one launch per iteration of the hierarchical loop) against the separate kernels it fuses --
    merge_sorted, upsample_step, ray_points -- bit for bit, on rows with ties between old and new depths."""
    g = torch.Generator().manual_seed(100 * n + kp)
    o, d, near, far = syn.make_rays(B, seed=3)
    z = torch.sort(near + (far - near) * torch.rand(B, n, generator=g), dim=-1)[0]
    sdf = torch.randn(B, n, generator=g) * 0.3
    prev_z = prev_sdf = None
    if kp:
        prev_z = torch.sort(near + (far - near) * torch.rand(B, kp, generator=g), dim=-1)[0]
        prev_z[:, 0] = z[:, n // 2]                                   # a tie: the new sample goes AFTER the equal old one
        prev_z = torch.sort(prev_z, dim=-1)[0]
        prev_sdf = torch.randn(B, kp, generator=g) * 0.3
    u = torch.linspace(0.5 / k, 1.0 - 0.5 / k, k)
    cu = lambda t: None if t is None else t.to(DEV).contiguous()
    od, dd, zd, sd, pzd, psd, ud = map(cu, (o, d, z, sdf, prev_z, prev_sdf, u))
    if kp:
        z_ref, s_ref = ops.merge_sorted(zd, pzd, sd, psd)
    else:
        z_ref, s_ref = zd, sd
    nz_ref = ops.upsample_step(od, dd, z_ref, s_ref, k, 64.0, ud)
    p_ref = ops.ray_points(od, dd, nz_ref)
    z_got, s_got, nz_got, p_got = ops.upsample_iter(od, dd, zd, sd, pzd, psd, k, 64.0, ud)
    assert torch.equal(z_got, z_ref) and torch.equal(s_got, s_ref)
    assert torch.equal(nz_got, nz_ref)
    assert torch.equal(p_got.reshape(-1), p_ref.reshape(-1))
    assert ops.upsample_iter(od, dd, zd, sd, pzd, psd, k, 64.0, ud, want_pts=False)[3] is None
