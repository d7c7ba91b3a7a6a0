"""Shared helpers for the parity tests."""
import numpy as np
import torch

import factored_neus_b200 as fn

syn = fn.synthetic


def build_modules(states, device, render_conf=None, sdf_conf=None, color_conf=None):
    """Product modules (CUDA path) loaded with the given reference-keyed state dicts."""
    sdf_conf = sdf_conf or syn.SDF_CONF
    color_conf = color_conf or syn.COLOR_CONF
    sdf = fn.SDFNetwork(**sdf_conf)
    col = fn.RenderingNetwork(**color_conf)
    var = fn.SingleVarianceNetwork(0.3)
    ref = fn.RefColor(d_feature=color_conf["d_feature"], d_hidden=color_conf["d_hidden"])
    sdf.load_state_dict(states["sdf"])
    col.load_state_dict(states["color"])
    var.load_state_dict(states["var"])
    ref.load_state_dict(states["ref"])
    mods = dict(sdf=sdf.to(device), color=col.to(device), var=var.to(device), ref=ref.to(device))
    if hasattr(fn, "NeRF"):
        nerf = fn.NeRF(**syn.NERF_CONF)
        nerf.load_state_dict(states["nerf"])
        mods["nerf"] = nerf.to(device)
    if render_conf is not None:
        mods["renderer"] = fn.NeuSRenderer(**render_conf, nerf=mods.get("nerf"), sdf_network=mods["sdf"],
                                           deviation_network=mods["var"], color_network=mods["color"],
                                           refColor_network=mods["ref"])
    return mods


def grad_params(states):
    """Oracle-side leaf copies with requires_grad."""
    return {k: {n: t.clone().requires_grad_(True) for n, t in sd.items()} for k, sd in states.items()}


def max_err(a, b):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max()) if a.size else 0.0


def assert_close(a, b, atol, what, rtol=0.0):
    a_ = a.detach().cpu().double() if torch.is_tensor(a) else torch.as_tensor(np.asarray(a)).double()
    b_ = b.detach().cpu().double() if torch.is_tensor(b) else torch.as_tensor(np.asarray(b)).double()
    assert a_.shape == b_.shape, "%s: shape %s vs %s" % (what, tuple(a_.shape), tuple(b_.shape))
    assert torch.isfinite(a_).all(), "%s: non-finite values" % what
    err = (a_ - b_).abs()
    tol = atol + rtol * b_.abs()
    bad = err > tol
    assert not bad.any(), "%s: max abs err %.3e (tol %.1e, rtol %.1e), %d/%d bad" % (
        what, float(err.max()), atol, rtol, int(bad.sum()), err.numel())


def compare_param_grads(mods, P, nets, atol, rtol, tag=""):
    """Product .grad (CUDA modules) vs oracle .grad (CPU leaf dicts), every parameter."""
    worst = 0.0
    for net in nets:
        for name, p in mods[net].named_parameters():
            ref = P[net][name].grad
            ref = torch.zeros_like(P[net][name]) if ref is None else ref
            got = p.grad if p.grad is not None else torch.zeros_like(p)
            assert_close(got, ref, atol, "%s grad %s.%s" % (tag, net, name), rtol=rtol)
            worst = max(worst, max_err(got, ref))
    return worst
