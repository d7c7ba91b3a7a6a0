"""CPU-side checks: the C-ABI library loads and exports every symbol include/fneus.h declares, host-only
queries work, and the drop-in modules keep the reference's parameter names/shapes.  No kernels are launched."""
import ctypes
import os
import re

import pytest
import torch

import factored_neus_b200 as fn
from factored_neus_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
syn = fn.synthetic


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "fneus.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fneus_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(L.LIB_PATH), "run __graft_entry__.build() first"
    h = ctypes.CDLL(L.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(h, s), "libfneus_b200.so does not export %s" % s
    assert set(L.declared_symbols()) == set(syms), "ctypes table and header disagree: %s" % (
        set(L.declared_symbols()) ^ set(syms))


def test_status_strings_and_version():
    lib = L.lib()
    assert lib.fneus_abi_version() >= 1
    assert lib.fneus_status_string(0) == b"ok"
    assert b"null" in lib.fneus_status_string(4)
    with pytest.raises(RuntimeError):
        L.check(3, "unit")


def test_pack_sizes_match_reference_parameter_counts():
    lib = L.lib()
    sdf = fn.SDFNetwork(**syn.SDF_CONF)
    col = fn.RenderingNetwork(**syn.COLOR_CONF)
    ref = fn.RefColor()
    # 524 544 weights + biases for the SDF net; colour 273 414 params (incl. weight_g);
    assert lib.fneus_sdf_pack_floats(sdf.cfg) == sdf.flat_weights().numel() == 524544 + 8 * 256 - 39 + 257
    assert lib.fneus_color_pack_floats(col.cfg) == col.flat_weights().numel()
    assert lib.fneus_ref_pack_floats(ref.cfg) == ref.flat_weights().numel() == 543492
    assert lib.fneus_sdf_saved_floats(sdf.cfg, 10) > 0 and lib.fneus_sdf_scratch_floats(sdf.cfg, 10) > 0
    bad = L.SdfCfg(d_in=3, d_hidden=256, n_layers=40, d_out=257, multires=6, skip_layer=4, scale=1.0, beta=100.0)
    assert lib.fneus_sdf_pack_floats(bad) == -1


def test_state_dict_keys_match_reference_layout():
    st = syn.scene_states(seed=4)
    sdf = fn.SDFNetwork(**syn.SDF_CONF)
    col = fn.RenderingNetwork(**syn.COLOR_CONF)
    ref = fn.RefColor()
    var = fn.SingleVarianceNetwork(0.3)
    for mod, key in ((sdf, "sdf"), (col, "color"), (ref, "ref"), (var, "var")):
        sd = mod.state_dict()
        assert set(sd.keys()) == set(st[key].keys()), key
        for k in sd:
            assert tuple(sd[k].shape) == tuple(st[key][k].shape), (key, k)
        mod.load_state_dict(st[key])
    assert sdf.lin3.weight_v.shape == (217, 256) and sdf.lin8.weight_g.shape == (257, 1)
    assert sum(p.numel() for p in sdf.parameters()) == 529076 + 0  # SURVEY.md 8a-2 (weights+biases) + weight_g


def test_geometric_init_follows_reference_statistics():
    torch.manual_seed(0)
    sdf = fn.SDFNetwork(**syn.SDF_CONF)
    assert float(sdf.lin0.weight_v[:, 3:].abs().max()) == 0.0
    assert float(sdf.lin4.weight_v[:, -36:].abs().max()) == 0.0
    assert abs(float(sdf.lin8.weight_v.mean()) - (3.14159265 ** 0.5) / 16.0) < 1e-3
    assert float(sdf.lin8.bias[0]) == -0.5
    w = sdf.lin2.effective()
    assert torch.allclose(w, sdf.lin2.weight_v, atol=1e-6)


def test_no_cpu_fallback():
    sdf = fn.SDFNetwork(**syn.SDF_CONF)
    with pytest.raises(RuntimeError):
        sdf.sdf(torch.zeros(4, 3))


def test_fanout_gather_rows_gradients_match_autograd():
    """ops.FanOut / ops.GatherRows (dense + sparse consumers of one activation): same gradients as plain autograd."""
    import torch
    from factored_neus_b200 import ops
    torch.manual_seed(0)
    x0 = torch.randn(50, 8, requires_grad=True)
    rows = torch.tensor([3, 3, 7, 49, 0, 12])
    w_dense, w_rows = torch.randn(50, 8), torch.randn(6, 8)

    def loss_plain(x):
        return (x * 2.0 * w_dense).sum() + (x.index_select(0, rows) ** 2 * w_rows).sum()

    def loss_fan(x, use_dense=True, use_sparse=True):
        stash = {}
        xd, xs = ops.FanOut.apply(x * 2.0, stash)
        out = x.sum() * 0.0
        if use_dense:
            out = out + (xd * w_dense).sum()
        if use_sparse:
            out = out + ((ops.GatherRows.apply(xs, rows, stash) / 2.0) ** 2 * w_rows).sum()
        return out

    g_ref, = torch.autograd.grad(loss_plain(x0), x0)
    g_fan, = torch.autograd.grad(loss_fan(x0), x0)
    assert torch.allclose(g_ref, g_fan, atol=1e-6)
    # only one of the two consumers contributes
    g_d, = torch.autograd.grad(loss_fan(x0, use_sparse=False), x0)
    assert torch.allclose(g_d, 2.0 * w_dense, atol=1e-6)
    g_s, = torch.autograd.grad(loss_fan(x0, use_dense=False), x0)
    g_s_ref, = torch.autograd.grad((x0.index_select(0, rows) ** 2 * w_rows).sum(), x0)
    assert torch.allclose(g_s, g_s_ref, atol=1e-6)


def test_flat_adam_state_dict_round_trips_with_torch_adam():
    """exp_runner.py:261,273: the reference saves / loads ``optimizer.state_dict()``.  FlatAdam speaks torch.optim.Adam's
    layout in both directions (host-side bookkeeping only: no kernel runs)."""
    from factored_neus_b200.parallel import FlatAdam, GradBucket
    torch.manual_seed(0)
    shapes = [(5, 3), (5,), (2, 2, 2)]
    ref = [torch.nn.Parameter(torch.randn(s)) for s in shapes]
    topt = torch.optim.Adam(ref, lr=5e-4)
    for _ in range(3):
        for p in ref:
            p.grad = torch.randn_like(p)
        topt.step()
    sd = topt.state_dict()
    ours = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    opt = FlatAdam(GradBucket(ours), lr=5e-4, warm_up_end=10, end_iter=100)
    assert opt.param_groups[0]["lr"] == 0.0 and len(opt.state_dict()["state"]) == 0
    opt.load_state_dict(sd)
    assert int(opt.state[0]) == 3 and opt._host_it == 3
    assert abs(opt.param_groups[0]["lr"] - 5e-4 * 3 / 10) < 1e-12
    back = opt.state_dict()
    for i, p in enumerate(ref):
        assert torch.equal(back["state"][i]["exp_avg"], sd["state"][i]["exp_avg"])
        assert torch.equal(back["state"][i]["exp_avg_sq"], sd["state"][i]["exp_avg_sq"])
        assert float(back["state"][i]["step"]) == 3.0
    topt2 = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in ref], lr=5e-4)
    topt2.load_state_dict(back)                       # and torch accepts ours
    assert float(topt2.state_dict()["state"][0]["step"]) == 3.0
    with pytest.raises(ValueError):
        FlatAdam(GradBucket(ours[:2])).load_state_dict(sd)
    with pytest.warns(UserWarning):
        FlatAdam(GradBucket([torch.nn.Parameter(torch.zeros(3))])).set_iteration(50)


def test_pack_weights_direct_mode_is_opt_in():
    """direct accumulation into .grad must be an explicit opt-in (GradBucket(direct=True)), never inferred
    from the mere presence of a .grad tensor."""
    from factored_neus_b200.parallel import GradBucket
    p = [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(4))]
    assert not getattr(p[0], "_fneus_direct_grad", False)
    GradBucket(p, direct=False)
    assert p[0].grad is not None and not p[0]._fneus_direct_grad
    GradBucket(p)
    assert p[0]._fneus_direct_grad and p[1]._fneus_direct_grad


def test_sqrt_free_radius_thresholds_are_exact():
    """csrc/fneus_common.cuh radius_lt_1 / radius_lt_1p2: `sqrtf(x) < c` (what the reference's `norm < c` does) equals
    `x < T(c)` with T(1) = 1 and T(1.2f) = 0x3FB851EC, checked on every float within 2^17 ulps of the thresholds."""
    import numpy as np
    for c, bits in ((1.0, 0x3F800000), (1.2, 0x3FB851EC)):
        t = np.uint32(bits).view(np.float32)
        a = np.arange(bits - (1 << 17), bits + (1 << 17), dtype=np.uint32).view(np.float32)
        assert np.array_equal(np.sqrt(a) < np.float32(c), a < t)
