"""Fused flat Adam + on-device learning-rate schedule (fneus_adam_step) against torch.optim.Adam driven by the
reference's update_learning_rate (exp_runner.py:118,229-238)."""
import math

import pytest
import torch

from factored_neus_b200.parallel import FlatAdam, GradBucket

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _lr_at(it, base, alpha, warm, end):
    if it < warm:
        f = it / warm
    else:
        f = (math.cos(math.pi * (it - warm) / (end - warm)) + 1.0) * 0.5 * (1 - alpha) + alpha
    return base * f


def test_flat_adam_matches_torch_adam():
    torch.manual_seed(0)
    shapes = [(256, 39), (256,), (217, 256), (1,), (3, 7, 5)]
    ours = [torch.nn.Parameter(torch.randn(s, device=DEV)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    base, alpha, warm, end = 5e-4, 0.05, 5, 40
    bucket = GradBucket(ours)
    opt = FlatAdam(bucket, lr=base, lr_alpha=alpha, warm_up_end=warm, end_iter=end)
    topt = torch.optim.Adam(ref, lr=base)
    for it in range(25):
        grads = [torch.randn(s, device=DEV) * (1.0 + it) for s in shapes]
        for p, q, g in zip(ours, ref, grads):
            p.grad.copy_(g)                       # the bucket views stay in place
            q.grad = g.clone()
        for grp in topt.param_groups:             # reference: rate of the iteration count BEFORE the step
            grp["lr"] = _lr_at(it, base, alpha, warm, end)
        opt.step()
        topt.step()
        assert float(bucket.flat.abs().max()) == 0.0          # the step clears the gradient bucket
        assert abs(float(opt.state[1]) - _lr_at(it, base, alpha, warm, end)) < 1e-9
    assert int(opt.state[0]) == 25
    for p, q in zip(ours, ref):
        assert p.data_ptr() >= opt.flat_p.data_ptr()          # parameters are views of the flat buffer
        assert float((p - q).abs().max()) < 2e-6


def test_flat_adam_resume_equals_uninterrupted_run():
    """a resume through state_dict()/load_state_dict() continues exactly (moments + counter), where
    set_iteration() alone would restart with zero moments."""
    torch.manual_seed(1)
    shapes = [(64, 39), (64,), (17, 64)]
    init = [torch.randn(s, device=DEV) for s in shapes]
    grads = [[torch.randn(s, device=DEV) for s in shapes] for _ in range(12)]

    def make():
        ps = [torch.nn.Parameter(t.clone()) for t in init]
        return ps, FlatAdam(GradBucket(ps), lr=5e-4, warm_up_end=4, end_iter=30)

    def run(ps, opt, its):
        for it in its:
            for p, g in zip(ps, grads[it]):
                p.grad.copy_(g)
            opt.step()

    pa, oa = make()
    run(pa, oa, range(12))
    pb, ob = make()
    run(pb, ob, range(7))
    sd = ob.state_dict()
    pc = [torch.nn.Parameter(p.detach().clone()) for p in pb]
    oc = FlatAdam(GradBucket(pc), lr=5e-4, warm_up_end=4, end_iter=30)
    oc.load_state_dict(sd)
    assert oc.param_groups[0]["lr"] == ob.param_groups[0]["lr"]
    run(pc, oc, range(7, 12))
    for a, c in zip(pa, pc):
        assert float((a - c).abs().max()) == 0.0


def test_pack_weights_returns_grads_to_autograd_without_bucket():
    """Without a GradBucket the weight-pack backward is an ordinary autograd node: torch.autograd.grad sees the
    gradients and .grad is not touched behind autograd's back."""
    import factored_neus_b200 as fn
    syn = fn.synthetic
    col = fn.RenderingNetwork(**syn.COLOR_CONF).to(DEV)
    for p in col.parameters():
        p.grad = torch.zeros_like(p)            # a pre-existing .grad must NOT switch direct mode on
    w = col.flat_weights()
    gs = torch.autograd.grad((w * w).sum(), list(col.parameters()))
    assert all(g is not None and float(g.abs().max()) > 0 for g in gs)
    assert all(float(p.grad.abs().max()) == 0.0 for p in col.parameters())
