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
