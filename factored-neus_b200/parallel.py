"""Ray-sharded data parallelism for the stage-1 training step.

The reference is single-process (no torch.distributed call site anywhere); rays are independent through the
whole of ``NeuSRenderer.render`` and the only cross-ray reductions are the loss normalisers (mask_sum,
mask_sdf_sum: exp_runner.py:146,160) and the eikonal mean's denominator (renderer.py:372).  One process per
GPU: replicated weights, each rank renders its slice of the batch, the three denominators are all-reduced
(one tiny collective) so that every rank's loss is  local numerator / GLOBAL denominator, and the weight
gradients are summed with one all-reduce over a single flat FP32 bucket -- the result equals the
single-process gradient of the full batch.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn.functional as F


def shard_rays(n_rays: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of the global batch owned by ``rank``."""
    per = (n_rays + world - 1) // world
    lo = min(n_rays, rank * per)
    return lo, min(n_rays, lo + per)


def stage1_loss_sharded(renderer, out, true_rgb, mask, surface_weight=0.1, igr_weight=0.1, mask_weight=0.1,
                        group=None):
    """exp_runner.py:134-177 on a ray shard with batch-global normalisers.  Returns (loss_local, stats): the
    SUM over ranks of loss_local is the reference's full-batch loss."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    eik_num, eik_den = renderer.last_eikonal_parts
    if out["color_fine"].is_cuda:
        # product path: normalisers, then loss + gradients, one launch each (csrc/loss.cu)
        from . import ops
        hit_idx = renderer.last_hit_idx
        use_mask = mask_weight > 0.0
        den = ops.loss_norms(mask, hit_idx, eik_den.detach(), use_mask)
        if world > 1:
            dist.all_reduce(den, group=group)
        parts = ops.Stage1Loss.apply(out["color_fine"], out["surface_color"], out["weight_sum"], eik_num, true_rgb, mask,
                                     hit_idx, den, use_mask, surface_weight, igr_weight, mask_weight)
        stats = parts.detach()
        return parts[0], dict(color_loss=stats[1], surface_loss=stats[2], eikonal=stats[3], mask_loss=stats[4])
    # host-logic path (CPU tensors: the gloo world_size-2 test of the sharding arithmetic)
    mask = (mask > 0.5).to(true_rgb.dtype) if mask_weight > 0.0 else torch.ones_like(mask)
    hit = out["sdf_mask"]
    hit_f = hit.to(true_rgb.dtype)[:, None]
    n_local = mask.new_full((), float(mask.shape[0]))          # fill kernel: CUDA-graph capturable
    den = torch.stack([mask.sum(), (mask * hit_f).sum(), eik_den.detach(), n_local])
    if world > 1:
        dist.all_reduce(den, group=group)
    mask_sum, mask_sdf_sum, relax_sum, n_global = den[0] + 1e-5, den[1] + 1e-5, den[2] + 1e-5, den[3]
    color_loss = ((out["color_fine"] - true_rgb) * mask).abs().sum() / mask_sum
    surf_err = surface_weight * (out["surface_color"] - true_rgb) * mask * hit_f
    surf_loss = surf_err.abs().sum() / mask_sdf_sum
    eik_loss = eik_num / relax_sum
    ws = out["weight_sum"].clip(1e-3, 1.0 - 1e-3)
    mask_loss = F.binary_cross_entropy(ws, mask, reduction="sum") / n_global
    loss = color_loss + surf_loss + eik_loss * igr_weight + mask_loss * mask_weight
    return loss, dict(color_loss=color_loss, surface_loss=surf_loss, eikonal=eik_loss, mask_loss=mask_loss)


class GradBucket:
    """Single flat FP32 gradient bucket over a fixed parameter list; one all-reduce(sum) per step."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:          # parameters' .grad become views into the bucket: backward writes in place
            p.grad = self.flat[off: off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce(self):
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.flat, group=self.group)


class FlatAdam:
    """torch.optim.Adam(lr, betas=(0.9, 0.999), eps=1e-8) (exp_runner.py:118) + update_learning_rate
    (exp_runner.py:229-238) as ONE fused launch over flat buffers: the parameters are re-homed as views of one
    flat FP32 buffer (their values are kept), the gradients are the GradBucket's flat buffer, and the iteration
    counter / learning rate live on the device (``state``: [iterations done, last lr, 1-b1^t, 1-b2^t]), so a
    captured step needs no host-written scalar.  ``step()`` also clears the gradient bucket."""

    def __init__(self, bucket: "GradBucket", lr=5e-4, lr_alpha=0.05, warm_up_end=5000, end_iter=300000,
                 betas=(0.9, 0.999), eps=1e-8):
        self.bucket = bucket
        self.base_lr, self.lr_alpha, self.warm_up_end, self.end_iter = lr, lr_alpha, warm_up_end, end_iter
        self.betas, self.eps = betas, eps
        n = bucket.flat.numel()
        dev = bucket.flat.device
        self.flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in bucket.params:
                k = p.numel()
                self.flat_p[off: off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[off: off + k].view_as(p)
                off += k
        self.m = torch.zeros_like(self.flat_p)
        self.v = torch.zeros_like(self.flat_p)
        self.state = torch.zeros(4, dtype=torch.float32, device=dev)

    def set_iteration(self, it: int):
        """Resume support (exp_runner.py:266-275 restores iter_step)."""
        self.state[0] = float(it)

    def step(self):
        from . import ops
        ops.adam_step(self.flat_p, self.bucket.flat, self.m, self.v, self.state, self.base_lr, self.lr_alpha,
                      self.warm_up_end, self.end_iter, self.betas[0], self.betas[1], self.eps, 1.0, True)
