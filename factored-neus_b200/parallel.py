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


# ---------------------------------------------------------------------------------------------------------------------
# Tile-sharded inference (BASELINE.json configs[4]): full-image rendering (exp_runner.py:374-426 renders an
# image as a Python loop over 512-ray chunks) and the marching-cubes grid query (renderer.py:14-29: 64^3 chunks) have no
# cross-ray / cross-voxel dependency: tiles go round-robin to the ranks, x-slabs of the grid in contiguous blocks, and one
# gather brings the results to rank 0.  No other collective.
# ---------------------------------------------------------------------------------------------------------------------
def shard_tiles(n_tiles: int, rank: int, world: int):
    """Tile indices owned by ``rank``: round-robin (neighbouring tiles of an image cost about the same, so every rank gets
    the same mix of empty and surface tiles)."""
    return list(range(rank, n_tiles, world))


def gather_tiles(local, n_items: int, tile: int, rank: int, world: int, group=None, dst: int = 0):
    """local [n_local_tiles * tile, C] (this rank's tiles in ``shard_tiles`` order, the last global tile possibly partial
    but padded to ``tile`` rows) -> on ``dst`` the full [n_items, C] tensor in item order, elsewhere None."""
    n_tiles = (n_items + tile - 1) // tile
    if world == 1:
        return local[:n_items]
    per = (n_tiles + world - 1) // world                  # equal-sized messages: pad ranks that own one tile fewer
    C = local.shape[1]
    buf = local.new_zeros(per * tile, C)
    buf[: local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, parts, dst=dst, group=group)
    if rank != dst:
        return None
    full = local.new_empty(n_tiles * tile, C)
    view = full.view(n_tiles, tile, C)
    for r in range(world):
        idx = shard_tiles(n_tiles, r, world)
        view[idx] = parts[r].view(per, tile, C)[: len(idx)]
    return full[:n_items]


def render_image_sharded(renderer, rays_o, rays_d, near=None, far=None, tile=4096, group=None,
                         keys=("color_fine", "normals")):
    """Full-image render (exp_runner.py:374-426 ``validate_image``): every rank renders its round-robin tiles of the ray
    list, rank 0 receives [N, 3] per key.  ``normals`` = sum_samples gradients * weights[:, :n] * inside_sphere as the
    reference accumulates them (exp_runner.py:418-424).  rays are [N,3] on this rank's device; N need not divide."""
    from . import ops
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    N = rays_o.shape[0]
    n_tiles = (N + tile - 1) // tile
    outs = {k: [] for k in keys}
    with torch.no_grad():
        for t in shard_tiles(n_tiles, rank, world):
            lo, hi = t * tile, min(N, (t + 1) * tile)
            o, d = rays_o[lo:hi].contiguous(), rays_d[lo:hi].contiguous()
            if near is None:
                nr, fr = ops.near_far_from_sphere(o, d)
            else:
                nr, fr = near[lo:hi], far[lo:hi]
            out = renderer.render(o, d, nr, fr, perturb_overwrite=0, cos_anneal_ratio=1.0)
            n = renderer.n_samples + renderer.n_importance
            for k in keys:
                if k == "normals":
                    v = (out["gradients"] * (out["weights"][:, :n] * out["inside_sphere"])[:, :, None]).sum(dim=1)
                else:
                    v = out[k]
                if hi - lo < tile:
                    v = torch.cat([v, v.new_zeros(tile - (hi - lo), v.shape[1])])
                outs[k].append(v)
    res = {}
    for k in keys:
        local = torch.cat(outs[k]) if outs[k] else rays_o.new_zeros(0, 3)
        res[k] = gather_tiles(local, N, tile, rank, world, group)
    return res


def extract_fields_sharded(renderer, bound_min, bound_max, resolution, group=None, dst: int = 0):
    """renderer.py:14-29 over all ranks: rank r evaluates the x-slab [r R / world, (r+1) R / world) of the R^3 grid on its
    GPU; rank ``dst`` receives the whole u [R,R,R] (device tensor), the others None."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_rays(resolution, rank, world)
    slab = renderer.extract_fields(bound_min, bound_max, resolution, ix0=lo, ix1=hi)
    if world == 1:
        return slab
    per = (resolution + world - 1) // world
    buf = slab.new_zeros(per, resolution, resolution)
    buf[: hi - lo] = slab
    parts = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, parts, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([parts[r][: shard_rays(resolution, r, world)[1] - shard_rays(resolution, r, world)[0]]
                      for r in range(world)])


class GradBucket:
    """Single flat FP32 gradient bucket over a fixed parameter list; the parameters' ``.grad`` are views into it.

    ``direct=True`` (default) additionally registers the parameters for direct accumulation: the weight-pack backward
    (``ops.PackWeights``) then adds into the bucket views itself and hands autograd no gradient tensors.  Tensor /
    post-accumulate hooks, ``torch.autograd.grad`` and DDP do not see those gradients; pass ``direct=False`` to keep plain
    autograd semantics (the ``.grad`` views still alias the bucket).

    ``segments``: optional list of parameter lists (a partition of ``params``) in the order their gradients become FINAL
    during backward.  The bucket is laid out segment by segment and ``all_reduce_segment(i)`` reduces segment i on a side
    stream as soon as the caller knows it is final, so that the transfer overlaps the rest of the backward pass
    (the colour / RefColor slice travels while the SDF backward runs); ``all_reduce()`` reduces whatever has
    not been reduced yet and joins the side stream.  ``params`` keeps its own order (optimizer state-dict order)."""

    def __init__(self, params, group=None, direct=True, segments=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.direct = direct
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        ids = {id(p) for p in self.params}
        if segments is None:
            segments = [self.params]
        segments = [[p for p in seg if id(p) in ids] for seg in segments]
        placed = [id(p) for seg in segments for p in seg]
        if sorted(placed) != sorted(ids):
            raise ValueError("GradBucket: `segments` must partition the trainable parameters")
        self.offsets, self.seg_span, off = {}, [], 0
        for seg in segments:
            lo = off
            for p in seg:          # parameters' .grad become views into the bucket: backward writes in place
                self.offsets[id(p)] = off
                p.grad = self.flat[off: off + p.numel()].view_as(p)
                p._fneus_direct_grad = bool(direct)
                off += p.numel()
            self.seg_span.append((lo, off))
        self._reduced = [False] * len(self.seg_span)
        self._side = None

    def offset_of(self, p):
        return self.offsets[id(p)]

    def zero(self):
        self.flat.zero_()

    def _world(self):
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def all_reduce_segment(self, i):
        """Segment i is final: reduce it on the side stream (fork from the current stream; ``all_reduce`` joins)."""
        if self._world() == 1 or self._reduced[i]:
            return
        lo, hi = self.seg_span[i]
        if hi == lo:
            self._reduced[i] = True
            return
        if self.flat.is_cuda:
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.flat.device)
            cur = torch.cuda.current_stream()
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                dist.all_reduce(self.flat[lo:hi], group=self.group)
        else:
            dist.all_reduce(self.flat[lo:hi], group=self.group)
        self._reduced[i] = True

    def all_reduce(self):
        """Reduce every segment not reduced yet (on the current stream) and wait for the side stream."""
        if self._world() > 1:
            pend = [sp for sp, done in zip(self.seg_span, self._reduced) if not done and sp[1] > sp[0]]
            if len(pend) == len(self.seg_span):
                dist.all_reduce(self.flat, group=self.group)
            else:
                for lo, hi in pend:
                    dist.all_reduce(self.flat[lo:hi], group=self.group)
            if self._side is not None and any(self._reduced):
                torch.cuda.current_stream().wait_stream(self._side)
        self._reduced = [False] * len(self.seg_span)


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
        with torch.no_grad():
            for p in bucket.params:                      # same placement as the gradient bucket
                k, off = p.numel(), bucket.offset_of(p)
                self.flat_p[off: off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[off: off + k].view_as(p)
        self.m = torch.zeros_like(self.flat_p)
        self.v = torch.zeros_like(self.flat_p)
        self.state = torch.zeros(4, dtype=torch.float32, device=dev)
        self._host_it = 0                   # host mirror of state[0] (graph replays advance it through step())
        self._moments_restored = False

    # ---- torch.optim.Adam surface used by the reference's Runner (exp_runner.py:179-181,196,237,261,273) ----
    def _lr_at(self, it):
        import math
        if it < self.warm_up_end:
            f = it / self.warm_up_end
        else:
            prog = (it - self.warm_up_end) / max(1.0, self.end_iter - self.warm_up_end)
            f = (math.cos(math.pi * prog) + 1.0) * 0.5 * (1 - self.lr_alpha) + self.lr_alpha
        return self.base_lr * f

    @property
    def param_groups(self):
        """One group, like the reference's ``Adam(params_to_train, lr)``.  ``lr`` is the value the NEXT step will use
        (host restatement of the device schedule; the device-resident schedule is authoritative, so writes to this
        dict by ``update_learning_rate`` (exp_runner.py:237) are accepted and ignored)."""
        return [{"lr": self._lr_at(self._host_it), "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0,
                 "amsgrad": False, "params": list(range(len(self.bucket.params)))}]

    def zero_grad(self, set_to_none=False):
        """The fused step clears the bucket itself; an explicit call clears it too (the ``.grad`` views stay)."""
        self.bucket.zero()

    def state_dict(self):
        """``torch.optim.Adam.state_dict()`` layout (per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq``), so that
        the reference's ``save_checkpoint`` / ``load_checkpoint`` (exp_runner.py:253-278) work unchanged and
        checkpoints move freely between the two implementations."""
        it = float(self._host_it)
        state = {}
        for i, p in enumerate(self.bucket.params):
            k, off = p.numel(), self.bucket.offset_of(p)
            if it > 0:
                state[i] = {"step": torch.tensor(it), "exp_avg": self.m[off: off + k].view_as(p).clone(),
                            "exp_avg_sq": self.v[off: off + k].view_as(p).clone()}
        return {"state": state, "param_groups": [dict(self.param_groups[0], foreach=None, maximize=False,
                                                      capturable=False, differentiable=False, fused=None)]}

    def load_state_dict(self, sd):
        """Accepts this class's own ``state_dict()`` and a reference ``torch.optim.Adam`` one over the same parameter
        order: restores both moments and the step counter (bias corrections and schedule follow from it)."""
        groups = sd["param_groups"]
        order = [i for g in groups for i in g["params"]]
        if len(order) != len(self.bucket.params):
            raise ValueError("FlatAdam.load_state_dict: %d parameters in the checkpoint, %d here"
                             % (len(order), len(self.bucket.params)))
        steps = set()
        with torch.no_grad():
            for idx, p in zip(order, self.bucket.params):
                k, off = p.numel(), self.bucket.offset_of(p)
                st = sd["state"].get(idx)
                if st is None:
                    self.m[off: off + k].zero_()
                    self.v[off: off + k].zero_()
                else:
                    if st["exp_avg"].numel() != k:
                        raise ValueError("FlatAdam.load_state_dict: parameter %d has %d elements, checkpoint %d"
                                         % (idx, k, st["exp_avg"].numel()))
                    self.m[off: off + k].copy_(st["exp_avg"].reshape(-1))
                    self.v[off: off + k].copy_(st["exp_avg_sq"].reshape(-1))
                    steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError("FlatAdam.load_state_dict: parameters disagree on the step count %s" % sorted(steps))
        self._set_it(steps.pop() if steps else 0)
        self._moments_restored = True

    def _set_it(self, it):
        self._host_it = int(it)
        self.state[0] = float(it)

    def set_iteration(self, it: int):
        """Move the schedule / bias-correction counter WITHOUT restoring the moments -- only meaningful at iteration 0
        or for benchmarking at a given learning rate.  Resuming a run must go through ``load_state_dict``: with zeroed
        moments and bias corrections already near 1 the first updates would be ~3x a normal Adam step."""
        if it > 0 and not self._moments_restored and self._host_it == 0:
            import warnings
            warnings.warn("FlatAdam.set_iteration(%d) with zeroed moments: use load_state_dict() to resume a run" % it)
        self._set_it(it)

    def step(self):
        from . import ops
        ops.adam_step(self.flat_p, self.bucket.flat, self.m, self.v, self.state, self.base_lr, self.lr_alpha,
                      self.warm_up_end, self.end_iter, self.betas[0], self.betas[1], self.eps, 1.0, True)
        self._host_it += 1
