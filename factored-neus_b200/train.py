"""Stage-1 training step around the render path (reference: exp_runner.py:125-238), SURVEY.md section 8f-1.

``Stage1Trainer.step(batch)`` = render + stage-1 loss (exp_runner.py:134-177) + backward + flat-bucket
all-reduce (ray-sharded data parallelism) + fused Adam with the reference's warm-up / cosine schedule
(exp_runner.py:229-238).  All shapes are fixed, so the whole step can be captured into ONE CUDA graph
(``use_graph=True``): the iteration counter and learning rate live on the device and are advanced by the
optimiser launch itself (``parallel.FlatAdam`` -> ``fneus_adam_step``).
"""
from __future__ import annotations

import math

import torch

from .parallel import FlatAdam, GradBucket, stage1_loss_sharded


class Stage1Trainer:
    def __init__(self, renderer, networks, batch_size, lr=5e-4, lr_alpha=0.05, warm_up_end=5000, end_iter=300000,
                 igr_weight=0.1, mask_weight=0.1, surface_weight=0.1, anneal_end=0, use_white_bkgd=False,
                 use_graph=True):
        self.renderer = renderer
        self.networks = list(networks)
        self.batch_size = batch_size
        self.base_lr, self.lr_alpha, self.warm_up_end, self.end_iter = lr, lr_alpha, warm_up_end, end_iter
        self.igr_weight, self.mask_weight, self.surface_weight = igr_weight, mask_weight, surface_weight
        self.anneal_end = anneal_end
        self.use_white_bkgd = use_white_bkgd
        self.iter_step = 0
        params = [p for n in self.networks for p in n.parameters()]
        self.device = params[0].device
        # everything but the SDF network and the outside NeRF has its gradient complete when the SDF backward starts: that
        # slice is all-reduced on a side stream underneath it (parallel.GradBucket)
        sdf_net, nerf_net = renderer.sdf_network, renderer.nerf
        late = [p for n in self.networks if n is sdf_net or n is nerf_net for p in n.parameters()]
        late_ids = {id(p) for p in late}
        early = [p for p in params if id(p) not in late_ids]
        self.bucket = GradBucket(params, segments=[early, late])
        self.optimizer = FlatAdam(self.bucket, lr=lr, lr_alpha=lr_alpha, warm_up_end=warm_up_end, end_iter=end_iter)
        self.use_graph = use_graph
        self._graph = None
        self._static_batch = torch.zeros(batch_size, 10, device=self.device)
        self._static_loss = None
        self._stream = torch.cuda.Stream(device=self.device) if use_graph else None
        self._warm = 0
        self.last_stats = None
        # cos_anneal_ratio as a device scalar: written before every step, read by the compositing kernels, so that the
        # annealed (womask, anneal_end > 0) schedule runs inside the captured step too
        self._car_dev = torch.ones(1, dtype=torch.float32, device=self.device)

    # exp_runner.py:229-238
    def _lr_at(self, it):
        if it < self.warm_up_end:
            f = it / self.warm_up_end
        else:
            prog = (it - self.warm_up_end) / max(1, self.end_iter - self.warm_up_end)
            f = (math.cos(math.pi * prog) + 1.0) * 0.5 * (1 - self.lr_alpha) + self.lr_alpha
        return self.base_lr * f

    # exp_runner.py:223-227
    def cos_anneal_ratio(self):
        return 1.0 if self.anneal_end == 0 else min(1.0, self.iter_step / self.anneal_end)

    def _eager_step(self, batch):
        from . import ops
        ro, rd, rgb, m = ops.split_batch(batch)
        near, far = ops.near_far_from_sphere(ro, rd)                      # dataset.near_far_from_sphere
        bg = torch.ones([1, 3], device=batch.device) if self.use_white_bkgd else None
        out = self.renderer.render(ro, rd, near, far, background_rgb=bg, cos_anneal_ratio=self._car_dev)
        loss, stats = stage1_loss_sharded(self.renderer, out, rgb, m, self.surface_weight, self.igr_weight,
                                          self.mask_weight)
        from .ops import SdfValueGrad
        SdfValueGrad.pre_backward_hook = lambda: self.bucket.all_reduce_segment(0)
        try:
            loss.backward()                                              # the bucket is cleared by the optimiser step
        finally:
            SdfValueGrad.pre_backward_hook = None
        self.bucket.all_reduce()
        self.optimizer.step()
        return loss

    def step(self, batch):
        """batch [B,10] = (rays_o, rays_d, true_rgb, mask) as produced by Dataset.gen_random_rays_at
        (dataset.py:133-151).  Returns the (local-shard) loss tensor."""
        self._car_dev.fill_(self.cos_anneal_ratio())                      # outside the graph: replays read the new value
        if not self.use_graph:
            loss = self._eager_step(batch)
        else:
            cur = torch.cuda.current_stream()
            host_it = self.optimizer._host_it
            self._stream.wait_stream(cur)
            with torch.cuda.stream(self._stream):
                self._static_batch.copy_(batch, non_blocking=True)
                if self._warm < 3:                                         # warm-up on the capture stream
                    self._static_loss = self._eager_step(self._static_batch)
                    self._warm += 1
                else:
                    if self._graph is None:
                        self._graph = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(self._graph, stream=self._stream):
                            self._static_loss = self._eager_step(self._static_batch)
                    self._graph.replay()
            cur.wait_stream(self._stream)
            self.optimizer._host_it = host_it + 1       # one optimiser step ran, however it was issued (eager/capture+replay)
            loss = self._static_loss
        self.iter_step += 1
        return loss
