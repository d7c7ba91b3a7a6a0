#!/usr/bin/env python
"""Training throughput of the womask configuration (BASELINE.json configs[2]: outside NeRF with n_outside = 32, mask_weight 0)
on one GPU: render + stage-1 loss + backward + Adam through train.Stage1Trainer (whole step in one CUDA graph).

  python tools/bench_womask.py [--no-graph] [rays_per_step ...]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import factored_neus_b200 as fn  # noqa: E402
from factored_neus_b200 import ops  # noqa: E402
from factored_neus_b200.train import Stage1Trainer  # noqa: E402
from util import build_modules  # noqa: E402

syn = fn.synthetic
FLOP_PER_RAY = 1_718_092_800                # wmask step + 160 x 3 F_nerf


def main():
    use_graph = "--no-graph" not in sys.argv
    sizes = [int(a) for a in sys.argv[1:] if not a.startswith("--")] or [512, 4096]
    dev = "cuda:0"
    ops.set_precision("bf16")
    rows = []
    for B in sizes:
        m = build_modules(syn.scene_states(seed=4), dev, syn.RENDER_CONF_WOMASK)
        tr = Stage1Trainer(m["renderer"], [m["nerf"], m["sdf"], m["var"], m["color"], m["ref"]], B, warm_up_end=0,
                           end_iter=300000, mask_weight=0.0, use_graph=use_graph)
        o, d, _, _ = syn.make_rays(B, seed=1)
        rgb, mask = syn.make_targets(B, seed=2)
        batch = torch.cat([o, d, rgb, mask], dim=1).to(dev)
        for _ in range(5):                   # 3 eager warm-ups on the capture stream, capture, first replay
            loss = tr.step(batch)
        torch.cuda.synchronize()
        assert torch.isfinite(loss).all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        a.record()
        for _ in range(reps):
            tr.step(batch)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        rps = B / (ms * 1e-3)
        rows.append({"rays_per_step": B, "ms_per_step": ms, "train_rays_per_s": rps,
                     "algorithmic_tflops": rps * FLOP_PER_RAY / 1e12})
        print("womask train step, %5d rays: %.3f ms -> %.0f rays/s (%.1f TFLOP/s algorithmic)" % (
            B, ms, rps, rps * FLOP_PER_RAY / 1e12))
        del tr, m
        torch.cuda.empty_cache()
    print(json.dumps({"metric": "train_rays_per_s", "config": "womask n_outside=32", "precision": "bf16", "cuda_graph": use_graph,
                      "rows": rows}))


if __name__ == "__main__":
    main()
