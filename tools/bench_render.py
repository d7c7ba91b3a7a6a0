#!/usr/bin/env python
"""Render-only throughput (BASELINE.json metric "render rays/s"): NeuSRenderer.render under no_grad, unperturbed, the
wmask configuration (64+64 samples, 4 up-sampling steps), as validate_image / render_novel_image call it
(exp_runner.py:399,503: chunks of rays).  Device time by CUDA events around a CUDA-graph replay.

  python tools/bench_render.py [rays_per_call ...]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import factored_neus_b200 as fn  # noqa: E402
from factored_neus_b200 import ops  # noqa: E402
from util import build_modules  # noqa: E402

syn = fn.synthetic
FLOP_PER_RAY_RENDER = 457_698_304          # 368 F_sdf + 128 F_col + 2 F_ref


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [512, 4096]
    dev = "cuda:0"
    ops.set_precision("bf16")
    m = build_modules(syn.scene_states(seed=4), dev, syn.RENDER_CONF_WMASK)
    R = m["renderer"]
    out_rows = []
    for B in sizes:
        o, d, near, far = [t.to(dev) for t in syn.make_rays(B, seed=1)]

        def call():
            with torch.no_grad():
                return R.render(o, d, near, far, perturb_overwrite=0, cos_anneal_ratio=1.0)["color_fine"]

        stream = torch.cuda.Stream()
        stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(stream):
            for _ in range(3):
                call()
        torch.cuda.current_stream().wait_stream(stream)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        reps = 5
        with torch.cuda.graph(graph, stream=stream):
            for _ in range(reps):
                call()
        graph.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        graph.replay()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        rps = B / (ms * 1e-3)
        out_rows.append({"rays_per_call": B, "ms_per_call": ms, "render_rays_per_s": rps,
                         "algorithmic_tflops": rps * FLOP_PER_RAY_RENDER / 1e12})
        print("render-only, %5d rays per call: %.3f ms -> %.0f rays/s (%.1f TFLOP/s algorithmic)" % (
            B, ms, rps, rps * FLOP_PER_RAY_RENDER / 1e12))
    print(json.dumps({"metric": "render_rays_per_s", "precision": "bf16", "rows": out_rows}))


if __name__ == "__main__":
    main()
