#!/usr/bin/env python
"""HBM roofline of the bandwidth kernels (sampling and compositing) at a batch whose working set exceeds the
126 MB L2: B = 65 536 rays (at the benchmarked B = 512 these tensors are L2-resident and the launches are latency-bound).
Algorithmic bytes per ray as stated in DESIGN.md 4.3; time by CUDA events; peak from MEASURED_PEAKS.json.

  python tools/bench_bandwidth.py [B]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import factored_neus_b200 as fn  # noqa: E402
from factored_neus_b200 import ops  # noqa: E402

syn = fn.synthetic


def timeit(f, reps=10):
    """Device time per call: the calls are captured into one CUDA graph so that host launch overhead (ctypes + allocator,
    ~20 us per call) does not masquerade as kernel time."""
    stream = torch.cuda.Stream()
    stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(stream):
        for _ in range(3):
            f()
    torch.cuda.current_stream().wait_stream(stream)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=stream):
        for _ in range(reps):
            f()
    graph.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    graph.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


def timeit_eager(f, reps=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        f()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    dev = "cuda:0"
    peak = 6537.3
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    g = torch.Generator(device=dev).manual_seed(0)
    o, d, near, far = [t.to(dev) for t in syn.make_rays(B, seed=1)]
    rows = []

    def report(name, t, bytes_per_ray, note=""):
        gbs = B * bytes_per_ray / t / 1e9
        rows.append((name, t * 1e6, bytes_per_ray, gbs, gbs / peak))
        print("%-34s %9.1f us  %7d B/ray  %8.1f GB/s  %5.1f%% of %.0f GB/s %s" % (name, t * 1e6, bytes_per_ray, gbs,
                                                                                100 * gbs / peak, peak, note))

    # ---- up-sampling step (renderer.py:152-189 + 43-77) and sorted merge (191-205) for n = 64, 80, 96, 112 ----
    for n in (64, 80, 96, 112):
        k = 16
        z = (near + (far - near) * torch.linspace(0, 1, n, device=dev)[None, :]).contiguous()
        pts_r = torch.linalg.norm(o[:, None, :] + d[:, None, :] * z[:, :, None], dim=-1)
        sdf = (pts_r - 0.5).contiguous()
        u = torch.linspace(0.5 / k, 1 - 0.5 / k, k, device=dev)
        t = timeit(lambda: ops.upsample_step(o, d, z, sdf, k, 64.0, u))
        report("upsample_step n=%d k=16" % n, t, 8 * n + 4 * k + 24)
        new_z = ops.upsample_step(o, d, z, sdf, k, 64.0, u)
        new_sdf = torch.rand(B, k, device=dev, generator=g)
        t = timeit(lambda: ops.merge_sorted(z, new_z, sdf, new_sdf))
        report("merge_sorted n=%d k=16" % n, t, 16 * (n + k))
    # ---- compositing forward / backward (renderer.py:245-372), n = 128 ----
    n = 128
    z = (near + (far - near) * torch.linspace(0, 1, n, device=dev)[None, :]).contiguous()
    dists, mid_z, pts, dirs = ops.core_geometry(o, d, z, 2.0 / 64)
    t = timeit(lambda: ops.core_geometry(o, d, z, 2.0 / 64))
    report("core_geometry n=128", t, 4 * n + 24 + n * (4 + 4 + 12 + 12))
    sdf = (torch.linalg.norm(pts, dim=-1) - 0.5).contiguous().requires_grad_(True)
    nrm = torch.nn.functional.normalize(pts, dim=-1).contiguous().requires_grad_(True)
    rgb = torch.rand(B * n, 3, device=dev, generator=g).requires_grad_(True)
    inv_s = torch.full((1, 1), 20.0, device=dev, requires_grad=True)

    def fwd():
        return ops.Composite.apply(sdf, nrm, rgb, inv_s, None, None, dists, pts, d, None, n, 0, 1.0)

    with torch.no_grad():
        t = timeit(fwd)
    # reads sdf 4 + normal 12 + rgb 12 + dists 4 + pts 12 per sample, writes weights 4 + cdf 4 + inside 4 (+ per-ray I/O)
    report("composite_fwd n=128", t, n * (4 + 12 + 12 + 4 + 12 + 4 + 4 + 4) + 100)
    out = fwd()
    gc, gw = torch.rand_like(out[0]), torch.rand_like(out[1]) * 1e-3

    def bwd():
        torch.autograd.grad([out[0], out[1]], [sdf, nrm, rgb], [gc, gw], retain_graph=True)

    t = timeit_eager(bwd)
    # reads the forward's 44 B + upstream d_weights 4, writes d_sdf 4 + d_normal 12 + d_rgb 12 per sample
    report("composite_bwd n=128 (+torch sum)", t, n * (44 + 4 + 4 + 12 + 12) + 100)
    print(json.dumps({"rays": B, "hbm_peak_gbs": peak,
                      "kernels": [{"kernel": r[0], "us": r[1], "bytes_per_ray": r[2], "gbs": r[3], "frac": r[4]} for r in rows]}))


if __name__ == "__main__":
    main()
