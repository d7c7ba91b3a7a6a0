#!/usr/bin/env python
"""Warm device times of the four chain kernels' callers on N points (GPU box): sdf forward+normal, sdf backward,
colour forward, colour backward.  Usage: python tools/chain_timing.py [N] [tag]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import factored_neus_b200 as fn  # noqa: E402

syn = fn.synthetic


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    tag = sys.argv[2] if len(sys.argv) > 2 else ""
    dev = "cuda:0"
    fn.ops.set_precision("bf16")
    lib = fn._lib.lib()
    sdf = fn.SDFNetwork(**syn.SDF_CONF)
    sdf.load_state_dict(syn.sdf_state(4, syn.SDF_CONF, 0.03))
    sdf = sdf.to(dev)
    col = fn.RenderingNetwork(**syn.COLOR_CONF)
    col.load_state_dict(syn.scene_states(seed=4)["color"])
    col = col.to(dev)
    x = (torch.rand(N, 3, device=dev) * 2 - 1)
    nrm = torch.nn.functional.normalize(torch.randn(N, 3, device=dev), dim=-1).requires_grad_(True)
    dirs = torch.nn.functional.normalize(torch.randn(N, 3, device=dev), dim=-1)
    feat = torch.randn(N, 256, device=dev).requires_grad_(True)
    best = [1e9] * 4
    lib.fneus_prof_enable(1)
    for it in range(12):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        e[0].record()
        s, f, n = sdf.value_feature_normal(x, want_normal=True)
        e[1].record()
        (s.sum() + f.sum() * 0.01 + n.sum()).backward()
        e[2].record()
        rgb = col(x, nrm, dirs, feat)
        e[3].record()
        rgb.sum().backward()
        e[4].record()
        torch.cuda.synchronize()
        if it >= 2:
            for i in range(4):
                best[i] = min(best[i], e[i].elapsed_time(e[i + 1]))
    print("%-28s sdf fwd %.3f  sdf bwd(+wgrad) %.3f  colour fwd %.3f  colour bwd(+wgrad) %.3f ms" % ((tag,) + tuple(best)))


if __name__ == "__main__":
    main()
