#!/usr/bin/env python
"""Per-step timeline of the fused SDF chain kernels (CTA 0, first epilogue thread), from the clock stamps the kernel
records under debug flag bit 6.  Usage (GPU box):  python tools/timeline_sdf.py [N]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import factored_neus_b200 as fn  # noqa: E402

syn = fn.synthetic


def read_timeline(lib):
    buf = (ctypes.c_ulonglong * 8192)()
    fn._lib.check(lib.fneus_debug_timeline(ctypes.cast(buf, ctypes.c_void_p), 8192), "timeline")
    a = np.frombuffer(buf, dtype=np.uint64).copy()
    n = int(a[8191])
    return [(int(v >> np.uint64(56)), int(v & np.uint64((1 << 56) - 1))) for v in a[:n]]


def report(tag, tl, mhz=1965.0):
    print("== %s: %d stamps" % (tag, len(tl)))
    if not tl:
        return
    t0 = tl[0][1]
    if os.environ.get("TL_RAW"):
        nstep = 0
        for kind, t in tl:
            if kind == 1:
                nstep += 1
                base = t
            if nstep in (2, 11):
                print("   raw step %d  stamp %2d  +%.3f us" % (nstep - 1, kind, (t - base) / mhz))
    step, rows, cur = 0, [], None
    for kind, t in tl:
        us = (t - t0) / mhz
        if kind == 1:
            cur = {"start": us, "aux_wait": 0.0, "blk": 0.0, "nblk": 0}
        elif kind == 2:
            cur["acc"] = us
        elif kind == 3:
            cur["opfree"] = us
        elif kind == 4:
            cur["t4"] = us
        elif kind == 5:
            cur["aux_wait"] += us - cur["t4"]
            cur["t5"] = us
        elif kind in (8, 9, 10, 11):
            prev = cur.get("tp", cur["t5"]) if kind != 8 else cur["t5"]
            cur["ph%d" % kind] = cur.get("ph%d" % kind, 0.0) + us - prev
            cur["tp"] = us
        elif kind == 6:
            cur["blk"] += us - cur["t5"]
            cur["ph6"] = cur.get("ph6", 0.0) + us - cur.get("tp", cur["t5"])
            cur["nblk"] += 1
        elif kind == 7:
            cur["end"] = us
            rows.append(cur)
    print(" step   start   wait_mma  wait_opfree  wait_aux  compute   total | ldtm   math  pack+st fence arrive")
    for i, r in enumerate(rows[:24]):
        print("  %2d  %7.2f   %7.2f    %7.2f    %7.2f  %7.2f  %7.2f | %5.2f  %5.2f  %5.2f  %5.2f  %5.2f" % (
            i, r["start"], r["acc"] - r["start"], r["opfree"] - r["acc"], r["aux_wait"], r["blk"], r["end"] - r["start"],
            r.get("ph8", 0), r.get("ph9", 0), r.get("ph10", 0), r.get("ph11", 0), r.get("ph6", 0)))
    tot = rows[-1]["end"] - rows[0]["start"]
    print(" %d steps in %.1f us: wait_mma %.1f, wait_opfree %.1f, wait_aux %.1f, compute %.1f" % (
        len(rows), tot, sum(r["acc"] - r["start"] for r in rows), sum(r["opfree"] - r["acc"] for r in rows),
        sum(r["aux_wait"] for r in rows), sum(r["blk"] for r in rows)))


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    dev = "cuda:0"
    fn.ops.set_precision("bf16")
    lib = fn._lib.lib()
    sdf = fn.SDFNetwork(**syn.SDF_CONF)
    sdf.load_state_dict(syn.sdf_state(4, syn.SDF_CONF, 0.03))
    sdf = sdf.to(dev)
    x = (torch.rand(N, 3, device=dev) * 2 - 1)
    variants = [(0, "elected arrivals + suspend hint")]
    for xf, label in variants:
      print("#### variant:", label)
      for it in range(2):
        lib.fneus_debug_flags((64 if it == 1 else 0) | xf)
        s, f, n = sdf.value_feature_normal(x, want_normal=True)
        torch.cuda.synchronize()
        if it == 1:
            report("forward (value + gradient chain)", read_timeline(lib))
        (s.sum() + f.sum() * 0.01 + n.sum()).backward()
        torch.cuda.synchronize()
        if it == 1:
            report("backward (sweep + value path)", read_timeline(lib))
      lib.fneus_debug_flags(xf)
      for it in range(2):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        s, f, n = sdf.value_feature_normal(x, want_normal=True)
        e[1].record()
        (s.sum() + f.sum() * 0.01 + n.sum()).backward()
        e[2].record()
        torch.cuda.synchronize()
        print("sdf net [%s]: forward+normal %.3f ms, backward %.3f ms (incl. weight gradients)" % (
            label, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
    lib.fneus_debug_flags(0)
    # colour network (ReLU chain, <5, false> instantiation): debug flag bit 7
    col = fn.RenderingNetwork(**syn.COLOR_CONF)
    col.load_state_dict(syn.scene_states(seed=4)["color"])
    col = col.to(dev)
    nrm = torch.nn.functional.normalize(torch.randn(N, 3, device=dev), dim=-1).requires_grad_(True)
    dirs = torch.nn.functional.normalize(torch.randn(N, 3, device=dev), dim=-1)
    feat = torch.randn(N, 256, device=dev).requires_grad_(True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for it in range(3):
        lib.fneus_debug_flags(128 if it == 2 else 0)
        ev[0].record()
        rgb = col(x, nrm, dirs, feat)
        ev[1].record()
        torch.cuda.synchronize()
        if it == 2:
            report("colour forward", read_timeline(lib))
        rgb.sum().backward()
        ev[2].record()
        torch.cuda.synchronize()
        if it == 2:
            report("colour backward", read_timeline(lib))
        print("colour net: forward %.3f ms, backward %.3f ms (incl. weight gradients)" % (
            ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])))
    lib.fneus_debug_flags(0)
    # warm kernel times of the SDF chains
    for it in range(3):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        s, f, n = sdf.value_feature_normal(x, want_normal=True)
        e[1].record()
        (s.sum() + f.sum() * 0.01 + n.sum()).backward()
        e[2].record()
        torch.cuda.synchronize()
        print("sdf net: forward+normal %.3f ms, backward %.3f ms (incl. weight gradients)" % (
            e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))


if __name__ == "__main__":
    main()
