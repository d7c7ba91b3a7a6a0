#!/usr/bin/env python
"""bench.py -- the hot path of Factored-NeuS on B200: BASELINE.json's metric on BASELINE.json's configurations.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME] [--no-extras]

Headline (default ``--config wmask512``, BASELINE.json configs[1]): train rays/s of the wmask stage-1 step -- one "step" =
NeuSRenderer.render (perturbed, 64+64 samples, 4 up-sampling steps) + stage-1 loss (exp_runner.py:134-177) + backward
(incl. the SDF double backward) + Adam on 512 synthetic rays per GPU, networks at random geometric init ("sphere-SDF
scene").  N > 1: one process per GPU under torchrun, rays sharded (weak scaling), exact global loss
normalisers, one flat-bucket NCCL all-reduce.

The other configurations of BASELINE.json are measured in the same run on bounded samples and reported under
``extra_configs`` of the ONE JSON line (or as the headline with ``--config``):
  womask4096    configs[2]: womask step (outside NeRF, n_outside = 32), 4096 rays per GPU, ray-sharded
  render_image  configs[4]: render-only rays/s over tiles of a 1600x1200 synthetic camera, tiles round-robin over the GPUs
  grid512       configs[4]: 512^3 SDF grid query (extract_fields), x-slabs over the GPUs, gathered to rank 0
  lvis          configs[3]: stage-2 light-visibility trace, secondary rays/s (4 directions x 512 coarse + 32 fine samples)
  bandwidth     sampling / compositing kernels at 65 536 rays against the HBM roofline (single GPU)

``--impl reference`` times the reference algorithm's CPU restatement (oracle/neus_oracle.py; the reference is pure
PyTorch and /root/reference does not exist on the GPU box) on the host cores, same config and metric.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic FLOPs (2 x MACs of the dense layers; input gradient = 1x forward, backward = 2x)
F_SDF, F_SDF_ONLY, F_COL, F_NERF, F_REF = 1_049_088, 983_552, 542_720, 1_208_320, 1_082_880
FLOP_PER_RAY = {
    "wmask512": 880 * F_SDF + 384 * F_COL + 6 * F_REF,                       # 1 138 099 200
    "womask4096": 880 * F_SDF + 384 * F_COL + 6 * F_REF + 160 * 3 * F_NERF,  # 1 718 092 800
    "render_image": 368 * F_SDF + 128 * F_COL + 2 * F_REF,                   # 457 698 304
    "lvis": 610 * F_SDF,                                                     # ~0.640 GFLOP per secondary ray
}
SURFACE_W, IGR_W = 0.1, 0.1
CONFIGS = ("wmask512", "womask4096", "render_image", "grid512", "lvis", "bandwidth")
PROF_NAMES = ["gemm_fwd", "gemm_bwd_data", "gemm_wgrad", "sampling", "composite", "elementwise", "tc_gemm",
              "chain_sdf_fwd", "chain_sdf_bwd", "chain_relu", "tc_wgrad_group"]
TENSOR_CLASSES = (0, 1, 2, 6, 7, 8, 9, 10)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons, sampled every ~100 ms from before the warm-up (the first nvidia-smi call on a
    fresh box takes longer than a short timed region) and summarised over the samples taken between ``mark_begin`` and
    ``mark_end`` (the device-timed, per-kernel and end-to-end regions, which all run the same step)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.t0, self.t1 = None, None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def run(self):
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in o.strip().split(",")]
                if len(f) >= 6:
                    self.rows.append(f + [time.time()])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        rows = [r for r in self.rows if self.t0 is not None and self.t0 <= r[-1] <= (self.t1 or time.time())]
        window = "timed regions"
        if not rows:                       # region shorter than one nvidia-smi call: the samples closest to it
            mid = 0.5 * ((self.t0 or 0.0) + (self.t1 or time.time()))
            rows = sorted(self.rows, key=lambda r: abs(r[-1] - mid))[:3]
            window = "nearest samples"
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                "samples": len(rows), "window": window}


# =====================================================================================================================
# CPU arm (oracle port of the reference algorithm)
# =====================================================================================================================
def _oracle_setup(womask):
    import torch
    from oracle import neus_oracle as O
    import factored_neus_b200 as fn
    syn = fn.synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    st = syn.scene_states(seed=4, jitter=0.0)
    P = {k: {n: t.clone().requires_grad_(True) for n, t in sd.items()} for k, sd in st.items()}
    return torch, O, syn, cores, P


def cpu_step_fn(config, rays):
    """(step callable, units per step, unit name, sample description) of the CPU restatement for one configuration."""
    womask = config == "womask4096"
    torch, O, syn, cores, P = _oracle_setup(womask)
    if config in ("wmask512", "womask4096"):
        o, d, near, far = syn.make_rays(rays, seed=1)
        true_rgb, mask = syn.make_targets(rays, seed=2)
        conf = O.RENDER_CONF_WOMASK if womask else O.RENDER_CONF_WMASK

        def step():
            for sd in P.values():
                for t in sd.values():
                    t.grad = None
            out = O.render(P, o, d, near, far, conf=conf, cos_anneal_ratio=1.0)
            loss, _ = O.stage1_loss(out, true_rgb, mask, SURFACE_W, IGR_W, 0.0 if womask else 0.1)
            loss.backward()
            return float(loss)
        return step, rays, "rays", "%d synthetic rays/step (oracle port of renderer.py render + stage-1 loss + backward)" % rays
    if config == "render_image":
        o, d, near, far = syn.make_rays(rays, seed=1)

        def step():
            out = O.render(P, o, d, near, far, conf=O.RENDER_CONF_WMASK, perturb_overwrite=0, cos_anneal_ratio=1.0)
            return float(out["color_fine"].sum())
        return step, rays, "rays", "%d rays/step, render only (the reference builds the autograd graph for the normals too)" % rays
    if config == "grid512":
        R = 48

        def step():
            with torch.no_grad():
                u = O.extract_fields(P["sdf"], torch.tensor([-1.01] * 3), torch.tensor([1.01] * 3), R, chunk=64)
            return float(u.sum())
        return step, R ** 3, "voxels", "%d^3 grid (oracle extract_fields, 64^3 chunks)" % R
    if config == "lvis":
        m = rays
        import numpy as np
        rs = np.random.RandomState(11)
        surf = rs.standard_normal((m, 3))
        surf = torch.from_numpy((0.5 * surf / np.linalg.norm(surf, axis=1, keepdims=True)).astype(np.float32))
        normal = torch.nn.functional.normalize(surf, dim=-1)
        r_theta, rand_z = torch.rand(m, 4) * 6.2831853, torch.rand(m, 4) * 0.95
        Pd = {k: {n: t.detach() for n, t in sd.items()} for k, sd in P.items()}

        def step():
            lv, rad, _, _ = O.trace_visibility(Pd, surf, normal, r_theta, rand_z)
            return float(lv.sum())
        return step, 4 * m, "secondary rays", "%d surface points x 4 directions (oracle trace_visibility)" % m
    raise ValueError(config)


def cpu_baseline(config, rays, steps=2):
    step, units, unit, sample = cpu_step_fn(config, rays)
    best = None
    for i in range(steps + 1):
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if i > 0:
            best = dt if best is None else min(best, dt)
    return {"value": units / best, "unit": unit + "/s", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": sample + ", best of %d steps after 1 warm-up (CPU torch)" % steps}


CPU_SAMPLE = {"wmask512": 512, "womask4096": 256, "render_image": 512, "grid512": 0, "lvis": 64}
METRIC = {"wmask512": ("train_rays_per_s", "rays/s"), "womask4096": ("train_rays_per_s", "rays/s"),
          "render_image": ("render_rays_per_s", "rays/s"), "grid512": ("sdf_grid_voxels_per_s", "voxels/s"),
          "lvis": ("lvis_secondary_rays_per_s", "rays/s"), "bandwidth": ("hbm_gbs", "GB/s")}
WORKLOAD = {
    "wmask512": "wmask stage-1 train step (render fwd + loss + bwd incl. SDF double backward + Adam), 64+64 samples, 4 up-sample "
                "steps, sphere-SDF scene at geometric init",
    "womask4096": "womask stage-1 train step (outside NeRF, n_outside=32, mask_weight 0) + Adam, 64+64+32 samples",
    "render_image": "render only (no_grad, unperturbed), 4096-ray tiles of a 1600x1200 synthetic pinhole camera",
    "grid512": "extract_fields: -sdf on a 512^3 grid over [-1.01, 1.01]^3",
    "lvis": "stage-2 light-visibility trace: 4 directions per surface point, 512 coarse + 32 importance samples, first-hit shading",
    "bandwidth": "sampling / compositing kernels at 65 536 rays",
}


def run_reference(args):
    """CPU arm: the oracle port of the reference's algorithm for the configuration, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    config = args.config
    if config == "bandwidth":
        print(json.dumps({"impl": "reference", "unavailable": "the bandwidth kernels have no separate CPU counterpart"}))
        return
    rays = args.ref_rays if config in ("wmask512",) else CPU_SAMPLE[config]
    step, units, unit, sample = cpu_step_fn(config, rays)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = units * args.steps / dt
    metric, munit = METRIC[config]
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": v, "unit": munit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD[config], "units_per_step": units},
        "cpu_baseline": {"value": v, "unit": munit, "cores": os.cpu_count() or 1, "kind": "port",
                         "sample": sample + " x %d steps" % args.steps},
        "e2e": {"value": v, "unit": munit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# =====================================================================================================================
# GPU arm
# =====================================================================================================================
class Ctx:
    pass


def _prof_collect(lib, L, n):
    ncls = lib.fneus_prof_classes()
    ms = (ctypes.c_double * ncls)(); ln = (ctypes.c_longlong * ncls)()
    fl = (ctypes.c_double * ncls)(); by = (ctypes.c_double * ncls)()
    L.check(lib.fneus_prof_collect(ms, ln, fl, by), "prof_collect")
    out = {}
    for c in range(ncls):
        if ln[c]:
            t = ms[c] / n
            out[PROF_NAMES[c] if c < len(PROF_NAMES) else "class%d" % c] = {
                "ms_per_step": t, "launches_per_step": ln[c] / n, "gflop_per_step": fl[c] / n / 1e9,
                "designed_mb_per_step": by[c] / n / 1e6,
                "tflops": fl[c] / n / (t * 1e-3) / 1e12 if t > 0 else 0.0,
                "gbs": by[c] / n / (t * 1e-3) / 1e9 if t > 0 else 0.0}
    raw = {"ms": list(ms), "launches": list(ln), "flops": list(fl), "bytes": list(by)}
    return out, raw


def _profile_pass(C, fn_step, n):
    """Same work with eager launches bracketed by CUDA events inside the library: per-class time, launches, algorithmic
    FLOPs and designed DRAM bytes."""
    lib, L, torch = C.lib, C.L, C.torch
    lib.fneus_prof_collect(None, None, None, None)
    lib.fneus_prof_enable(1)
    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    C.barrier()
    ev[0].record()
    for _ in range(n):
        fn_step()
    ev[1].record()
    C.barrier()
    classes, raw = _prof_collect(lib, L, n)
    lib.fneus_prof_enable(0)
    return classes, raw, ev[0].elapsed_time(ev[1]) / n


def _tensor_roofline(C, classes, raw, n, dev_ms_per_step, flop_per_unit, units_per_s, kernel_note):
    pk = C.peaks
    t_ms = sum(raw["ms"][c] for c in TENSOR_CLASSES if c < len(raw["ms"]))
    t_fl = sum(raw["flops"][c] for c in TENSOR_CLASSES if c < len(raw["flops"]))
    achieved = t_fl / (t_ms * 1e-3) / 1e12 if t_ms > 0 else 0.0
    # the largest single launch of the step and the DRAM bytes it moves by design (counted by the library at launch)
    top = max((k for k in classes if k.startswith("chain_") or k.startswith("tc_")),
              key=lambda k: classes[k]["ms_per_step"] / max(1.0, classes[k]["launches_per_step"]), default=None)
    traffic = None
    if top is not None:
        traffic = classes[top]["designed_mb_per_step"] * 1e6 / max(1.0, classes[top]["launches_per_step"])
    return {"bound": "tensor", "achieved": achieved, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": achieved / pk["tf_sust"],
            "traffic": traffic,
            "traffic_source": ("designed DRAM bytes per launch of `%s`, counted by the library when it launches the kernel "
                               "(csrc/sdf_chain.cuh sdf_chain_bytes); ncu dram__bytes of the same kernels: profiles/r2_*" % top),
            "peak_source": pk["src"], "kernel": kernel_note,
            "kernel_share_of_step": (t_ms / n) / dev_ms_per_step if dev_ms_per_step > 0 else None,
            "measured": "CUDA events around every launch of %d eager steps" % n,
            "step_algorithmic_tflops": flop_per_unit * units_per_s / 1e12 if flop_per_unit else None}


def _build_nets(C, womask):
    fn, torch = C.fn, C.torch
    syn = fn.synthetic
    st = syn.scene_states(seed=4, jitter=0.0)
    sdf = fn.SDFNetwork(**syn.SDF_CONF); sdf.load_state_dict(st["sdf"])
    col = fn.RenderingNetwork(**syn.COLOR_CONF); col.load_state_dict(st["color"])
    var = fn.SingleVarianceNetwork(0.3); var.load_state_dict(st["var"])
    ref = fn.RefColor(); ref.load_state_dict(st["ref"])
    nets = [sdf.to(C.dev), var.to(C.dev), col.to(C.dev), ref.to(C.dev)]
    nerf = None
    if womask:
        nerf = fn.NeRF(**syn.NERF_CONF); nerf.load_state_dict(st["nerf"])
        nerf = nerf.to(C.dev)
        nets = [nerf] + nets                                                 # exp_runner.py:89-96: the NeRF comes first
    R = fn.NeuSRenderer(**(syn.RENDER_CONF_WOMASK if womask else syn.RENDER_CONF_WMASK), nerf=nerf, sdf_network=nets[-4],
                        deviation_network=nets[-3], color_network=nets[-2], refColor_network=nets[-1])
    return R, nets


def _stats(xs):
    xs = sorted(xs)
    q = lambda p: xs[min(len(xs) - 1, int(p * len(xs)))]
    return {"n": len(xs), "median_ms": q(0.5), "p10_ms": q(0.1), "p90_ms": q(0.9), "min_ms": xs[0], "max_ms": xs[-1]}


def bench_train(C, args, config, steps, stats_steps):
    """wmask512 / womask4096: whole training step, CUDA graph, device-timed K steps + >= 200-step statistics + per-kernel
    profile pass + end-to-end pass (pinned H2D of the rays, D2H of the loss)."""
    torch, dist, fn = C.torch, C.dist, C.fn
    from factored_neus_b200 import ops as _ops
    from factored_neus_b200.parallel import FlatAdam, GradBucket, stage1_loss_sharded
    syn = fn.synthetic
    womask = config == "womask4096"
    B = args.rays if (config == args.config and args.rays) else (4096 if womask else 512)
    mask_w = 0.0 if womask else 0.1
    R, nets = _build_nets(C, womask)
    params = [p for n in nets for p in n.parameters()]
    # gradient slices in the order they become final during backward: [variance, colour, RefColor] are complete when the SDF
    # backward starts and travel (side stream) while it runs; the SDF (and NeRF) slices follow at the end
    early = [p for n in nets[-3:] for p in n.parameters()]
    late = [p for n in nets[:-3] for p in n.parameters()]
    bucket = GradBucket(params, segments=[early, late])
    _ops.SdfValueGrad.pre_backward_hook = (lambda: bucket.all_reduce_segment(0)) if C.world > 1 else None
    opt = FlatAdam(bucket, lr=5e-4, warm_up_end=5000, end_iter=300000)
    opt._moments_restored = True                                             # benchmark at the full learning rate
    opt.set_iteration(5000)
    car = torch.ones(1, device=C.dev)                                        # device scalar: womask anneals it per iteration
    o, d, near, far = syn.make_rays(B, seed=1 + C.rank)
    true_rgb, mask = syn.make_targets(B, seed=100 + C.rank)
    host = torch.cat([o, d, true_rgb, mask], dim=1).pin_memory()             # [B,10] like dataset.gen_random_rays_at
    dev_batch = host.to(C.dev)

    def step(batch):
        ro, rd, rgb, m = _ops.split_batch(batch)                             # as train.Stage1Trainer._eager_step
        nr, fr = _ops.near_far_from_sphere(ro, rd)                           # dataset.near_far_from_sphere
        out = R.render(ro, rd, nr, fr, cos_anneal_ratio=car)
        loss, _ = stage1_loss_sharded(R, out, rgb, m, SURFACE_W, IGR_W, mask_w)
        loss.backward()
        bucket.all_reduce()
        opt.step()                                                           # also clears the gradient bucket
        return loss

    for _ in range(max(3, args.warmup)):
        step(dev_batch)
    torch.cuda.synchronize()
    graph, static_batch, static_loss = None, dev_batch.clone(), None
    if not args.no_graph:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=C.stream):
            static_loss = step(static_batch)
        graph.replay()
        torch.cuda.synchronize()
        if not torch.isfinite(static_loss).all():
            raise RuntimeError("non-finite loss from the graphed step")

    def run_step(batch):
        if graph is None:
            return step(batch)
        if batch is not static_batch:
            static_batch.copy_(batch, non_blocking=True)
        graph.replay()
        return static_loss

    def timed(n):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        C.barrier()
        for i in range(n):
            C.flush.zero_()                                                  # L2 flush between timed iterations
            ev[i][0].record()
            run_step(static_batch)
            ev[i][1].record()
        C.barrier()
        return [a.elapsed_time(b) for a, b in ev]

    per = timed(steps)
    dev_ms = sum(per)
    stat = _stats(timed(stats_steps)) if stats_steps else None
    classes, raw, prof_ms = _profile_pass(C, lambda: step(dev_batch), min(steps, 5))
    launches = int(sum(raw["launches"])) // min(steps, 5)
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()
    C.barrier()
    t_e2e = 0.0
    for i in range(steps):
        C.flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        batch = host.to(C.dev, non_blocking=True)                            # H2D of this step's rays [B,10]
        loss = run_step(batch)
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)         # D2H of the step's result
        torch.cuda.synchronize()
        t_e2e += time.perf_counter() - t0
    C.barrier()
    dev_ms, e2e_ms = C.max_over_ranks([dev_ms, t_e2e * 1e3])
    total = B * C.world * steps
    value = total / (dev_ms * 1e-3)
    res = {
        "metric": "train_rays_per_s", "value": value, "unit": "rays/s", "steps": steps, "ms_per_step": dev_ms / steps,
        "scaling": "weak",
        "config": {"workload": WORKLOAD[config], "rays_per_gpu_per_step": B, "global_rays_per_step": B * C.world,
                   "parallelism": "dp%d" % C.world,
                   "l2": "256 MiB flush between timed iterations; per-step working set >> 126 MB L2",
                   "cuda_graph": graph is not None,
                   "precision_path": ("fp16 forward / bf16 backward operands on tcgen05, fp32 accumulate" if args.precision == "bf16"
                                      else "fp32-simt")},
        "e2e": {"value": total / (e2e_ms * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": host.numel() * 4 * C.world,
                "d2h_bytes_per_step": 4 * C.world, "ms_per_step": e2e_ms / steps},
        "gpu_launches": launches * steps,
        "roofline": _tensor_roofline(C, classes, raw, min(steps, 5), dev_ms / steps, FLOP_PER_RAY[config], value / C.world,
                                     "sdf_chain_kernel<SDF fwd|SDF bwd|ReLU> + tc_gemm_wgrad_group_kernel (tcgen05 dense MLP "
                                     "layers)" if args.precision == "bf16" else "gemm_mk_kernel/gemm_wgrad_kernel (dense MLP layers)"),
        "kernel_classes": classes,
    }
    if stat:
        stat["rays_per_s_at_median"] = B * C.world / (stat["median_ms"] * 1e-3)
        res["step_time_stats"] = stat
    _ops.SdfValueGrad.pre_backward_hook = None
    del graph
    return res


def bench_render(C, args, steps):
    """render_image: a bounded sample of the 1600x1200 frame (4096-ray tiles, taken across the whole image), tiles
    round-robin over the ranks, colours + normals gathered to rank 0 (strong scaling: the sample is the same at every N)."""
    torch, fn = C.torch, C.fn
    from factored_neus_b200 import ops as _ops
    from factored_neus_b200.parallel import render_image_sharded
    syn = fn.synthetic
    H, W, tile = 1200, 1600, 4096
    n_tiles_img = (H * W + tile - 1) // tile
    n_tiles = args.render_tiles
    R, _ = _build_nets(C, False)
    kinv, pose = syn.pinhole_camera(H, W)
    px, py = syn.image_pixels(H, W, 1, C.dev)
    # sample: every (n_tiles_img // n_tiles)-th tile of the frame, so that empty and surface regions are represented
    pick = torch.arange(n_tiles, device=C.dev) * (n_tiles_img // n_tiles)
    idx = (pick[:, None] * tile + torch.arange(tile, device=C.dev)[None, :]).reshape(-1).clamp_(max=H * W - 1)
    rays, near, far = _ops.gen_rays(px[idx], py[idx], kinv.to(C.dev), pose.to(C.dev))
    o, d = rays[:, :3].contiguous(), rays[:, 3:6].contiguous()
    N = o.shape[0]

    def run():
        return render_image_sharded(R, o, d, near, far, tile=tile)

    for _ in range(2):
        run()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    C.barrier()
    for i in range(steps):
        ev[i][0].record()
        out = run()
        ev[i][1].record()
    C.barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    classes, raw, _ = _profile_pass(C, run, 1)
    # end to end: rays from pinned host memory, colours + normals back to the host on rank 0
    host_rays = torch.cat([o, d], 1).cpu().pin_memory()
    C.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        hr = host_rays.to(C.dev, non_blocking=True)
        out = render_image_sharded(R, hr[:, :3].contiguous(), hr[:, 3:6].contiguous(), None, None, tile=tile)
        if C.rank == 0:
            img = torch.cat([out["color_fine"], out["normals"]], 1).cpu()
        torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    dev_ms, e2e_ms = C.max_over_ranks([dev_ms, e2e_ms])
    value = N * steps / (dev_ms * 1e-3)
    return {"metric": "render_rays_per_s", "value": value, "unit": "rays/s", "steps": steps, "ms_per_step": dev_ms / steps,
            "scaling": "strong",
            "config": {"workload": WORKLOAD["render_image"], "rays_per_step": N, "tiles_per_step": n_tiles, "tile": tile,
                       "frame_rays": H * W, "frame_seconds_at_this_rate": H * W / value, "sharding": "tiles round-robin over %d GPUs, "
                       "one gather to rank 0" % C.world},
            "e2e": {"value": N * steps / (e2e_ms * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": N * 24, "d2h_bytes_per_step": N * 24},
            "gpu_launches": int(sum(raw["launches"])) * steps,
            "roofline": _tensor_roofline(C, classes, raw, 1, dev_ms / steps, FLOP_PER_RAY["render_image"], value / C.world,
                                         "sdf_chain_kernel<SDF fwd> (5 passes) + sdf_chain_kernel<ReLU>"),
            "kernel_classes": classes}


def bench_grid(C, args, steps):
    """grid512: extract_fields at R = 512 (134 M voxels), x-slabs over the ranks, gathered to rank 0."""
    torch, fn = C.torch, C.fn
    from factored_neus_b200.parallel import extract_fields_sharded
    R, _ = _build_nets(C, False)
    res_ = args.grid_resolution
    bmin, bmax = torch.tensor([-1.01] * 3), torch.tensor([1.01] * 3)

    def run():
        return extract_fields_sharded(R, bmin, bmax, res_)

    run()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    C.barrier()
    for i in range(steps):
        ev[i][0].record()
        u = run()
        ev[i][1].record()
    C.barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    del u
    classes, raw, _ = _profile_pass(C, run, 1)
    C.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        u = run()
        if C.rank == 0:
            uh = u.cpu()                                                     # what marching cubes consumes (renderer.py:27)
        torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    dev_ms, e2e_ms = C.max_over_ranks([dev_ms, e2e_ms])
    vox = res_ ** 3
    value = vox * steps / (dev_ms * 1e-3)
    return {"metric": "sdf_grid_voxels_per_s", "value": value, "unit": "voxels/s", "steps": steps, "ms_per_step": dev_ms / steps,
            "scaling": "strong",
            "config": {"workload": WORKLOAD["grid512"].replace("512", str(res_)), "resolution": res_, "voxels": vox,
                       "sharding": "x-slabs over %d GPUs, one gather to rank 0" % C.world},
            "e2e": {"value": vox * steps / (e2e_ms * 1e-3), "unit": "voxels/s", "h2d_bytes_per_step": 24,
                    "d2h_bytes_per_step": vox * 4},
            "gpu_launches": int(sum(raw["launches"])) * steps,
            "roofline": _tensor_roofline(C, classes, raw, 1, dev_ms / steps, F_SDF_ONLY, value / C.world,
                                         "sdf_chain_kernel<SDF fwd> (sdf-only value chain)"),
            "kernel_classes": classes}


def bench_lvis(C, args, steps):
    """lvis: calLvis.cal_indiLgt ground truth on surface points of the init sphere, points sharded over the ranks (no
    collective)."""
    torch, fn = C.torch, C.fn
    import numpy as np
    from factored_neus_b200 import lvis as LV
    R, nets = _build_nets(C, False)
    m = args.lvis_points
    rs = np.random.RandomState(11 + C.rank)
    surf = rs.standard_normal((m, 3))
    surf = torch.from_numpy((0.5 * surf / np.linalg.norm(surf, axis=1, keepdims=True)).astype(np.float32))
    normal = torch.nn.functional.normalize(surf, dim=-1)
    g = torch.Generator().manual_seed(3)
    r_theta, rand_z = torch.rand(m, 4, generator=g) * 6.2831853, torch.rand(m, 4, generator=g) * 0.95
    host = torch.cat([surf, normal, r_theta, rand_z], 1).pin_memory()
    dv = host.to(C.dev)

    def run(t):
        return LV.trace_visibility(t[:, :3].contiguous(), t[:, 3:6].contiguous(), nets[0], nets[1], nets[2],
                                   t[:, 6:10].contiguous(), t[:, 10:14].contiguous())

    run(dv)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    C.barrier()
    for i in range(steps):
        ev[i][0].record()
        run(dv)
        ev[i][1].record()
    C.barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    classes, raw, _ = _profile_pass(C, lambda: run(dv), 1)
    C.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        lv, rad, _ = run(host.to(C.dev, non_blocking=True))
        res = torch.cat([lv.reshape(m, -1), rad.reshape(m, -1)], 1).cpu()
        torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    dev_ms, e2e_ms = C.max_over_ranks([dev_ms, e2e_ms])
    rays = 4 * m * C.world
    value = rays * steps / (dev_ms * 1e-3)
    return {"metric": "lvis_secondary_rays_per_s", "value": value, "unit": "rays/s", "steps": steps, "ms_per_step": dev_ms / steps,
            "scaling": "weak",
            "config": {"workload": WORKLOAD["lvis"], "surface_points_per_gpu": m, "n_dirs": 4, "n_coarse": 512, "n_importance": 32,
                       "sdf_evaluations_per_s": value * 610, "sharding": "surface points over %d GPUs, no collective" % C.world},
            "e2e": {"value": rays * steps / (e2e_ms * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": host.numel() * 4 * C.world,
                    "d2h_bytes_per_step": m * 16 * 4 * C.world},
            "gpu_launches": int(sum(raw["launches"])) * steps,
            "roofline": _tensor_roofline(C, classes, raw, 1, dev_ms / steps, FLOP_PER_RAY["lvis"], value / C.world,
                                         "sdf_chain_kernel<SDF fwd> (512 coarse sdf-only evaluations per ray dominate)"),
            "kernel_classes": classes}


def bench_bandwidth(C, args):
    """Sampling / compositing kernels at 65 536 rays (working set > L2) against the measured HBM peak; algorithmic bytes per
    ray as stated in DESIGN.md.  Each kernel is timed by CUDA events around a graph of 10 launches."""
    torch, fn = C.torch, C.fn
    from factored_neus_b200 import ops
    syn = fn.synthetic
    B, dev, peak = args.bw_rays, C.dev, C.peaks["hbm"]
    g = torch.Generator(device=dev).manual_seed(0)
    o, d, near, far = [t.to(dev) for t in syn.make_rays(B, seed=1)]
    rows = []

    def timeit(f, reps=10):
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
        a.record(); graph.replay(); b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps * 1e-3

    def report(name, t, bytes_per_ray):
        gbs = B * bytes_per_ray / t / 1e9
        rows.append({"kernel": name, "us": t * 1e6, "bytes_per_ray": bytes_per_ray, "gbs": gbs, "frac": gbs / peak})

    for n in (64, 112):
        k = 16
        z = (near + (far - near) * torch.linspace(0, 1, n, device=dev)[None, :]).contiguous()
        sdf = (torch.linalg.norm(o[:, None, :] + d[:, None, :] * z[:, :, None], dim=-1) - 0.5).contiguous()
        u = torch.linspace(0.5 / k, 1 - 0.5 / k, k, device=dev)
        report("upsample_step n=%d k=16" % n, timeit(lambda: ops.upsample_step(o, d, z, sdf, k, 64.0, u)), 8 * n + 4 * k + 24)
        new_z = ops.upsample_step(o, d, z, sdf, k, 64.0, u)
        new_sdf = torch.rand(B, k, device=dev, generator=g)
        report("merge_sorted n=%d k=16" % n, timeit(lambda: ops.merge_sorted(z, new_z, sdf, new_sdf)), 16 * (n + k))
    n = 128
    z = (near + (far - near) * torch.linspace(0, 1, n, device=dev)[None, :]).contiguous()
    dists, mid_z, pts, dirs = ops.core_geometry(o, d, z, 2.0 / 64)
    report("core_geometry n=128", timeit(lambda: ops.core_geometry(o, d, z, 2.0 / 64)), 4 * n + 24 + n * (4 + 4 + 12 + 12))
    sdf = (torch.linalg.norm(pts, dim=-1) - 0.5).contiguous().requires_grad_(True)
    nrm = torch.nn.functional.normalize(pts, dim=-1).contiguous().requires_grad_(True)
    rgb = torch.rand(B * n, 3, device=dev, generator=g).requires_grad_(True)
    inv_s = torch.full((1, 1), 20.0, device=dev, requires_grad=True)
    fwd = lambda: ops.Composite.apply(sdf, nrm, rgb, inv_s, None, None, dists, pts, d, None, n, 0, 1.0)
    with torch.no_grad():
        report("composite_fwd n=128", timeit(fwd), n * (4 + 12 + 12 + 4 + 12 + 4 + 4 + 4) + 100)
    out = fwd()
    gc, gw = torch.rand_like(out[0]), torch.rand_like(out[1]) * 1e-3
    bwd = lambda: torch.autograd.grad([out[0], out[1]], [sdf, nrm, rgb, inv_s], [gc, gw], retain_graph=True)
    for _ in range(3):
        bwd()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        bwd()
    b.record()
    torch.cuda.synchronize()
    report("composite_bwd n=128", a.elapsed_time(b) / 10 * 1e-3, n * (44 + 4 + 4 + 12 + 12) + 100)
    worst = min(rows, key=lambda r: r["frac"])
    tot_bytes = sum(r["bytes_per_ray"] for r in rows) * B
    tot_t = sum(r["us"] for r in rows) * 1e-6
    return {"metric": "hbm_gbs", "value": tot_bytes / tot_t / 1e9, "unit": "GB/s", "steps": 10, "ms_per_step": tot_t * 1e3,
            "scaling": "replicas only",
            "config": {"workload": WORKLOAD["bandwidth"], "rays": B},
            "roofline": {"bound": "hbm", "achieved": tot_bytes / tot_t / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": tot_bytes / tot_t / 1e9 / peak, "traffic": None, "peak_source": C.peaks["src"],
                         "kernel": "all sampling / compositing kernels, bytes-weighted; worst: %s (%.0f%%)" % (
                             worst["kernel"], 100 * worst["frac"])},
            "kernels": rows}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import factored_neus_b200 as fn
    from factored_neus_b200 import _lib as L
    from factored_neus_b200 import ops as _ops

    C = Ctx()
    C.torch, C.dist, C.fn, C.L = torch, dist, fn, L
    C.world = int(os.environ.get("WORLD_SIZE", "1"))
    C.rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # libraries (NCCL's version banner) write to file descriptor 1: keep stdout for the ONE JSON line
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    C.dev = torch.device("cuda", local)
    if C.world > 1:
        dist.init_process_group("nccl", device_id=C.dev)
    C.peaks = _peaks()
    C.lib = L.lib()
    _ops.set_precision(args.precision)
    if args.debug_flags:
        C.lib.fneus_debug_flags(args.debug_flags)
    C.flush = torch.empty(256 << 20, dtype=torch.uint8, device=C.dev)        # > 126 MB L2

    def barrier():
        if C.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        t = torch.tensor(vals, device=C.dev, dtype=torch.float64)
        if C.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    C.barrier, C.max_over_ranks = barrier, max_over_ranks
    sampler = ClockSampler(local)
    sampler.start()
    # every step (warm-up, capture, replay, eager) runs on one non-default stream so that autograd's AccumulateGrad nodes
    # and the flat gradient bucket live on the capture stream
    C.stream = torch.cuda.Stream()
    C.stream.wait_stream(torch.cuda.current_stream())
    torch.cuda.set_stream(C.stream)

    def run_config(name, headline):
        steps = args.steps if headline else min(args.steps, args.extra_steps)
        if name in ("wmask512", "womask4096"):
            return bench_train(C, args, name, steps, args.stats_steps if headline else 0)
        if name == "render_image":
            return bench_render(C, args, max(1, min(steps, 3)) if not headline else steps)
        if name == "grid512":
            return bench_grid(C, args, max(1, min(steps, 2)) if not headline else steps)
        if name == "lvis":
            return bench_lvis(C, args, max(1, min(steps, 2)) if not headline else steps)
        if name == "bandwidth":
            return bench_bandwidth(C, args)
        raise ValueError(name)

    sampler.mark_begin()
    try:
        head = run_config(args.config, True)
    except Exception as ex:
        if args.config in ("wmask512", "womask4096") and not args.no_graph and C.world == 1:
            # a failed capture leaves the CUDA RNG in capture mode: re-run this process eagerly instead
            print("bench: graphed step failed (%s%s); re-running with --no-graph" % (str(ex).splitlines()[0], _hang_note()),
                  file=sys.stderr)
            sys.stdout.flush()
            os.dup2(real_stdout, 1)
            os.execv(sys.executable, [sys.executable] + sys.argv + ["--no-graph"])
        raise
    extras = {}
    if not args.no_extras:
        for name in CONFIGS:
            if name == args.config or (name == "bandwidth" and C.world > 1):
                continue
            torch.cuda.empty_cache()
            try:
                extras[name] = run_config(name, False)
            except Exception as ex:                                          # an extra must never cost the headline line
                extras[name] = {"error": (str(ex).splitlines()[0][:300] + _hang_note())[:2000]}
    sampler.mark_end()
    sampler.stop_flag = True

    if C.rank == 0:
        metric, unit = METRIC[args.config]
        line = {"metric": metric, "value": head["value"], "unit": unit, "n_gpus": C.world, "steps": head["steps"],
                "warmup": max(3, args.warmup), "ms_per_step": head["ms_per_step"], "higher_is_better": True,
                "scaling": head["scaling"], "vs_baseline": None,
                "dtype": "fp16+bf16" if args.precision == "bf16" else "f32", "data": "synthetic"}
        for k in ("config", "e2e", "gpu_launches", "roofline", "kernel_classes", "step_time_stats", "kernels"):
            if k in head:
                line[k] = head[k]
        line["clocks"] = sampler.summary()
        if extras:
            line["extra_configs"] = extras
        if C.world == 1 and not args.no_cpu_baseline and args.config != "bandwidth":
            line["cpu_baseline"] = cpu_baseline(args.config, args.cpu_rays if args.config == "wmask512" else CPU_SAMPLE[args.config])
            if not args.no_extras:
                for name in extras:
                    if name in CPU_SAMPLE and "error" not in extras[name]:
                        try:
                            extras[name]["cpu_baseline"] = cpu_baseline(name, CPU_SAMPLE[name], steps=1)
                        except Exception as ex:
                            extras[name]["cpu_baseline"] = {"error": str(ex).splitlines()[0][:200]}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line))
        sys.stdout.flush()
    if C.world > 1:
        # NCCL communicators captured inside a CUDA graph can stall destroy_process_group(): the result is out,
        # so flush and leave without the collective teardown
        sys.stdout.flush()
        sys.stderr.flush()
        torch.cuda.synchronize()
        os._exit(0)


def _hang_note():
    """The library's wait watchdog record, if a kernel trapped on a wait that never completed (csrc/api.cu)."""
    try:
        import factored_neus_b200 as fn
        rec = (ctypes.c_ulonglong * 64)()
        fn._lib.lib().fneus_debug_hang_record(ctypes.cast(rec, ctypes.c_void_p))
        if rec[0]:
            return (" | wait watchdog: block (%d,%d,%d) thread %d of grid %d x %d threads, barrier smem 0x%x parity %d, "
                    "complete %x, smem words %s" % (
                        rec[1] >> 32, rec[4] >> 32, rec[4] & 0xFFFFFFFF, rec[1] & 0xFFFFFFFF, rec[2] >> 32,
                        rec[2] & 0xFFFFFFFF, rec[3] >> 32, rec[3] & 0xFFFFFFFF, rec[5],
                        " ".join("%x" % rec[8 + i] for i in range(48))))
    except Exception:
        pass
    return ""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="wmask512", choices=CONFIGS, help="headline configuration (BASELINE.json configs)")
    ap.add_argument("--no-extras", action="store_true", help="measure only the headline configuration")
    ap.add_argument("--extra-steps", type=int, default=5, help="timed steps of each extra configuration")
    ap.add_argument("--stats-steps", type=int, default=200, help="additional steps for the median / spread of the headline")
    ap.add_argument("--rays", type=int, default=0, help="rays per GPU per step of the headline train configuration")
    ap.add_argument("--render-tiles", type=int, default=32, help="4096-ray tiles of the render_image sample")
    ap.add_argument("--grid-resolution", type=int, default=512)
    ap.add_argument("--lvis-points", type=int, default=8192, help="surface points per GPU of the lvis sample")
    ap.add_argument("--bw-rays", type=int, default=65536)
    ap.add_argument("--ref-rays", type=int, default=512,
                    help="rays per step of the CPU reference arm (default: the benchmarked 512-ray step, ~1.2 s each)")
    ap.add_argument("--cpu-rays", type=int, default=512, help="rays of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--debug-flags", type=int, default=0, help="library tuning/bisect flags (development only)")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of the whole-step CUDA graph")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"],
                    help="dense layers: bf16 = tcgen05 tensor cores (FP16 forward / BF16 backward operands, FP32 accumulate), "
                         "fp32 = CUDA-core anchor")
    args = ap.parse_args()
    try:
        # `kill -USR1 <pid>` (or `timeout -s USR1 ...`) dumps every thread's Python stack to stderr: where a stuck run is
        import faulthandler
        import signal
        faulthandler.register(signal.SIGUSR1, all_threads=True)
    except (ImportError, AttributeError, ValueError):
        pass
    wd = int(os.environ.get("FNEUS_BENCH_WATCHDOG", "0"))
    if wd > 0:
        # development aid: dump every thread's Python stack to stderr and exit if the run is still going after `wd` seconds
        import faulthandler
        faulthandler.dump_traceback_later(wd, exit=True)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
