#!/usr/bin/env python
"""bench.py -- train rays/s (fwd + bwd + eikonal) of the wmask stage-1 step (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--rays B]

One "step" = one pass of the hot path over one batch of synthetic rays: NeuSRenderer.render (perturbed,
64+64 samples, 4 up-sampling steps) + stage-1 loss (colour L1, surface L1 x0.1, eikonal x0.1, mask BCE x0.1;
exp_runner.py:134-177) + backward (incl. the SDF double backward) + Adam step.  Networks are at random geometric
init ("sphere-SDF scene"); data is synthetic.  N > 1 runs one process per GPU under torchrun, rays
sharded (weak scaling: 512 rays per GPU), exact global loss normalisers, one flat-bucket NCCL all-reduce.

``--impl reference`` times the reference algorithm's CPU restatement (oracle/neus_oracle.py; the reference is
pure PyTorch and /root/reference does not exist on the GPU box) on the host cores, same config and metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_RAY_TRAIN_WMASK = 1_138_099_200          # 880 F_sdf + 384 F_col + 6 F_ref
SURFACE_W, IGR_W, MASK_W = 0.1, 0.1, 0.1


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


def run_reference(args):
    """CPU arm: the oracle port of the reference's render + loss + backward, all host threads."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import neus_oracle as O
    import factored_neus_b200 as fn
    syn = fn.synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = args.ref_rays
    st = syn.scene_states(seed=4, jitter=0.0)
    P = {k: {n: t.clone().requires_grad_(True) for n, t in sd.items()} for k, sd in st.items()}
    o, d, near, far = syn.make_rays(B, seed=1)
    true_rgb, mask = syn.make_targets(B, seed=2)

    def step():
        for sd in P.values():
            for t in sd.values():
                t.grad = None
        out = O.render(P, o, d, near, far, conf=O.RENDER_CONF_WMASK, cos_anneal_ratio=1.0)
        loss, _ = O.stage1_loss(out, true_rgb, mask, SURFACE_W, IGR_W, MASK_W)
        loss.backward()
        return float(loss)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = B * args.steps / dt
    sample = "%d synthetic rays/step x %d steps (oracle port of renderer.py render + stage-1 loss + backward)" % (
        B, args.steps)
    print(json.dumps({
        "impl": "reference", "metric": "train_rays_per_s", "value": v, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "wmask stage-1 train step (fwd+bwd+eikonal), 64+64 samples, 4 up-sample steps",
                   "rays_per_step": B},
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline(rays, steps=2):
    import torch
    from oracle import neus_oracle as O
    import factored_neus_b200 as fn
    syn = fn.synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    st = syn.scene_states(seed=4, jitter=0.0)
    P = {k: {n: t.clone().requires_grad_(True) for n, t in sd.items()} for k, sd in st.items()}
    o, d, near, far = syn.make_rays(rays, seed=1)
    true_rgb, mask = syn.make_targets(rays, seed=2)
    best = None
    for i in range(steps + 1):
        t0 = time.perf_counter()
        out = O.render(P, o, d, near, far, conf=O.RENDER_CONF_WMASK, cos_anneal_ratio=1.0)
        loss, _ = O.stage1_loss(out, true_rgb, mask, SURFACE_W, IGR_W, MASK_W)
        loss.backward()
        dt = time.perf_counter() - t0
        if i > 0:
            best = dt if best is None else min(best, dt)
    return {"value": rays / best, "unit": "rays/s", "cores": cores, "kind": "port",
            "sample": "%d rays, best of %d steps after 1 warm-up, oracle port (CPU torch)" % (rays, steps)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import factored_neus_b200 as fn
    from factored_neus_b200 import _lib as L
    from factored_neus_b200 import ops as _ops
    from factored_neus_b200.parallel import FlatAdam, GradBucket, stage1_loss_sharded
    syn = fn.synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # libraries (NCCL's version banner) write to file descriptor 1: keep stdout for the ONE JSON line
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.rays
    st = syn.scene_states(seed=4, jitter=0.0)
    sdf = fn.SDFNetwork(**syn.SDF_CONF); sdf.load_state_dict(st["sdf"])
    col = fn.RenderingNetwork(**syn.COLOR_CONF); col.load_state_dict(st["color"])
    var = fn.SingleVarianceNetwork(0.3); var.load_state_dict(st["var"])
    ref = fn.RefColor(); ref.load_state_dict(st["ref"])
    nets = [sdf.to(dev), var.to(dev), col.to(dev), ref.to(dev)]
    R = fn.NeuSRenderer(**syn.RENDER_CONF_WMASK, sdf_network=nets[0], deviation_network=nets[1],
                        color_network=nets[2], refColor_network=nets[3])
    params = [p for n in nets for p in n.parameters()]
    bucket = GradBucket(params)
    opt = FlatAdam(bucket, lr=5e-4, warm_up_end=5000, end_iter=300000)   # fused flat Adam + on-device LR schedule
    opt.set_iteration(5000)                                              # past the warm-up: full learning rate

    o, d, near, far = syn.make_rays(B, seed=1 + rank)
    true_rgb, mask = syn.make_targets(B, seed=100 + rank)
    host = torch.cat([o, d, true_rgb, mask], dim=1).pin_memory()          # [B,10] like dataset.gen_random_rays_at
    dev_batch = host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2

    def step(batch):
        ro, rd, rgb, m = batch[:, :3], batch[:, 3:6], batch[:, 6:9], batch[:, 9:10]
        near, far = _ops.near_far_from_sphere(ro, rd)                       # dataset.near_far_from_sphere
        out = R.render(ro, rd, near, far, cos_anneal_ratio=1.0)
        loss, _ = stage1_loss_sharded(R, out, rgb, m, SURFACE_W, IGR_W, MASK_W)
        loss.backward()
        bucket.all_reduce()
        opt.step()                                                          # also clears the gradient bucket
        return loss

    sampler = ClockSampler(local)
    sampler.start()
    lib = L.lib()
    _ops.set_precision(args.precision)
    if args.debug_flags:
        lib.fneus_debug_flags(args.debug_flags)
    # every step (warm-up, capture, replay, eager) runs on one non-default stream so that autograd's
    # AccumulateGrad nodes and the flat gradient bucket live on the capture stream
    work_stream = torch.cuda.Stream()
    work_stream.wait_stream(torch.cuda.current_stream())
    torch.cuda.set_stream(work_stream)
    for _ in range(max(3, args.warmup)):
        step(dev_batch)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- whole-step CUDA graph (fixed shapes; removes ~2k launches of host overhead per step) ----
    graph, static_batch, static_loss = None, dev_batch.clone(), None
    if not args.no_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=work_stream):
                static_loss = step(static_batch)
            graph.replay()
            torch.cuda.synchronize()
            if not torch.isfinite(static_loss).all():
                raise RuntimeError("non-finite loss from the graphed step")
        except Exception as ex:
            # a failed capture leaves the CUDA RNG in capture mode: re-run this process eagerly instead
            if rank == 0:
                print("bench: CUDA graph capture failed (%s); re-running with --no-graph" % str(ex).splitlines()[0],
                      file=sys.stderr)
            if world == 1:
                sys.stdout.flush()
                os.dup2(real_stdout, 1)                                     # the re-executed process prints the JSON line
                os.execv(sys.executable, [sys.executable] + sys.argv + ["--no-graph"])
            raise

    def run_step(batch):
        if graph is None:
            return step(batch)
        if batch is not static_batch:
            static_batch.copy_(batch, non_blocking=True)
        graph.replay()
        return static_loss

    # ---------------- device-resident timed region (value) -------------------------------------
    barrier()
    sampler.mark_begin()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.zero_()                                                       # L2 flush between timed iterations
        ev[i][0].record()
        run_step(static_batch)
        ev[i][1].record()
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)

    # ---------------- per-kernel CUDA-event pass (roofline): same step, eager launches bracketed by events ----
    ncls = lib.fneus_prof_classes()
    import ctypes
    ms_c = (ctypes.c_double * ncls)(); ln_c = (ctypes.c_longlong * ncls)()
    fl_c = (ctypes.c_double * ncls)(); by_c = (ctypes.c_double * ncls)()
    lib.fneus_prof_collect(None, None, None, None)
    lib.fneus_prof_enable(1)
    prof_steps = min(args.steps, 5)
    pev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    barrier()
    pev[0].record()
    for i in range(prof_steps):
        step(dev_batch)
    pev[1].record()
    barrier()
    prof_ms = pev[0].elapsed_time(pev[1])
    L.check(lib.fneus_prof_collect(ms_c, ln_c, fl_c, by_c), "prof_collect")
    lib.fneus_prof_enable(0)
    launches = int(sum(ln_c)) // max(1, prof_steps) * args.steps

    # ---------------- end-to-end timed region (host buffers) -----------------------------------
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()
    barrier()
    t_e2e = 0.0
    for i in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        batch = host.to(dev, non_blocking=True)                             # H2D of this step's rays [B,10]
        loss = run_step(batch)
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)        # D2H of the step's result
        torch.cuda.synchronize()
        t_e2e += time.perf_counter() - t0
    barrier()
    sampler.mark_end()
    sampler.stop_flag = True

    t = torch.tensor([dev_ms, t_e2e * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    if rank == 0:
        pk = _peaks()
        total_rays = B * world * args.steps
        value = total_rays / (dev_ms * 1e-3)
        e2e_v = total_rays / (e2e_ms * 1e-3)
        names = ["gemm_fwd", "gemm_bwd_data", "gemm_wgrad", "sampling", "composite", "elementwise", "tc_mlp"]
        per_class = {names[c]: {"ms_per_step": ms_c[c] / prof_steps, "launches_per_step": ln_c[c] / prof_steps,
                                "gflop_per_step": fl_c[c] / prof_steps / 1e9} for c in range(ncls)}
        # dominant kernel: the dense-layer GEMMs (one kernel template, three operand layouts)
        gemm_ms = ms_c[0] + ms_c[1] + ms_c[2] + ms_c[6]
        gemm_fl = fl_c[0] + fl_c[1] + fl_c[2] + fl_c[6]
        achieved = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        line = {
            "metric": "train_rays_per_s", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
            "data": "synthetic",
            "config": {"workload": "wmask stage-1 train step (render fwd + loss + bwd incl. SDF double backward + "
                                   "Adam), 64+64 samples, 4 up-sample steps, sphere-SDF scene at geometric init",
                       "rays_per_gpu_per_step": B, "global_rays_per_step": B * world, "parallelism": "dp%d" % world,
                       "l2": "256 MiB flush between timed iterations; per-step working set ~1.5 GB >> 126 MB L2",
                       "cuda_graph": graph is not None,
                       "precision_path": "bf16 operands on tcgen05, fp32 accumulate/activations" if args.precision == "bf16"
                       else "fp32-simt"},
            "e2e": {"value": e2e_v, "unit": "rays/s", "h2d_bytes_per_step": host.numel() * 4 * world,
                    "d2h_bytes_per_step": 4 * world, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                         "frac": achieved / pk["tf_sust"],
                         # DRAM bytes per launch (read + write) of the largest single launch, sdf_chain_kernel<1> on
                         # 65 536 points, from the ncu --set full capture summarised in profiles/ (not measured live)
                         "traffic": 1.816e9 if args.precision == "bf16" and B == 512 else None,
                         "traffic_detail": ({"sdf_chain_kernel<0> fwd": 0.698e9, "sdf_chain_kernel<1> bwd": 1.816e9,
                                             "tc_gemm_wgrad_group_kernel (12 jobs)": 0.793e9,
                                             "source": "profiles/r1_final_ncu_chain_kernels.md"}
                                            if args.precision == "bf16" and B == 512 else None),
                         "peak_source": pk["src"],
                         "kernel": ("sdf_chain_kernel<SDF fwd|SDF bwd|ReLU> + tc_gemm_wgrad_group_kernel (tcgen05 dense MLP layers)" if args.precision == "bf16"
                                    else "gemm_mk_kernel/gemm_wgrad_kernel (dense MLP layers)"),
                         "kernel_share_of_step": (gemm_ms / prof_steps) / (dev_ms / args.steps),
                         "measured": "CUDA events around every launch of %d eager steps (%.2f ms/step with events)"
                                     % (prof_steps, prof_ms / prof_steps),
                         "step_algorithmic_tflops": FLOP_PER_RAY_TRAIN_WMASK * value / 1e12},
            "kernel_classes": per_class,
            "clocks": sampler.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.cpu_rays)
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        # NCCL communicators captured inside a CUDA graph can stall destroy_process_group(): the result is out,
        # so flush and leave without the collective teardown
        sys.stdout.flush()
        sys.stderr.flush()
        torch.cuda.synchronize()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=512, help="rays per GPU per step")
    ap.add_argument("--ref-rays", type=int, default=512,
                    help="rays per step of the CPU reference arm (default: the benchmarked 512-ray step, ~1.2 s each)")
    ap.add_argument("--cpu-rays", type=int, default=512, help="rays of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--debug-flags", type=int, default=0, help="library tuning/bisect flags (development only)")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of the whole-step CUDA graph")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"],
                    help="dense layers: bf16 = tcgen05 tensor cores (FP32 accumulate), fp32 = CUDA-core anchor")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
