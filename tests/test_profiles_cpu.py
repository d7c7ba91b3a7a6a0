"""The committed profile evidence stays machine-readable: the summary tools parse the committed ncu exports and the bench
lines carry the keys the measurement contract names (no GPU needed)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")


def test_launch_list_is_dominated_by_the_tensor_core_kernels():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from launch_summary import load
    seq = load(os.path.join(PROF, "r1_final_launches.csv"))
    assert len(seq) > 300
    idx = [i for i, s in enumerate(seq) if "upsample_step" in s[0]]
    step = seq[idx[-8]:idx[-4]]
    tot = sum(v for _, v, _ in step)
    tc = sum(v for k, v, _ in step if "sdf_chain_kernel" in k or "relu_chain_pair" in k or "wgrad_group" in k)
    assert 0.6 < tc / tot < 0.95
    assert not any("mlp_chain_kernel" in k for k, _, _ in step)       # the superseded kernel is gone from the path


def test_ncu_summary_tool_reads_the_raw_pages():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"),
                          os.path.join(PROF, "r1_final_ncu_chain_full_raw.csv"),
                          os.path.join(PROF, "r1_final_ncu_relu_pair_full_raw.csv")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    rows = [l for l in out.stdout.splitlines() if l.startswith("| `")]
    assert len(rows) == 7 and any("sdf_chain_kernel<1>" in r for r in rows)


def test_bench_lines_follow_the_contract():
    for name, gpus in (("r1_final_bench_line.json", 1), ("r1_final_bench_line_2gpu.json", 2),
                       ("r1_final_bench_line_4gpu.json", 4), ("r1_final_bench_line_8gpu.json", 8)):
        d = json.loads(open(os.path.join(PROF, name)).read().strip().splitlines()[-1])
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks"):
            assert k in d, (name, k)
        assert d["n_gpus"] == gpus and d["metric"] == "train_rays_per_s" and d["gpu_launches"] > 0
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"] * 1.02
        assert 0.0 < d["roofline"]["frac"] < 1.0 and d["roofline"]["bound"] == "tensor"
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if gpus == 1:
            assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    ref = json.load(open(os.path.join(PROF, "r1_final_reference_arm_line.json")))
    assert ref["impl"] == "reference" and ref["e2e"]["h2d_bytes_per_step"] == 0


def test_clock_sampler_windows_its_samples():
    """bench.ClockSampler.summary: samples inside [mark_begin, mark_end] decide; with none inside (a timed region shorter
    than one nvidia-smi call) the nearest samples are reported and labelled as such."""
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler(0)
    mk = lambda sm, t, pw="Not Active": [str(sm), "1965", "Not Active", "Not Active", "Not Active", pw, t]
    s.rows = [mk(600, 10.0), mk(1965, 20.0, "Active"), mk(1950, 21.0), mk(700, 30.0)]
    s.t0, s.t1 = 19.5, 21.5
    r = s.summary()
    assert r["sm_mhz"] in (1950.0, 1965.0) and r["samples"] == 2 and r["window"] == "timed regions"
    assert r["reasons"] == ["sw_power_cap"]
    s.t0, s.t1 = 24.9, 25.1
    r = s.summary()
    assert r["window"] == "nearest samples" and r["samples"] == 3
    s.rows = []
    assert s.summary()["reasons"] == ["unavailable"]


def test_round2_bench_lines_carry_every_baseline_configuration():
    """Round 2: ONE line per run with the headline (wmask512) and the other BASELINE.json configurations under
    extra_configs, each with its own roofline / e2e; the 1-GPU line also times the CPU port per configuration."""
    base = None
    for gpus in (1, 2, 4, 8):
        d = json.loads(open(os.path.join(PROF, "r2_bench_line_%dgpu.json" % gpus)).read().strip().splitlines()[-1])
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks", "step_time_stats"):
            assert k in d, (gpus, k)
        assert d["n_gpus"] == gpus and d["metric"] == "train_rays_per_s" and d["warmup"] >= 3
        assert d["step_time_stats"]["n"] >= 200
        assert d["roofline"]["traffic"] and d["roofline"]["traffic"] > 1e9           # live designed bytes, not a constant table
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        ex = d["extra_configs"]
        for name in ("womask4096", "render_image", "grid512", "lvis"):
            assert "error" not in ex[name], (gpus, name, ex[name])
            assert ex[name]["value"] > 0 and 0.0 < ex[name]["roofline"]["frac"] < 1.0 and "e2e" in ex[name]
        if gpus == 1:
            base = d
            assert d["cpu_baseline"]["kind"] == "port"
            assert "bandwidth" in ex and len(ex["bandwidth"]["kernels"]) >= 5
            assert all("cpu_baseline" in ex[n] for n in ("womask4096", "render_image", "grid512", "lvis"))
        else:
            assert d["value"] / (gpus * base["value"]) > 0.9                               # >= 7x at 8 GPUs
            assert ex["womask4096"]["value"] / (gpus * base["extra_configs"]["womask4096"]["value"]) > 0.9


def test_round2_ncu_exports_parse():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"),
                          os.path.join(PROF, "r2_ncu_chain_full_raw.csv")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    rows = [l for l in out.stdout.splitlines() if l.startswith("| `")]
    # one step = 5 SDF-forward + 2 colour + 2 RefColor-pair + 1 SDF-backward chain launches and 3 weight-gradient groups
    assert len(rows) == 13 and any("sdf_chain_kernel<1>" in r for r in rows) and any("wgrad_group" in r for r in rows)
    hist = open(os.path.join(PROF, "r2_sass_histogram.md")).read()
    assert "UTCHMMA" in hist and "| **all tensor-core kernels**" in hist
