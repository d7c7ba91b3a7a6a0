"""Time SDF forward (value+feature, no graph) on N points for each debug-flag setting."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import factored_neus_b200 as fn
from factored_neus_b200 import ops, _lib as L
syn = fn.synthetic
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
st = syn.scene_states(seed=4)
sdf = fn.SDFNetwork(**syn.SDF_CONF); sdf.load_state_dict(st["sdf"]); sdf = sdf.cuda()
x = (torch.rand(N, 3, device="cuda") * 2 - 1)
w = sdf.flat_weights().detach()
def run(prec, flags, reps=10):
    ops.set_precision(prec)
    L.lib().fneus_debug_flags(flags)
    for _ in range(3):
        ops.sdf_forward_nograd(sdf.cfg, w, x, want_feat=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        ops.sdf_forward_nograd(sdf.cfg, w, x, want_feat=True)
    b.record(); torch.cuda.synchronize()
    L.lib().fneus_debug_flags(0)
    return a.elapsed_time(b) / reps
for prec, flags, label in (("fp32", 0, "fp32 simt"), ("bf16", 0, "bf16 full"), ("bf16", 1, "no A loads"), ("bf16", 2, "no epilogue"),
                           ("bf16", 4, "no MMA"), ("bf16", 1 + 8, "no A, epi transposes only"), ("bf16", 1 + 16, "no A, epi no stores"), ("bf16", 3, "no A, no epilogue"), ("bf16", 7, "nothing")):
    print("%-20s %8.3f ms per SDF forward of %d points (9 layers)" % (label, run(prec, flags), N))

def run_sdf_only(prec, reps=10):
    ops.set_precision(prec)
    for _ in range(3):
        ops.sdf_forward_nograd(sdf.cfg, w, x, want_feat=False)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        ops.sdf_forward_nograd(sdf.cfg, w, x, want_feat=False)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for prec in ("fp32", "bf16"):
    t = run_sdf_only(prec)
    print("%-20s %8.3f ms per sdf-only forward of %d points -> %.1f M pts/s, %.1f TFLOP/s" % (
        prec + " sdf only", t, N, N / t / 1e3, N * 983552 / t / 1e9))
