"""north_star gate for the tensor-core path, AS WRITTEN: PSNR after 1k training iterations within 0.1 dB of the FP32
path, end to end (trained AND rendered on the tensor-core path); two runs per arm, ensemble means, no allowance for noise.

Teacher = the same architecture at another random init (seed 123) rendered unperturbed on the FP32 path; both
students start from seed 4, see the identical ray order and identical perturbation random numbers, and are
evaluated on held-out rays.  (SURVEY.md 8d "correctness gates".)"""
import math
import os

import pytest
import torch

import factored_neus_b200 as fn
from factored_neus_b200 import ops
from factored_neus_b200.train import Stage1Trainer
from util import build_modules, syn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ITERS = int(os.environ.get("FNEUS_PSNR_ITERS", "1000"))
B = 512
# A single checkpoint's held-out PSNR moves by a few 0.1 dB from one checkpoint to the next (minibatch noise), so the reading
# compared is the mean over the checkpoints of the last 200 iterations.
EVAL_SPAN, EVAL_EVERY = 200, 10
GATE_DB = 0.1
REPEATS = int(os.environ.get("FNEUS_PSNR_REPEATS", "2"))
WARM_UP_END = int(os.environ.get("FNEUS_PSNR_WARM_UP_END", "100"))
END_ITER = int(os.environ.get("FNEUS_PSNR_END_ITER", str(ITERS)))
# Peak learning rate of both arms: 1e-4 = what the reference's schedule (5e-4 with a 5000-iteration warm-up,
# exp_runner.py:229-238) reaches at iteration 1000.  At 5e-4 from iteration 100 on (an earlier version of this test) the
# optimisation is chaotic enough that two runs of the SAME arm end 0.1-0.25 dB apart (the weight gradients are summed with
# FP32 atomics) -- the bound could not be resolved; at 1e-4 both arms reach the same ~50.6 dB and repeat to ~0.04 dB.
LR = float(os.environ.get("FNEUS_PSNR_LR", "1e-4"))


def _render_rgb(R, o, d, near, far):
    outs = []
    with torch.no_grad():
        for i in range(0, o.shape[0], 2048):
            out = R.render(o[i:i + 2048], d[i:i + 2048], near[i:i + 2048], far[i:i + 2048], perturb_overwrite=0,
                           cos_anneal_ratio=1.0)
            outs.append(out["color_fine"].detach())
    return torch.cat(outs)


def _psnr(a, b):
    return 20.0 * math.log10(1.0 / math.sqrt(float(((a - b) ** 2).mean())))


def test_bf16_training_tracks_fp32_psnr():
    ops.set_precision("fp32")
    teacher = build_modules(syn.scene_states(seed=123, jitter=0.03), DEV, syn.RENDER_CONF_WMASK)
    n_train, n_test = 16 * B, 2048
    rays = [t.to(DEV) for t in syn.make_rays(n_train + n_test, seed=77)]
    true_rgb = _render_rgb(teacher["renderer"], *rays)
    batches = torch.cat([rays[0], rays[1], true_rgb, torch.ones_like(rays[2])], dim=1)       # [N,10]
    train, test = batches[:n_train], batches[n_train:]
    order = torch.randperm(n_train // B, generator=torch.Generator().manual_seed(5))

    def run(prec):
        ops.set_precision(prec)
        m = build_modules(syn.scene_states(seed=4), DEV, syn.RENDER_CONF_WMASK)
        # both arms run the captured step: identical launch sequence and identical Philox offsets for the perturbation
        tr = Stage1Trainer(m["renderer"], [m["sdf"], m["var"], m["color"], m["ref"]], B, lr=LR, warm_up_end=WARM_UP_END,
                           end_iter=END_ITER, use_graph=True)
        torch.manual_seed(11)
        evals, evals32 = [], []
        for it in range(ITERS):
            k = int(order[it % len(order)])
            tr.step(train[k * B:(k + 1) * B])
            if it + 1 > ITERS - EVAL_SPAN and (ITERS - 1 - it) % EVAL_EVERY == 0:
                torch.cuda.synchronize()
                args = (test[:, :3], test[:, 3:6], rays[2][n_train:], rays[3][n_train:])
                evals.append(_psnr(_render_rgb(m["renderer"], *args), test[:, 6:9]))
                if prec == "bf16":         # the same weights rendered by the FP32 path: what was LEARNED
                    ops.set_precision("fp32")
                    evals32.append(_psnr(_render_rgb(m["renderer"], *args), test[:, 6:9]))
                    ops.set_precision("bf16")
        return evals, evals32

    # Two runs per arm: the weight gradients are summed with FP32 atomics (split-K), so two runs of the same arm are not
    # bit-identical; the spread between the repeats is printed next to the difference it has to be read against.
    mean = lambda v: sum(v) / len(v)
    fmt = lambda v: " ".join("%.2f" % x for x in v)
    runs32, runs16, runs16_as32 = [], [], []
    for rep in range(REPEATS):
        e32, _ = run("fp32")
        e16, e16_as32 = run("bf16")
        runs32.append(mean(e32)); runs16.append(mean(e16)); runs16_as32.append(mean(e16_as32))
        print("run %d, held-out PSNR at the last %d checkpoints:\n  fp32-trained, fp32 render: %s\n  tc-trained,   tc render:   %s\n"
              "  tc-trained,   fp32 render: %s" % (rep, len(e32), fmt(e32), fmt(e16), fmt(e16_as32)))
    ops.set_precision("fp32")
    p32, p16, p16_as32 = mean(runs32), mean(runs16), mean(runs16_as32)
    spread32 = max(runs32) - min(runs32)
    spread16 = max(runs16) - min(runs16)
    # standard error of a mean of n runs ~ range / (2 sqrt(n)) for n = 2; of the difference: root sum of squares
    se = 0.5 * math.sqrt(spread32 ** 2 + spread16 ** 2) / math.sqrt(REPEATS / 2.0)
    print("PSNR after %d iterations (%d runs per arm): fp32 %.3f dB (runs %s) | tensor-core end to end %.3f dB (runs %s, diff "
          "%+.3f) | tensor-core-trained rendered in fp32 %.3f dB (diff %+.3f) | render-only rounding cost %+.3f dB | "
          "run-to-run spread fp32 %.3f, tensor-core %.3f -> standard error of the difference %.3f dB"
          % (ITERS, REPEATS, p32, fmt(runs32), p16, fmt(runs16), p16 - p32, p16_as32, p16_as32 - p32, p16 - p16_as32,
             spread32, spread16, se))
    # The north_star bound AS WRITTEN: end to end, |difference| <= 0.1 dB (ensemble means, no allowance for the spread).
    assert abs(p16 - p32) <= GATE_DB, (
        "tensor-core end-to-end PSNR %.3f vs fp32 %.3f: outside %.1f dB (standard error %.3f)" % (p16, p32, GATE_DB, se))
    assert abs(p16_as32 - p32) <= GATE_DB, (
        "tensor-core-trained weights (fp32 render) %.3f vs fp32 %.3f" % (p16_as32, p32))
    # the deterministic component: what the tensor-core renderer's operand rounding costs on FIXED weights (no trajectory
    # noise in this one: the same weights rendered by both paths)
    assert abs(p16 - p16_as32) <= 0.03, "operand rounding of the tensor-core renderer costs %.3f dB on fixed weights" % (
        p16 - p16_as32)
