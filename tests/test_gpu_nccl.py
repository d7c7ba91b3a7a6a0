"""Ray-sharded data parallelism on REAL GPUs (world_size 2, NCCL): the CUDA branch of the sharded step -- CUDA render on each
rank's slice, device-side loss normalisers, fused stage-1 loss, backward, flat-bucket all-reduce -- must reproduce the
single-process full-batch loss and gradient.  (tests/test_dp_gloo.py covers the host logic on CPU with the oracle.)
Skipped with fewer than 2 GPUs."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(dev, precision):
    import factored_neus_b200 as fn
    from factored_neus_b200 import ops
    from factored_neus_b200.parallel import GradBucket
    syn = fn.synthetic
    ops.set_precision(precision)
    st = syn.scene_states(seed=4, jitter=0.03)
    sdf = fn.SDFNetwork(**syn.SDF_CONF); sdf.load_state_dict(st["sdf"])
    col = fn.RenderingNetwork(**syn.COLOR_CONF); col.load_state_dict(st["color"])
    var = fn.SingleVarianceNetwork(0.3); var.load_state_dict(st["var"])
    ref = fn.RefColor(); ref.load_state_dict(st["ref"])
    nets = [sdf.to(dev), var.to(dev), col.to(dev), ref.to(dev)]
    R = fn.NeuSRenderer(**syn.RENDER_CONF_WMASK, sdf_network=nets[0], deviation_network=nets[1], color_network=nets[2],
                        refColor_network=nets[3])
    bucket = GradBucket([p for n in nets for p in n.parameters()])
    return R, bucket


def _step(R, bucket, o, d, rgb, mask, dev, group_reduce):
    from factored_neus_b200 import ops
    from factored_neus_b200.parallel import stage1_loss_sharded
    near, far = ops.near_far_from_sphere(o.to(dev), d.to(dev))
    out = R.render(o.to(dev), d.to(dev), near, far, perturb_overwrite=0, cos_anneal_ratio=1.0)
    loss, _ = stage1_loss_sharded(R, out, rgb.to(dev), mask.to(dev), 0.1, 0.1, 0.1)
    bucket.zero()
    loss.backward()
    if group_reduce:
        bucket.all_reduce()
    return loss.detach()


def _data(B):
    import factored_neus_b200 as fn
    syn = fn.synthetic
    o, d, _, _ = syn.make_rays(B, seed=1)
    rgb, _ = syn.make_targets(B, seed=2)
    mask = (torch.arange(B) % 3 != 0).float()[:, None]            # unequal shard normalisers
    return o, d, rgb, mask


def _worker(rank, world, port, B, precision, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from factored_neus_b200.parallel import shard_rays
    R, bucket = _build(dev, precision)
    o, d, rgb, mask = _data(B)
    lo, hi = shard_rays(B, rank, world)
    loss = _step(R, bucket, o[lo:hi], d[lo:hi], rgb[lo:hi], mask[lo:hi], dev, True)
    dist.all_reduce(loss)
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({"flat": bucket.flat.cpu(), "loss": loss.cpu()}, tmp)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("precision,rtol", [("fp32", 2e-5), ("bf16", 5e-5)])
def test_two_rank_nccl_step_equals_full_batch(tmp_path, precision, rtol):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    B, world = 44, 2
    tmp = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(world, _free_port(), B, precision, tmp), nprocs=world, join=True)
    got = torch.load(tmp)
    from factored_neus_b200 import ops
    dev = torch.device("cuda", 0)
    R, bucket = _build(dev, precision)
    o, d, rgb, mask = _data(B)
    loss = _step(R, bucket, o, d, rgb, mask, dev, False)
    ops.set_precision("fp32")
    flat = bucket.flat.cpu()
    scale = max(1.0, float(flat.abs().max()))
    print("2-rank NCCL [%s]: loss %.6f vs %.6f, max grad diff %.3e (scale %.3e)" % (
        precision, float(got["loss"]), float(loss), float((got["flat"] - flat).abs().max()), scale))
    diffs, off = [], 0
    for p_ in bucket.params:                                   # where the largest differences sit (diagnostic)
        k = p_.numel()
        diffs.append((float((got["flat"][off:off + k] - flat[off:off + k]).abs().max()),
                      float(flat[off:off + k].abs().max()), tuple(p_.shape)))
        off += k
    print("largest per-parameter differences (diff, scale, shape):", sorted(diffs, reverse=True)[:4])
    assert abs(float(got["loss"]) - float(loss)) <= rtol * max(1.0, abs(float(loss)))
    assert float((got["flat"] - flat).abs().max()) <= rtol * scale
