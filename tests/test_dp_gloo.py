"""Ray-sharded data parallelism (factored-neus_b200/parallel.py) on CPU: world_size 2, gloo.
The per-shard render is done by the CPU oracle (tests may use it); what is under test is the host logic:
shard boundaries, global loss normalisers, flat gradient bucket + all-reduce == single-process full batch."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


SMALL_SDF = dict(d_in=3, d_out=33, d_hidden=32, n_layers=4, skip_in=(2,), multires=2, bias=0.5, scale=1.0)
SMALL_COL = dict(d_feature=32, d_in=9, d_out=3, d_hidden=32, n_layers=2, multires_view=2)
SMALL_RENDER = dict(n_samples=16, n_importance=8, n_outside=0, up_sample_steps=2, perturb=0.0)


def _small_states():
    import factored_neus_b200 as fn
    syn = fn.synthetic
    sdf_conf = dict(syn.SDF_CONF, d_out=33, d_hidden=32, n_layers=4, skip_in=(2,), multires=2)
    col_conf = dict(syn.COLOR_CONF, d_feature=32, d_hidden=32, n_layers=2, multires_view=2)
    st = {"sdf": syn.sdf_state(4, sdf_conf, 0.03), "color": syn.color_state(5, col_conf, 0.03),
          "var": syn.variance_state(0.3), "ref": syn.refcolor_state(7, 32, 32)}
    return st


class _FakeRenderer:
    pass


def _shard_loss(P, o, d, near, far, rgb, mask):
    from oracle import neus_oracle as O
    from factored_neus_b200.parallel import stage1_loss_sharded
    out = O.render(P, o, d, near, far, conf=SMALL_RENDER, perturb_overwrite=0, cos_anneal_ratio=0.5,
                   sdf_conf=SMALL_SDF, color_conf=SMALL_COL)
    R = _FakeRenderer()
    R.last_eikonal_parts = (out["eik_num"], out["eik_den"])
    return stage1_loss_sharded(R, out, rgb, mask, 0.1, 0.1, 0.1)[0]


def _worker(rank, world, port, B, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import factored_neus_b200 as fn
    from factored_neus_b200.parallel import GradBucket, shard_rays
    syn = fn.synthetic
    torch.set_num_threads(2)
    st = _small_states()
    P = {k: {n: t.clone().requires_grad_(True) for n, t in sd.items()} for k, sd in st.items()}
    params = [t for sd in P.values() for t in sd.values()]
    bucket = GradBucket(params)
    o, d, near, far = syn.make_rays(B, seed=1)
    rgb, mask = syn.make_targets(B, seed=2)
    mask = (torch.arange(B) % 3 != 0).float()[:, None]            # non-trivial mask -> unequal shard normalisers
    lo, hi = shard_rays(B, rank, world)
    loss = _shard_loss(P, o[lo:hi], d[lo:hi], near[lo:hi], far[lo:hi], rgb[lo:hi], mask[lo:hi])
    bucket.zero()
    loss.backward()
    bucket.all_reduce()
    tot = loss.detach().clone()
    dist.all_reduce(tot)
    if rank == 0:
        torch.save({"flat": bucket.flat.clone(), "loss": tot}, tmp)
    dist.destroy_process_group()


def test_sharded_step_equals_full_batch(tmp_path):
    B, world = 22, 2
    tmp = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(world, _free_port(), B, tmp), nprocs=world, join=True)
    got = torch.load(tmp)
    import factored_neus_b200 as fn
    from oracle import neus_oracle as O
    syn = fn.synthetic
    st = _small_states()
    P = {k: {n: t.clone().requires_grad_(True) for n, t in sd.items()} for k, sd in st.items()}
    o, d, near, far = syn.make_rays(B, seed=1)
    rgb, _ = syn.make_targets(B, seed=2)
    mask = (torch.arange(B) % 3 != 0).float()[:, None]
    out = O.render(P, o, d, near, far, conf=SMALL_RENDER, perturb_overwrite=0, cos_anneal_ratio=0.5,
                   sdf_conf=SMALL_SDF, color_conf=SMALL_COL)
    loss, _ = O.stage1_loss(out, rgb, mask, 0.1, 0.1, 0.1)
    loss.backward()
    flat = torch.cat([(t.grad if t.grad is not None else torch.zeros_like(t)).reshape(-1)
                      for sd in P.values() for t in sd.values()])
    assert abs(float(got["loss"]) - float(loss)) < 1e-5
    assert float((got["flat"] - flat).abs().max()) < 2e-5 * max(1.0, float(flat.abs().max()))


def test_shard_rays_cover_batch():
    from factored_neus_b200.parallel import shard_rays
    for n, w in ((512, 8), (10, 4), (3, 8), (0, 2)):
        spans = [shard_rays(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


# ------------------------------------------------------------------------------------------ tile-sharded inference (host logic)
def _tile_worker(rank, world, port, N, tile, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from factored_neus_b200.parallel import gather_tiles, shard_tiles
    n_tiles = (N + tile - 1) // tile
    mine = shard_tiles(n_tiles, rank, world)
    rows = []
    for t in mine:                                   # "render" of tile t: item index in column 0, 2 * index in column 1
        idx = torch.arange(t * tile, (t + 1) * tile, dtype=torch.float32)
        v = torch.stack([idx, 2 * idx], dim=1)
        v[idx >= N] = 0.0
        rows.append(v)
    local = torch.cat(rows) if rows else torch.zeros(0, 2)
    full = gather_tiles(local, N, tile, rank, world)
    if rank == 0:
        torch.save(full, tmp)
    else:
        assert full is None
    dist.destroy_process_group()


@pytest.mark.parametrize("N,tile,world", [(130, 64, 3), (5, 8, 2)])
def test_gather_tiles_restores_item_order(tmp_path, N, tile, world):
    """Round-robin tile ownership + one gather == the unsharded result, incl. a partial last tile, ranks with one tile
    fewer than others, and ranks with no tile at all."""
    tmp = str(tmp_path / "tiles.pt")
    mp.spawn(_tile_worker, args=(world, _free_port(), N, tile, tmp), nprocs=world, join=True)
    full = torch.load(tmp)
    idx = torch.arange(N, dtype=torch.float32)
    assert full.shape == (N, 2)
    assert torch.equal(full[:, 0], idx) and torch.equal(full[:, 1], 2 * idx)


def test_shard_tiles_partition():
    from factored_neus_b200.parallel import shard_tiles
    for n, w in ((469, 8), (3, 8), (0, 2), (16, 4)):
        owned = sorted(t for r in range(w) for t in shard_tiles(n, r, w))
        assert owned == list(range(n))
        sizes = [len(shard_tiles(n, r, w)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1
