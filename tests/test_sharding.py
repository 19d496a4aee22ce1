"""The N > 1 path on CPU: sharding helpers, and a world_size-2 gloo run in which each rank renders its band /
item range with the CPU oracle standing in for the device, gathers, and checks the stitched result against the
unsharded one (both modes are exchange-free)."""
import os
import socket

import numpy as np
import pytest

from rasterize_b200 import sharding


def test_shard_range_covers_and_balances():
    for n in (0, 1, 7, 100, 100003):
        for world in (1, 2, 3, 8):
            ranges = [sharding.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            assert max(e - b for b, e in ranges) - min(e - b for b, e in ranges) <= 1
    w = np.array([1] * 50 + [10] * 50, dtype=float)
    ranges = [sharding.shard_range(100, r, 2, w) for r in range(2)]
    assert ranges[0][1] == ranges[1][0] and ranges[0][0] == 0 and ranges[1][1] == 100
    s0, s1 = w[ranges[0][0]:ranges[0][1]].sum(), w[ranges[1][0]:ranges[1][1]].sum()
    assert abs(s0 - s1) <= 10


def test_band_rows():
    for h in (1, 7, 64, 453, 4096, 32768):
        for world in (1, 2, 4, 8):
            bands = [sharding.band_rows(h, r, world) for r in range(world)]
            assert bands[0][0] == 0 and bands[-1][1] == h
            assert all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
            assert all(b[0] % 8 == 0 for b in bands if b[0] < h)
    assert sharding.band_rows(32768, 3, 8) == (12288, 16384)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist

    import oracle as O
    import assets
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # --- band mode: one canvas, rows split across ranks
        p = assets.load_path("squirrel")
        e = assets.expected()["paths"]["squirrel"]["c1"]
        w, h = e["size"]
        tr = np.array(e["tr"])
        op = O.OraclePath.from_flat(p.points, p.kinds, p.subpath_offsets, p.closed)
        y0, y1 = sharding.band_rows(h, rank, world)
        band = np.zeros((y1 - y0, w))
        op.mask(sharding.band_transform(tr, y0), O.NONZERO, band)
        stitched = torch.zeros((h, w), dtype=torch.float64)
        stitched[y0:y1] = torch.from_numpy(band)
        dist.all_reduce(stitched)  # test-side gather only (bands are disjoint); the product gathers with cudaMemcpyAsync
        full = np.zeros((h, w))
        op.mask(tr, O.NONZERO, full)
        err_band = float(np.abs(stitched.numpy() - full).max())
        # --- batch mode: independent glyphs split by index
        n = 12
        b, eidx = sharding.shard_range(n, rank, world, weights=[18] * n)
        sums = torch.zeros(n, dtype=torch.float64)
        for i in range(b, eidx):
            img = np.zeros((64, 64))
            O.OraclePath.glyph(i + 1).mask(O.IDENTITY, O.NONZERO, img)
            sums[i] = img.sum()
        dist.all_reduce(sums)
        ref = []
        for i in range(n):
            img = np.zeros((64, 64))
            O.OraclePath.glyph(i + 1).mask(O.IDENTITY, O.NONZERO, img)
            ref.append(img.sum())
        err_batch = float(np.abs(sums.numpy() - np.array(ref)).max())
        # max-over-ranks reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            out.put((err_band, err_batch, float(t.item())))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_band_and_batch_sharding():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    err_band, err_batch, tmax = q.get(timeout=5)
    assert err_band < 1e-12   # bands stitched == unsharded mask
    assert err_batch == 0.0
    assert tmax == 2.0


def test_band_costs_balance_blocks():
    """Cost-balanced blocks of scanline bands (bench.py config 5 at N > 1): the blocks tile the canvas, and a block in the dense
    middle of the outline gets fewer rows than an empty one."""
    import numpy as np
    from rasterize_b200 import sharding
    h, nb = 32768, 256
    rng = np.random.default_rng(0)
    y0 = rng.uniform(9000, 23000, 5000)
    lines = np.stack([rng.uniform(0, 32768, 5000), y0, rng.uniform(0, 32768, 5000), y0 + rng.uniform(-900, 900, 5000)], axis=1)
    costs = sharding.band_costs(lines, h, nb, 32768)
    assert costs.shape == (nb,) and np.all(costs >= 128 - 1e-9) and costs[0] == 128 and costs[nb // 2] > 2 * 128
    for world in (1, 2, 3, 8):
        cuts = [sharding.shard_range(nb, k, world, costs)[0] for k in range(world)] + [nb]
        assert cuts[0] == 0 and all(a <= b for a, b in zip(cuts, cuts[1:]))
        assert [sharding.shard_range(nb, k, world, costs) for k in range(world)] == list(zip(cuts[:-1], cuts[1:]))
        block = [costs[a:b].sum() for a, b in zip(cuts[:-1], cuts[1:])]
        assert max(block) <= costs.sum() / world + costs.max()  # within one band of the ideal
    cuts = [sharding.shard_range(nb, k, 8, costs)[0] for k in range(8)] + [nb]
    sizes = np.diff(cuts)
    assert sizes[0] > sizes[3] and sizes[7] > sizes[4]
