"""Data-parallel host logic on CPU: world_size 2 over gloo (sharding + the single logits all-gather)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vit_pytorch_b200 import ViT
from vit_pytorch_b200.parallel import all_gather_logits, data_parallel_forward, shard_bounds, shard_sizes


def test_shard_bounds_cover_the_batch():
    for total in (0, 1, 7, 8, 512, 1023):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(shard_sizes(total, world)) - min(shard_sizes(total, world)) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    torch.set_num_threads(1)
    model = ViT(image_size=16, patch_size=8, num_classes=5, dim=32, depth=1, heads=2, mlp_dim=32, dim_head=16).eval()
    g = torch.Generator().manual_seed(7)
    imgs = torch.randn(total, 3, 16, 16, generator=g)
    lo, hi = shard_bounds(total, rank, world)
    out = data_parallel_forward(model, imgs[lo:hi], total=total)
    full = model(imgs)
    ok = out.shape == full.shape and torch.allclose(out, full, atol=1e-6)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [6, 5])
def test_two_rank_gloo_all_gather(total):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(ret) == {0: True, 1: True}


def test_all_gather_without_process_group_is_identity():
    x = torch.randn(3, 4)
    assert all_gather_logits(x) is x


def test_balanced_bounds_minimise_the_largest_slice():
    from vit_pytorch_b200.parallel import balanced_bounds
    import itertools
    import random
    rnd = random.Random(0)
    for n, world in ((1, 4), (5, 2), (9, 3), (12, 4), (7, 8)):
        costs = [rnd.randint(1, 100) ** 2 for _ in range(n)]
        b = balanced_bounds(costs, world)
        assert len(b) == world and b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        got = max(sum(costs[lo:hi]) for lo, hi in b)
        # brute force over all contiguous partitions into at most `world` slices
        best = min(max(sum(costs[a:c]) for a, c in zip((0,) + cuts, cuts + (n,)))
                   for k in range(min(world, n)) for cuts in itertools.combinations(range(1, n), k))
        assert got <= best * (1 + 1e-6), (costs, world, got, best)
    assert balanced_bounds([], 3) == [(0, 0)] * 3


def _navit_worker(rank, world, port, ret):
    from vit_pytorch_b200.na_vit import NaViT
    from vit_pytorch_b200.parallel import navit_data_parallel_forward
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    torch.set_num_threads(1)
    model = NaViT(image_size=32, patch_size=8, num_classes=5, dim=32, depth=1, heads=2, mlp_dim=32, dim_head=16).eval()
    g = torch.Generator().manual_seed(3)
    sizes = [(32, 32), (8, 8), (8, 16), (16, 8), (8, 8), (24, 32), (8, 8)]     # one large image first: 1 | 6 split
    imgs = [torch.randn(3, h, w, generator=g) for h, w in sizes]
    out = navit_data_parallel_forward(model, imgs)
    full = model(imgs)
    ret[rank] = bool(out.shape == full.shape and torch.allclose(out, full, atol=1e-5))
    dist.destroy_process_group()


def test_two_rank_gloo_navit_balanced_split():
    """Variable-resolution images are split by per-layer work, not by count; logits come back in input order."""
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_navit_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret[r] for r in range(world))
