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
