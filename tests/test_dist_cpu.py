"""CPU (gloo, world_size 2): the multi-GPU plumbing of the view-parallel path — view sharding, the flat
gradient bucket and its single sum-allreduce (gscream_b200/dist.py).  The rasterizer itself is not called."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gscream_b200.dist import GradBucket, allreduce_bucket, shard_views


def test_shard_views_partitions_every_view_exactly_once():
    for n, w in ((8, 1), (8, 2), (8, 8), (5, 2), (3, 4)):
        seen = sorted(v for r in range(w) for v in shard_views(n, w, r))
        assert seen == list(range(n))
    assert shard_views(8, 4, 1) == [1, 5]
    with pytest.raises(ValueError):
        shard_views(4, 2, 2)


def test_bucket_layout_is_one_flat_buffer_of_contiguous_blocks():
    P, C = 11, 32
    b = GradBucket(P, C)
    assert b.numel == P * (3 + 3 + C + 1 + 1 + 3 + 4) and b.nbytes() == b.numel * 4
    # the all-reduced range is a prefix that starts with the colour block and leaves out the per-view means2D block
    assert b.reduced_numel == P * (3 + C + 1 + 1 + 3 + 4) and b.colors_numel == P * C
    assert b.layout[0][0] == "colors" and b.layout[-1][0] == "means2D" and b.layout[-1][1] == b.reduced_numel
    total = 0
    for name, off, cols in b.layout:
        v = b.views[name]
        assert v.shape == (P, cols) and v.is_contiguous()
        assert v.data_ptr() == b.flat.data_ptr() + off * 4
        total += v.numel()
    assert total == b.numel
    b.views["colors"].fill_(2.0)
    assert b.flat.sum().item() == 2.0 * P * C
    out = b.as_backward_out()
    assert out["dL_dmeans3D"].data_ptr() == b.views["means3D"].data_ptr() and out["dL_dcov3D"] is None


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_views, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P, C = 7, 3
    bucket = GradBucket(P, C)
    # stand-in for "backward of view v adds its gradient": view v contributes (v+1) to every entry
    for v in shard_views(n_views, world, rank):
        bucket.flat.add_(float(v + 1))
    local_sum = float(sum(v + 1 for v in shard_views(n_views, world, rank)))
    allreduce_bucket(bucket)
    # the per-view means2D block is not part of the collective: it keeps this rank's own sum
    assert torch.all(bucket.views["means2D"] == local_sum)
    q.put((rank, bucket.flat[:bucket.reduced_numel].clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_sums_the_buckets_of_all_ranks():
    world, n_views = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_views, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = float(sum(v + 1 for v in range(n_views)))
    for r in range(world):
        assert torch.all(results[r] == expect)
    assert torch.equal(results[0], results[1])
