"""GPU (-m gpu): the view-parallel path on real hardware — every rank renders its share of a batch of views into its flat
gradient bucket, one NCCL sum-allreduce, and the result must equal the SUM OVER ALL VIEWS of the gradients the reference's own
CUDA build (oracle/_ref) returns for each view.

Two ways of running the same check:
  * under `torchrun --nproc-per-node N` (WORLD_SIZE > 1): N ranks, one GPU each, NCCL — the real thing;
        gpurun --gpus 2 -- python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
               --master-port 29511 -m pytest tests/test_dist_gpu.py -m gpu -q
  * plain `pytest -m gpu` on one GPU: the same batch, all views on the one rank (no collective), which still pins the
    accumulate-into-bucket arithmetic to the reference build.
"""
import os

import numpy as np
import pytest
import torch

import _ref_utils as ru
from _cases import GRAD_KEYS
from gscream_b200 import _lib, scenes
from gscream_b200.dist import GradBucket, allreduce_bucket, render_views_into_bucket, shard_views

pytestmark = pytest.mark.gpu
REL = 1e-5
BUCKET_OF = {"dL_dmeans3D": "means3D", "dL_dmeans2D": "means2D", "dL_dcolors": "colors", "dL_dopacity": "opacities",
             "dL_duncertainty": "uncertainties", "dL_dscales": "scales", "dL_drotations": "rotations"}


def _init():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    return world, rank, torch.device("cuda", local)


@pytest.mark.parametrize("C", [32, 3])
def test_allreduced_bucket_equals_sum_of_reference_gradients(C):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    _lib.load()
    if not ru.ref_available(C):
        pytest.skip("oracle/_ref/dgr%d not built" % C)
    world, rank, dev = _init()
    P, W, H = 60000, 480, 272
    n_views = 2 * world
    scene = scenes.make_scene(P, W, H, C, 4242 + C, scale_mult=2.0, bg_value=0.1 if C == 3 else 0.0)
    cams = [scenes.make_camera(W, H, yaw_deg=3.0 * (v - (n_views - 1) / 2)) for v in range(n_views)]
    ups = [scenes.make_upstream_grads(C, W, H, 9000 + v) for v in range(n_views)]

    # ---- ours: local views into the bucket, one collective ----
    s = {k: v.to(dev) for k, v in scene.items()}
    mine = shard_views(n_views, world, rank)
    bucket = GradBucket(P, C, device=dev)
    bucket.flat.fill_(123.0)   # stale contents: overwrite=True must not need a zeroed bucket
    render_views_into_bucket(s, [{k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in cams[v].items()} for v in mine],
                             [tuple(t.to(dev) for t in ups[v]) for v in mine], bucket, overwrite=True)
    allreduce_bucket(bucket)
    torch.cuda.synchronize()

    # ---- checker: the reference build on EVERY view (each rank recomputes the full sum on its own GPU) ----
    ref_mod = ru.load_ref(C)
    total = {k: 0.0 for k in GRAD_KEYS}
    spread = {k: 0.0 for k in GRAD_KEYS}
    local_m2d, local_spread = 0.0, 0.0
    for v in range(n_views):
        r = ru.run_impl(ref_mod, scene, cams[v], ups[v], device=str(dev))
        r2 = ru.run_impl(ref_mod, scene, cams[v], ups[v], device=str(dev))
        for k in GRAD_KEYS:
            total[k] = total[k] + r[k].astype(np.float64)
            spread[k] += float(np.abs(r2[k] - r[k]).max())
        if v in mine:
            local_m2d = local_m2d + r["dL_dmeans2D"].astype(np.float64)
            local_spread += float(np.abs(r2["dL_dmeans2D"] - r["dL_dmeans2D"]).max())
    got = bucket.views["means2D"].cpu().numpy()
    assert np.abs(got - local_m2d).max() <= REL * np.abs(local_m2d).max() * len(mine) ** 0.5 + 8.0 * local_spread, "means2D (local block)"
    for k in GRAD_KEYS:
        if k == "dL_dmeans2D":
            continue  # per-view densification signal: stays local (checked below)
        got = bucket.views[BUCKET_OF[k]].cpu().numpy()
        ref = total[k]
        rel = REL if k in ("dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_duncertainty") else 3 * REL
        tol = rel * np.abs(ref).max() * n_views ** 0.5 + 8.0 * spread[k]
        err = float(np.abs(got - ref).max())
        assert err <= tol, (k, "rank %d of %d" % (rank, world), err, float(tol))
    if world > 1:
        # every rank holds the same parameter gradients after the collective
        import torch.distributed as dist
        mine_sum = bucket.flat[:bucket.reduced_numel].double().sum().reshape(1)
        gathered = [torch.zeros_like(mine_sum) for _ in range(world)]
        dist.all_gather(gathered, mine_sum)
        assert all(torch.equal(g, gathered[0]) for g in gathered)
