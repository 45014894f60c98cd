"""N>1 host logic on CPU: world_size-2 gloo run of the shard / gather path (the CUDA model is replaced by a
deterministic stand-in because the product has no CPU compute path)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _stand_in(batch):
    x, b = batch['image'], batch['bboxes']
    B, P = b.shape[:2]
    m = x.mean(dim=(1, 2, 3)).view(B, 1, 1, 1, 1)
    return {'keypoints': m.expand(B, P, 1, 17, 2).clone() + b[:, :, None, None, :2],
            'visibilities': torch.ones(B, P, 1, 17, 3) * b[:, :, None, None, 2:3],
            'heatmap': m.view(B, 1, 1, 1, 1).expand(B, P, 17, 4, 4).clone()}


def _worker(rank, world, port, batch_size, q):
    sys.path.insert(0, os.path.join(ROOT, "keypoint-detection_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dll.parallel import predict_data_parallel, shard_bounds
    g = torch.Generator().manual_seed(0)
    batch = {'image': torch.randn(batch_size, 1, 8, 8, generator=g), 'bboxes': torch.rand(batch_size, 3, 4, generator=g)}
    small = batch_size < world                 # a rank without images: its empty result carries the model's 56x56 heat-map shape
    out = predict_data_parallel(_stand_in, batch, gather_heatmap=not small)
    full = _stand_in(batch)
    ok = all(torch.equal(out[k], full[k]) for k in (('keypoints', 'visibilities') if small else ('keypoints', 'visibilities', 'heatmap')))
    lo, hi = shard_bounds(batch_size, rank, world)
    ok = ok and out['local']['keypoints'].shape[0] == hi - lo
    # pre-sharded form (what bench.py runs at N > 1): the rank passes only its own slice
    mine = {'image': batch['image'][lo:hi], 'bboxes': batch['bboxes'][lo:hi]}
    out2 = predict_data_parallel(_stand_in, mine, gather_heatmap=not small, presharded=True, global_batch=batch_size)
    ok = ok and torch.equal(out2['keypoints'], full['keypoints']) and torch.equal(out2['visibilities'], full['visibilities'])
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("batch_size", [8, 7, 1])
def test_two_rank_shard_and_gather(batch_size):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + batch_size
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch_size, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_shard_bounds_cover_the_batch():
    sys.path.insert(0, os.path.join(ROOT, "keypoint-detection_b200"))
    from dll.parallel import shard_bounds, shard_batch
    for b in (1, 7, 8, 4096):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(b, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == b
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    batch = {'image': torch.zeros(5, 1, 4, 4), 'bboxes': [torch.arange(5 * 2 * 4.).view(5, 2, 4)]}
    s = shard_batch(batch, 1, 2)
    assert s['image'].shape[0] == 2 and torch.equal(s['bboxes'][0], batch['bboxes'][0][3:5])
