"""Data-parallel inference over several B200s: one process per GPU, independent replicas, result gather only.

The forward has no cross-image dependency (eval-mode BN, per-image SE / attention / top-k, per-ROI head), so the
batch is split contiguously across ranks and the only collective is the gather of the small result tensors
(`keypoints` 136 B + `visibilities` 204 B per person; the 213 KB/person heat maps only on request) over
NCCL / NVLink.  The reference itself has no distributed code.
"""
from typing import Dict, Optional

import torch
import torch.distributed as dist


def shard_bounds(batch_size: int, rank: int, world: int):
    """Contiguous split: rank r owns images [lo, hi); the first (batch_size % world) ranks get one extra image."""
    base, rem = divmod(batch_size, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: Dict, rank: int, world: int) -> Dict:
    """Slice 'image' and 'bboxes' ([B,P,4] tensor or [tensor] list form) of a global batch for this rank."""
    x = batch['image']
    lo, hi = shard_bounds(x.size(0), rank, world)
    out = {'image': x[lo:hi]}
    if 'bboxes' in batch:
        b = batch['bboxes']
        if isinstance(b, list) and len(b) > 0 and torch.is_tensor(b[0]) and b[0].dim() == 3:
            out['bboxes'] = [b[0][lo:hi]]
        elif torch.is_tensor(b):
            out['bboxes'] = b[lo:hi]
        else:
            out['bboxes'] = b
    return out


def _gather_rows(t: torch.Tensor, counts, group) -> torch.Tensor:
    """all_gather of tensors that differ only in dim 0 (ragged split): pad to the max row count, gather, trim."""
    world = len(counts)
    mx = max(counts)
    if t.size(0) < mx:
        pad = torch.zeros((mx - t.size(0),) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        t = torch.cat([t, pad], 0)
    t = t.contiguous()
    full = torch.empty((world * mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(full, t, group=group)
    return torch.cat([full[r * mx:r * mx + counts[r]] for r in range(world)], 0)


def gather_outputs(local: Dict, global_batch: int, group=None, gather_heatmap: bool = False) -> Dict:
    """Assemble the global result on every rank from per-rank forward outputs (all ranks must use the same P)."""
    world = dist.get_world_size(group)
    counts = [shard_bounds(global_batch, r, world)[1] - shard_bounds(global_batch, r, world)[0] for r in range(world)]
    out = {'keypoints': _gather_rows(local['keypoints'], counts, group),
           'visibilities': _gather_rows(local['visibilities'], counts, group)}
    if gather_heatmap:
        out['heatmap'] = _gather_rows(local['heatmap'], counts, group)
    return out


def _persons(batch: Dict) -> int:
    """Padded person count P of the global batch (0 when no boxes are supplied)."""
    b = batch.get('bboxes') if isinstance(batch, dict) else None
    if isinstance(b, list) and len(b) > 0 and torch.is_tensor(b[0]):
        return int(b[0].size(-2)) if b[0].dim() >= 2 else 0
    if torch.is_tensor(b) and b.dim() == 3:
        return int(b.size(1))
    return 0


def _empty_shard_result(model, batch: Dict) -> Dict:
    """Result of a rank whose shard holds no image: zero rows with the trailing shapes the other ranks produce
    ([b,P,1,K,2] / [b,P,1,K,3] / [b,P,K,56,56]; the reference's rank-changing no-box shapes when P == 0), so the
    gather sees tensors that differ in dim 0 only.  The model is not called (its 0-image path has other shapes)."""
    K = int(getattr(model, 'num_keypoints', 17))
    try:
        dev = next(model.parameters()).device
    except (AttributeError, StopIteration, TypeError):
        dev = batch['image'].device
    P = _persons(batch)
    if P == 0:
        return {'keypoints': torch.zeros(0, 1, K, 2, device=dev), 'visibilities': torch.zeros(0, 1, K, device=dev),
                'heatmap': torch.zeros(0, 1, K, 56, 56, device=dev), 'boxes': []}
    return {'keypoints': torch.zeros(0, P, 1, K, 2, device=dev), 'visibilities': torch.zeros(0, P, 1, K, 3, device=dev),
            'heatmap': torch.zeros(0, P, K, 56, 56, device=dev), 'boxes': []}


@torch.no_grad()
def predict_data_parallel(model, batch: Dict, group=None, gather_heatmap: bool = False, presharded: bool = False,
                          global_batch: Optional[int] = None) -> Dict:
    """Every rank passes the same global batch (default), or -- `presharded=True` -- only its own contiguous slice of a
    `global_batch`-image batch split by `shard_bounds`; returns the gathered global result plus the rank's own under 'local'.

    Works with any backend torch.distributed offers: NCCL on the GPUs (product path), gloo in the CPU unit tests
    (with a stand-in `model` callable -- the CUDA model itself has no CPU path).
    """
    if not dist.is_initialized():
        return model.predict(batch) if hasattr(model, 'predict') else model(batch)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if presharded:
        if global_batch is None:
            raise ValueError("presharded=True needs global_batch (the image count over all ranks)")
        lo, hi = shard_bounds(global_batch, rank, world)
        if batch['image'].size(0) != hi - lo:
            raise ValueError(f"rank {rank} owns images [{lo}, {hi}) of {global_batch} but was given {batch['image'].size(0)}")
        local, total = batch, global_batch
    else:
        local, total = shard_batch(batch, rank, world), batch['image'].size(0)
    if local['image'].size(0) == 0:
        res = _empty_shard_result(model, batch)         # global batch < world size: this rank owns no image
    else:
        res = model.predict(local) if hasattr(model, 'predict') else model(local)
    out = gather_outputs(res, total, group, gather_heatmap)
    out['local'] = res
    return out
