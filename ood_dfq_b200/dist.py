"""Data-parallel glue of the quantisation path (one process per GPU, torch.distributed / NCCL).

The path shards trivially over the batch: every image's fake-quant is independent given the
replicated ranges and weights.  Collectives that exist:

* gradients: one flat all-reduce per step (``step.FlatGrads``);
* BN partial sums: one packed all-reduce per hooked forward (``bns.BNStatLoss(sync=True)``);
* activation ranges: ``reduce_minmax`` below, once, when calibration ends.

All three are device-agnostic torch.distributed calls, so the host logic is covered by
world_size-2 gloo tests on CPU (tests/test_dist_gloo.py).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def _range_modules(model, act_types):
    if act_types is None:
        from .quantization_utils.quant_modules import QuantAct
        act_types = (QuantAct,)
    return [m for m in model.modules() if isinstance(m, act_types)]


def reduce_minmax(model, group=None, act_types=None):
    """Average the calibrated activation ranges over ranks in ONE all-reduce.

    Reference: ``Trainer.reduce_minmax`` (trainer_direct.py:368-374) issues two 4-byte
    all-reduces per QuantAct (34-38 NCCL latencies) and divides by the world size; the
    semantics -- arithmetic mean of the per-rank EMA states, not a min/max -- are kept.
    """
    mods = _range_modules(model, act_types)
    if not mods or not (dist.is_available() and dist.is_initialized()):
        return model
    world = dist.get_world_size(group)
    packed = torch.cat([m.x_min.reshape(-1)[:1] for m in mods] + [m.x_max.reshape(-1)[:1] for m in mods])
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    packed = packed / world
    n = len(mods)
    for i, m in enumerate(mods):
        m.x_min = packed[i: i + 1].clone()
        m.x_max = packed[n + i: n + i + 1].clone()
    return model


def shard_batch(global_batch: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Contiguous per-rank slice of a global batch (what DistributedSampler gives each rank in order)."""
    per = global_batch.shape[0] // world
    return global_batch[rank * per: (rank + 1) * per]
