"""Multi-GPU plumbing for the batch path: proofs are independent, so a batch shards by proof index with no collective on
the data path.  One process per GPU (torchrun); the only exchange is the final gather of verdict bits.

The C library can also shard one call over several devices of one process (bn254v_init(devices, n)); this module is the
one-process-per-GPU form used by bench.py.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous index range [lo, hi) of rank `rank` (same rule as the C library's per-device split)."""
    return n * rank // world, n * (rank + 1) // world


def pack_verdicts(status: np.ndarray) -> np.ndarray:
    """status bytes -> accept bits (1 = Ok(true)), packed 8 per byte, padded to a whole byte."""
    return np.packbits(np.asarray(status) == 0)


def unpack_verdicts(bits: np.ndarray, n: int) -> np.ndarray:
    return np.unpackbits(np.asarray(bits, dtype=np.uint8))[:n].astype(bool)


def gather_verdicts(status: np.ndarray, n_total: int, dist, device=None) -> np.ndarray:
    """All ranks contribute the verdict bits of their shard (in rank order, shards from `shard_bounds`); every rank
    gets the accept bits of the whole batch.  `dist` is torch.distributed (nccl on GPUs, gloo in the CPU tests)."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    max_bytes = max((hi - lo + 7) // 8 for lo, hi in sizes)
    mine = np.zeros(max_bytes, dtype=np.uint8)
    packed = pack_verdicts(status)
    mine[:packed.size] = packed
    t = torch.from_numpy(mine)
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    res = np.zeros(n_total, dtype=bool)
    for r, (lo, hi) in enumerate(sizes):
        res[lo:hi] = unpack_verdicts(out[r].cpu().numpy(), hi - lo)
    return res
