"""Multi-GPU plumbing of the hot path.  One process per GPU; torch.distributed only.

Training shards the SAMPLES (rows of A): rank r owns rows [r*N/G, (r+1)*N/G) of the images, x and x_gt for
the whole run.  Per cascade level every rank forms its partial [A^T A | A^T b] and ONE all-reduce (sum,
fp32) over NVLink merges them; the lambda rule uses the GLOBAL sample count and the solve runs redundantly
on every rank (deterministic, no broadcast).  Inference shards the face batch with no collective at all.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Rows [begin, end) owned by `rank`: contiguous, sizes differ by at most one, every row owned once."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    base, extra = divmod(n, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def global_count(n_local: int, group=None, device=None) -> int:
    """Total number of training rows over the group (the N of the MatrixNorm lambda rule, regressors.hpp:135)."""
    if group is None and not (dist.is_available() and dist.is_initialized()):
        return n_local
    t = torch.tensor([n_local], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


def allreduce_gram(G: torch.Tensor, group=None, D: Optional[int] = None, band: int = 1024) -> torch.Tensor:
    """The one collective per cascade level: sums the packed [A^T A | A^T b] buffer over the ranks, in place.

    Only the upper triangle of A^T A (and the right-hand-side columns) is ever read by the solve, so with `D` given the
    buffer is sent as row bands of `band` rows, each from its first diagonal column to the end of the row: about half the
    bytes of the full buffer (0.62 instead of 1.17 GB for config 4, 5.7 instead of 11.1 GB for config 5).  The bands are
    staged through one contiguous buffer (two HBM-speed copies) because the collective needs contiguous memory.  Elements
    below a band's first column keep this rank's partial sums; nothing reads them."""
    if group is None and not (dist.is_available() and dist.is_initialized()):
        return G
    rows, ld = G.shape
    if D is None or band < 1 or rows <= 2 * band:
        dist.all_reduce(G, op=dist.ReduceOp.SUM, group=group)
        return G
    starts = list(range(0, rows, band))
    sizes = [(min(b0 + band, rows) - b0) * (ld - b0) for b0 in starts]
    flat = torch.empty(sum(sizes), dtype=G.dtype, device=G.device)
    off = 0
    for b0, sz in zip(starts, sizes):
        b1 = min(b0 + band, rows)
        flat[off:off + sz].view(b1 - b0, ld - b0).copy_(G[b0:b1, b0:])
        off += sz
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for b0, sz in zip(starts, sizes):
        b1 = min(b0 + band, rows)
        G[b0:b1, b0:].copy_(flat[off:off + sz].view(b1 - b0, ld - b0))
        off += sz
    return G


def gather_rows(x_local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather of the per-rank landmark rows (only needed when a training callback wants the full current_x,
    superviseddescent.hpp:217)."""
    if group is None and not (dist.is_available() and dist.is_initialized()):
        return x_local
    world = dist.get_world_size(group)
    counts = [torch.zeros(1, dtype=torch.int64, device=x_local.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([x_local.shape[0]], dtype=torch.int64, device=x_local.device), group=group)
    counts = [int(c.item()) for c in counts]
    most = max(counts)
    padded = torch.zeros((most, x_local.shape[1]), dtype=x_local.dtype, device=x_local.device)   # all_gather needs equal shapes
    padded[:x_local.shape[0]] = x_local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
