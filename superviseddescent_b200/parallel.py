"""Multi-GPU plumbing of the training path.  One process per GPU.

Training shards the SAMPLES (rows of A): rank r owns rows [r*N/G, (r+1)*N/G) of the images, x and x_gt for the whole
run.  Per cascade level every rank forms its partial [A^T A | A^T b]; the exchange and the solve are the C ABI's
(include/sd_b200.h, "multi-GPU training"): NCCL collectives issued by libsd_b200.so on the context's stream.

  replicated  : sd_allreduce_gram (upper row bands only) + sd_solve_gram on every rank
  distributed : sd_reduce_scatter_gram (band p -> rank p % G) + sd_solve_gram_dist (block-row-cyclic blocked Cholesky,
                panel broadcast over NVLink, every rank updates the block rows it owns)

torch.distributed is only the bootstrap here: it carries the 128-byte NCCL id from rank 0 to the other ranks (any
backend -- a C++ host would use MPI or a file).  Inference shards the face batch with no collective at all.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

PANEL = 256          # rows per ownership unit == one Cholesky panel (csrc/sd_linalg.cu: two 128-blocks)
DIST_SOLVE_MIN_D = 24000   # feature dimension from which the distributed factorisation pays


def shard_range(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Rows [begin, end) owned by `rank`: contiguous, sizes differ by at most one, every row owned once."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    base, extra = divmod(n, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def panel_owner(row: int, world_size: int, panel: int = PANEL) -> int:
    """Rank that owns global row `row` of [AtA|Atb] in the distributed factorisation (block-row-cyclic)."""
    return (row // panel) % world_size


def cg_slab(n: int, world_size: int, rank: int, align: int = 16) -> Tuple[int, int]:
    """Rows [k0, k1) of the n x n system whose part of the product S P `rank` computes in the shared conjugate-gradient route
    (multiples of `align` rows: one pipeline stage of the tensor-core product).  Mirror of sd_cg_slab() in csrc/sd_internal.cuh."""
    if world_size <= 1:
        return 0, n
    per = ((n + world_size - 1) // world_size + align - 1) // align * align
    return min(rank * per, n), min((rank + 1) * per, n)


def band_offsets(D: int, W: int, band: int = PANEL) -> List[int]:
    """Offsets (in floats) of the packed row bands that travel in the Gram exchange: band p = rows [p*band, ...) from
    column p*band to W.  Mirror of band_layout() in csrc/sd_comm.cu; offsets[-1] is the total."""
    off = [0]
    for r0 in range(0, D, band):
        rows = min(band, D - r0)
        off.append(off[-1] + rows * (W - r0))
    return off


def global_count(n_local: int, group=None, device=None) -> int:
    """Total number of training rows over the group (the N of the MatrixNorm lambda rule, regressors.hpp:135)."""
    if group is None and not (dist.is_available() and dist.is_initialized()):
        return n_local
    t = torch.tensor([n_local], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


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


class Communicator:
    """sd_comm of the C ABI, bootstrapped from a torch.distributed process group (which only carries the NCCL id)."""

    def __init__(self, ctx, group=None):
        from . import _capi
        self.ctx = ctx
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)
        lib = _capi.lib()
        ident = (C.c_uint8 * 128)()
        if self.rank == 0 and self.size > 1:
            rc = lib.sd_comm_get_unique_id(ident)
            if rc:
                raise _capi.SdError(rc, "sd_comm_get_unique_id failed (is libnccl.so.2 loadable?)")
        if self.size > 1:
            backend = dist.get_backend(group)
            t = torch.tensor(list(bytes(ident)), dtype=torch.uint8, device=(f"cuda:{ctx.device}" if backend == "nccl" else "cpu"))
            dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            ident = (C.c_uint8 * 128)(*t.cpu().tolist())
        self._h = C.c_void_p()
        rc = lib.sd_comm_create(ctx.h, ident, self.rank, self.size, C.byref(self._h))
        if rc:
            raise _capi.SdError(rc, lib.sd_last_error(ctx.h).decode())

    @property
    def h(self):
        return self._h

    def sum_int(self, value: int) -> int:
        from . import _capi
        v = C.c_int64(int(value))
        rc = _capi.lib().sd_comm_sum_int64(self.ctx.h, self._h, C.byref(v))
        if rc:
            raise _capi.SdError(rc, _capi.lib().sd_last_error(self.ctx.h).decode())
        return int(v.value)

    def allgather_rows(self, x_local: torch.Tensor, counts: Optional[List[int]] = None) -> torch.Tensor:
        """Equal row counts per rank: one sd_comm_allgather; ragged: padded to the largest shard."""
        from . import _capi
        n = x_local.shape[0]
        if counts is None:
            counts = [0] * self.size
            counts[self.rank] = n
            counts = [self.sum_int(c) for c in counts]
        most = max(counts)
        send = torch.zeros((most, x_local.shape[1]), dtype=x_local.dtype, device=x_local.device)
        send[:n] = x_local
        recv = torch.empty((self.size, most, x_local.shape[1]), dtype=x_local.dtype, device=x_local.device)
        rc = _capi.lib().sd_comm_allgather(self.ctx.h, self._h, C.c_void_p(send.data_ptr()), C.c_size_t(send.numel() * send.element_size()),
                                           C.c_void_p(recv.data_ptr()))
        if rc:
            raise _capi.SdError(rc, _capi.lib().sd_last_error(self.ctx.h).decode())
        return torch.cat([recv[r, :counts[r]] for r in range(self.size)], dim=0)

    def close(self):
        from . import _capi
        if self._h:
            self.ctx.sync()
            _capi.lib().sd_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
