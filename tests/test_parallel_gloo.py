"""world_size-2 gloo test (CPU) of the multi-GPU host logic: sample sharding + the one all-reduce of
[A^T A | A^T b] per level + the global-N lambda rule + replicated solve.

No GPU here, so the per-rank Gram is formed with numpy; everything else is the product code in
superviseddescent_b200/parallel.py that the NCCL path runs unchanged."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _solve(G, D, lam_param, n_global):
    AtA = G[:, :D].astype(np.float64)
    AtA = np.triu(AtA) + np.triu(AtA, 1).T
    lam = lam_param * np.linalg.norm(AtA) / n_global            # regressors.hpp:133-136
    reg = np.eye(D) * lam
    reg[-1, -1] = 0.0                                            # bias row not regularised (:143-146)
    return np.linalg.solve(AtA + reg, G[:, D:].astype(np.float64))


def _worker(rank, world, port, n, d, m, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from superviseddescent_b200 import parallel
    rng = np.random.default_rng(5)
    A = rng.random((n, d)).astype(np.float32)
    A[:, -1] = 1.0
    B = rng.standard_normal((n, m)).astype(np.float32)
    b, e = parallel.shard_range(n, world, rank)
    Al, Bl = A[b:e].astype(np.float64), B[b:e].astype(np.float64)
    G = torch.from_numpy(np.hstack([Al.T @ Al, Al.T @ Bl]).astype(np.float32))
    n_global = parallel.global_count(e - b)
    G_full = G.clone()
    parallel.allreduce_gram(G_full)                          # whole buffer
    parallel.allreduce_gram(G, None, d, band=5)              # upper row bands only (ragged last band: 24 = 4*5 + 4)
    iu = np.triu_indices(d)
    assert np.array_equal(G.numpy()[:, :d][iu], G_full.numpy()[:, :d][iu])       # every element the solve reads ...
    assert np.array_equal(G.numpy()[:, d:], G_full.numpy()[:, d:])              # ... including the right-hand sides
    X = _solve(G.numpy(), d, 1.5, n_global)
    x_local = torch.from_numpy(A[b:e, :4].copy())
    gathered = parallel.gather_rows(x_local)
    out.put((rank, n_global, X, gathered.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions_rows():
    from superviseddescent_b200 import parallel
    for n in (0, 1, 7, 10000, 100001):
        for w in (1, 2, 3, 8):
            ranges = [parallel.shard_range(n, w, r) for r in range(w)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(10, 2, 2)


def test_two_rank_gram_allreduce_matches_single_process():
    n, d, m, world = 301, 24, 6, 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, d, m, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(5)
    A = rng.random((n, d)).astype(np.float32)
    A[:, -1] = 1.0
    B = rng.standard_normal((n, m)).astype(np.float32)
    A64, B64 = A.astype(np.float64), B.astype(np.float64)
    X_ref = _solve(np.hstack([A64.T @ A64, A64.T @ B64]), d, 1.5, n)
    for rank, n_global, X, gathered in results:
        assert n_global == n
        assert np.max(np.abs(X - X_ref)) <= 1e-4 * np.max(np.abs(X_ref))
        assert np.array_equal(gathered, A[:, :4])
    assert np.array_equal(results[0][2], results[1][2])      # replicated solve: bit-identical on every rank
