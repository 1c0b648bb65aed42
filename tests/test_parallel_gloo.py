"""world_size-2 gloo test (CPU) of the multi-GPU protocol: sample sharding, the band-packed exchange of
[A^T A | A^T b] (all-reduce for the replicated route, per-band reduce to the block-row-cyclic owner for the distributed
one), the global-N lambda rule, a numpy model of the distributed blocked Cholesky's ownership / broadcast protocol and a
numpy model of the shared conjugate-gradient route.

No GPU here, so the per-rank Gram is formed with numpy and the collectives run over gloo; the layout and ownership
helpers are the product's (superviseddescent_b200/parallel.py mirrors csrc/sd_comm.cu).  The CUDA implementation of the
same protocol is checked on hardware by tests/test_gpu_multi.py."""
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


def _dist_cholesky_model(G, d, m, lam, rank, world, panel):
    """numpy model of sd_solve_gram_dist's protocol (csrc/sd_linalg.cu cholesky_solve with a communicator): block-row-cyclic
    panel ownership, the owner factors its panel and broadcasts the finished rows, every rank updates the rows it owns."""
    from superviseddescent_b200 import parallel
    G = G.astype(np.float64).copy()
    for i in range(d):
        if parallel.panel_owner(i, world, panel) != rank:
            G[i, :] = np.nan                                   # rows of other ranks are undefined on entry
    idx = np.arange(d)
    G[idx, idx] += np.where(idx == d - 1, 0.0, lam)            # regulariser on every (owned) diagonal entry
    for j in range(0, d, panel):
        k = min(panel, d - j)
        owner = parallel.panel_owner(j, world, panel)
        rows = torch.from_numpy(G[j:j + k, j:].copy())
        if rank == owner:
            blk = np.triu(G[j:j + k, j:j + k])
            U11 = np.linalg.cholesky(blk + np.triu(blk, 1).T).T
            P = np.linalg.solve(U11.T, G[j:j + k, j + k:])     # block-row solve, right-hand sides ride along
            rows = torch.from_numpy(np.hstack([U11, P]))
        dist.broadcast(rows, src=owner)                        # the panel broadcast
        G[j:j + k, j:] = rows.numpy()
        P = G[j:j + k, j + k:]
        for i in range(j + k, d):                              # trailing update of the rows this rank owns
            if parallel.panel_owner(i, world, panel) == rank:
                G[i, i:] -= P[:, i - j - k] @ P[:, i - j - k:]
    U, Y = np.triu(G[:, :d]), G[:, d:]
    return np.linalg.solve(U, Y)                               # replicated back substitution


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
    G = np.hstack([Al.T @ Al, Al.T @ Bl]).astype(np.float32)
    n_global = parallel.global_count(e - b)
    G_full = torch.from_numpy(G.copy())
    dist.all_reduce(G_full)                                  # reference: the whole buffer
    # the exchange of the product (csrc/sd_comm.cu): only the upper row bands travel, packed back to back
    band = 5                                                 # ragged last band: 24 = 4*5 + 4
    W = d + m
    off = parallel.band_offsets(d, W, band)
    flat = torch.empty(off[-1], dtype=torch.float32)
    for p, r0 in enumerate(range(0, d, band)):
        flat[off[p]:off[p + 1]] = torch.from_numpy(G[r0:r0 + band, r0:].copy()).reshape(-1)
    # (a) replicated route: all-reduce of the packed bands
    fa = flat.clone()
    dist.all_reduce(fa)
    Ga = G.copy()
    for p, r0 in enumerate(range(0, d, band)):
        Ga[r0:r0 + band, r0:] = fa[off[p]:off[p + 1]].reshape(-1, W - r0).numpy()
    iu = np.triu_indices(d)
    assert np.array_equal(Ga[:, :d][iu], G_full.numpy()[:, :d][iu])             # every element the solve reads ...
    assert np.array_equal(Ga[:, d:], G_full.numpy()[:, d:])                    # ... including the right-hand sides
    X = _solve(Ga, d, 1.5, n_global)
    # (b) distributed route: band p is reduced to rank p % world only, then the distributed factorisation
    Gs = np.full_like(G, np.nan)
    for p, r0 in enumerate(range(0, d, band)):
        piece = flat[off[p]:off[p + 1]].clone()
        dist.reduce(piece, dst=p % world)
        if p % world == rank:
            Gs[r0:r0 + band, r0:] = piece.reshape(-1, W - r0).numpy()
            assert np.array_equal(Gs[r0:r0 + band, r0:], Ga[r0:r0 + band, r0:])
    AtA = np.triu(Ga[:, :d].astype(np.float64))
    lam = 1.5 * np.linalg.norm(AtA + np.triu(AtA, 1).T) / n_global
    Xd = _dist_cholesky_model(np.where(np.isnan(Gs), 0.0, Gs), d, m, lam, rank, world, band)
    x_local = torch.from_numpy(A[b:e, :4].copy())
    gathered = parallel.gather_rows(x_local)
    out.put((rank, n_global, X, gathered.numpy(), Xd))
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


def test_band_layout_and_ownership():
    from superviseddescent_b200 import parallel
    D, W = 1000, 1044
    off = parallel.band_offsets(D, W)
    assert len(off) == 5 and off[1] == 256 * W and off[2] - off[1] == 256 * (W - 256)
    assert off[-1] - off[-2] == (1000 - 768) * (W - 768)
    assert [parallel.panel_owner(r, 3) for r in (0, 255, 256, 511, 512, 768, 999)] == [0, 0, 1, 1, 2, 0, 0]
    assert off[-1] < 0.65 * D * W                            # about half of the buffer travels


def test_two_rank_gram_exchange_and_distributed_solve_match_single_process():
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
    for rank, n_global, X, gathered, Xd in results:
        assert n_global == n
        assert np.max(np.abs(X - X_ref)) <= 1e-4 * np.max(np.abs(X_ref))
        assert np.max(np.abs(Xd - X_ref)) <= 1e-4 * np.max(np.abs(X_ref))      # distributed factorisation, same answer
        assert np.array_equal(gathered, A[:, :4])
    assert np.array_equal(results[0][2], results[1][2])      # replicated solve: bit-identical on every rank
    assert np.array_equal(results[0][4], results[1][4])      # distributed solve: every rank ends with the same X


def _cg_worker(rank, world, port, n, d, m, out):
    """numpy model of route 2 (csrc/sd_linalg.cu solve_gram_impl + csrc/sd_cg.cu): global centring, all-reduced Gram, bias column
    eliminated first, every rank downdates ONLY what its slab of the product reads (everything else is poisoned with NaN here),
    strip-major copy of the slab built from the upper triangle, CG with one all-reduce of Q per iteration."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from superviseddescent_b200 import parallel
    rng = np.random.default_rng(9)
    A = np.minimum(np.abs(rng.standard_normal((n, d))) * 0.08, 0.2)
    A[:, -1] = 1.0
    B = 0.05 * rng.standard_normal((n, m))
    b, e = parallel.shard_range(n, world, rank)
    # sd_centre_features: global column means (one all-reduce), bias column untouched
    sums = torch.from_numpy(A[b:e].sum(0))
    dist.all_reduce(sums)
    mu = sums.numpy() / n
    mu[-1] = 0.0
    Ac = A[b:e] - mu
    G = torch.from_numpy(np.hstack([Ac.T @ Ac, Ac.T @ B[b:e]]))
    dist.all_reduce(G)
    G = np.triu(G.numpy()[:, :d]), G.numpy()[:, d:]                # upper triangle + right-hand sides, as the exchange delivers
    U, R = G
    # lambda from the norm of the UNcentred matrix; the ranks split the rows of the sum (route 2: share_norm)
    full = U + np.triu(U, 1).T
    un = full + n * np.outer(mu, mu)
    un[:, -1] += n * mu
    un[-1, :] += n * mu
    part = torch.tensor([float(sum((un[i, i:] ** 2).sum() * 2 - un[i, i] ** 2 for i in range(d) if (i // 4) % world == rank))], dtype=torch.float64)
    dist.all_reduce(part)
    lam = 1.5 * np.sqrt(part.item()) / n
    assert abs(lam - 1.5 * np.linalg.norm(un) / n) <= 1e-12 * lam
    # last column first: s = bias column, pivot = n; what remains is (d-1) x (d-1)
    nn = d - 1
    s_col, piv, rb = U[:nn, -1].copy(), U[-1, -1], R[-1].copy()
    S = U[:nn, :nn] + lam * np.eye(nn)
    Y = R[:nn].copy()
    k0, k1 = parallel.cg_slab(nn, world, rank, align=4)
    need = np.zeros((nn, nn), bool)
    need[k0:k1, :] = True
    need[:, k0:k1] = True
    need &= np.triu(np.ones((nn, nn), bool))
    S = np.where(need, S - np.outer(s_col, s_col) / piv, np.nan)     # partial downdate: the rest is never prepared ...
    Y = Y - np.outer(s_col, rb) / piv
    # cg_pack_kernel: T[strip][k - k0][c] = S[k][strip*W + c] from the upper triangle only
    W = 8
    kp = (k1 - k0 + 3) // 4 * 4
    T = np.zeros(((nn + W - 1) // W, kp, W))
    for k in range(k0, k1):
        for j in range(nn):
            T[j // W, k - k0, j % W] = S[k, j] if j >= k else S[j, k]
    assert not np.isnan(T).any()                                       # ... and never read
    # lockstep CG on all right-hand sides; Q = sum over ranks of S[k0:k1, :]^T P[k0:k1, :]
    X = np.zeros_like(Y); Rr = Y.copy(); P = Y.copy()
    rs = (Rr * Rr).sum(0); bb = rs.copy()
    its = 0
    for its in range(1, 200):
        Q = np.zeros_like(P)
        for st in range(T.shape[0]):
            cols = slice(st * W, min(st * W + W, nn))
            Q[cols] = T[st, :k1 - k0, :cols.stop - cols.start].T @ P[k0:k1]
        Qt = torch.from_numpy(Q)
        dist.all_reduce(Qt)
        Q = Qt.numpy()
        alpha = rs / (P * Q).sum(0)
        X += alpha * P
        Rr -= alpha * Q
        rn = (Rr * Rr).sum(0)
        if np.sqrt(rn / bb).max() <= 1e-10:
            break
        P = Rr + (rn / rs) * P
        rs = rn
    bias = (rb - s_col @ X) / piv - mu[:nn] @ X                        # bias_finish_kernel, shifted back to uncentred rows
    out.put((rank, its, np.vstack([X, bias])))
    dist.barrier()
    dist.destroy_process_group()


def test_cg_slab_covers_rows_once():
    from superviseddescent_b200 import parallel
    for n in (1, 15, 16, 17050, 52700):
        for w in (1, 2, 3, 8):
            slabs = [parallel.cg_slab(n, w, r) for r in range(w)]
            assert slabs[0][0] == 0 and slabs[-1][1] == n
            assert all(slabs[i][1] == slabs[i + 1][0] for i in range(w - 1))
            assert all(k0 % 16 == 0 or k0 == n for k0, _ in slabs)       # empty slabs start at n
    assert parallel.cg_slab(17050, 8, 7) == (15008, 17050)                           # 2,144 rows per rank at config 4


def test_two_rank_shared_cg_route_matches_direct_solve():
    n, d, m, world = 400, 61, 6, 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cg_worker, args=(r, world, port, n, d, m, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(9)
    A = np.minimum(np.abs(rng.standard_normal((n, d))) * 0.08, 0.2)
    A[:, -1] = 1.0
    B = 0.05 * rng.standard_normal((n, m))
    X_ref = _solve(np.hstack([A.T @ A, A.T @ B]), d, 1.5, n)
    for rank, its, X in results:
        assert 1 < its < 100
        assert np.max(np.abs(X - X_ref)) <= 1e-8 * np.max(np.abs(X_ref))
    assert np.array_equal(results[0][2], results[1][2])      # every rank ends with the same model
