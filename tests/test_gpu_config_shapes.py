"""GPU parity of one training level at the SHAPES of BASELINE.json's training configs (the kernels bench.py times):
config 4: 22 landmarks, 5x5 cells of 11 px, K = 9  -> hog_patch_kernel<9,5,11>, D = 17,051 (134 Cholesky blocks)
config 5: 68 landmarks, same cells                 -> D = 52,701 (412 blocks)
on a reduced number of samples (the oracle's HOG and a float64 solve must finish in seconds; the feature dimension, which is
what selects kernels, tile schedules and panel counts, is the full one).

Truth: float64, from the oracle's float32 features.  N < D here (as in config 4 itself), so the float64 solve uses the ridge
identity on the CENTRED system -- with the bias column unregularised (regressors.hpp:143-146)
    w = Ac^T (Ac Ac^T + lambda I_N)^-1 bc ,   bias = mean(b) - mean(A) w ,   lambda = 1.5 ||A A^T||_F / N  (= ||A^T A||_F)
which is the exact minimiser of the reference's system (A^T A + Lambda) X = A^T b at N x N cost.

Gates (north_star): features, weights AND updated landmarks <= 1e-4 relative (max-norm / max-abs).  The float32 LAPACK
partial-pivot LU of the same regularised Gram (what Eigen's PartialPivLU does in the reference, blocked) is solved beside it
at config 4 and its own distance to the float64 truth is printed: it is the accuracy the reference itself has."""
import os
import time

import numpy as np
import pytest

import synth
from conftest import rel_err

pytestmark = pytest.mark.gpu

CFG = {
    "config4": dict(n=2000, size=128, landmarks=22, seed=2024),
    "config5": dict(n=400, size=256, landmarks=68, seed=2025),
}


def _shape_model(oracle, golden, landmarks):
    if landmarks == 22:
        om = oracle.Model(golden.model_path)
        return om.mean, om.landmark_ids, om.right_ids, om.left_ids, om.right_idx, om.left_idx
    mean = np.asarray(golden.mean68, dtype=np.float32).reshape(-1)
    ids = [str(i) for i in range(1, 69)]
    return mean, ids, ["37", "40"], ["43", "46"], [36, 39], [42, 45]


def _truth_dual(A, b, lam_param):
    n = A.shape[0]
    A64, b64 = A.astype(np.float64), b.astype(np.float64)
    Kf = A64 @ A64.T
    lam = float(np.float32(lam_param) * np.float32(np.linalg.norm(Kf)) / np.float32(n))      # regressors.hpp:135 (float)
    Aw = A64[:, :-1]
    mu, mb = Aw.mean(axis=0), b64.mean(axis=0)
    Ac, bc = Aw - mu, b64 - mb
    alpha = np.linalg.solve(Ac @ Ac.T + lam * np.eye(n), bc)
    w = Ac.T @ alpha
    return np.vstack([w, (mb - mu @ w)[None]]), lam


@pytest.mark.parametrize("name", ["config4", "config5"])
def test_training_level_at_config_shape(sd, oracle, golden, name):
    import torch
    cfg = CFG[name]
    n, size, L = cfg["n"], cfg["size"], cfg["landmarks"]
    mean, ids, right, left, ridx, lidx = _shape_model(oracle, golden, L)
    images = synth.smooth_images(n, size, size, seed=cfg["seed"])
    rng = np.random.default_rng(cfg["seed"])
    m = int(round(size * 0.05))
    box = (m, m, size - 2 * m, size - 2 * m)
    x0 = np.tile(oracle.align_mean(mean, box), (n, 1)).astype(np.float32)
    x_gt = np.stack([oracle.align_mean(mean, box, 1.0 + rng.normal(0, 0.04), 1.0 + rng.normal(0, 0.04), rng.normal(0, 0.04), rng.normal(0, 0.04))
                     for _ in range(n)]).astype(np.float32)
    hp, ohp = sd.HoGParam(1, 5, 11, 9, 1.0), oracle.HogParam(1, 5, 11, 9, 1.0)
    t0 = time.time()
    A_ref = oracle.hog_transform_batch(images, x0, ohp, ridx, lidx, threads=min(32, os.cpu_count() or 1))
    t_hog = time.time() - t0
    D = A_ref.shape[1]
    assert D == L * 25 * 31 + 1
    ied = np.array([oracle.get_ied(x0[i], ridx, lidx) for i in range(n)])
    nrm = (1.0 / ied).astype(np.float32)
    b = ((x0 - x_gt) * nrm[:, None]).astype(np.float32)
    t0 = time.time()
    X_ref, lam_ref = _truth_dual(A_ref, b, 1.5)
    t_truth = time.time() - t0
    upd = (A_ref.astype(np.float64) @ X_ref).astype(np.float32)
    nxt_ref = (x0 - upd * (np.float32(1.0) / nrm)[:, None]).astype(np.float32)

    ctx = sd.default_context()
    ht = sd.HogTransform(images, [hp], ids, right, left)
    A_gpu = ht(x0, 0).cpu().numpy()
    e_a = rel_err(A_gpu, A_ref)
    sdo = sd.SupervisedDescentOptimiser([sd.LinearRegressor(sd.Regulariser(sd.RegularisationType.MatrixNorm, 1.5, False))],
                                        sd.InterEyeDistanceNormalisation(ids, right, left))
    got = sdo.train(x_gt, x0, None, ht).cpu().numpy()
    X = sdo.regressors[0].x.cpu().numpy()
    lam = sdo.regressors[0].last_lambda
    e_w, e_x = rel_err(X, X_ref), rel_err(got, nxt_ref)
    e_p = rel_err(A_ref.astype(np.float64) @ X.astype(np.float64), A_ref.astype(np.float64) @ X_ref)
    print(f"{name}: N={n} D={D}; oracle HOG {t_hog:.1f} s, float64 truth {t_truth:.1f} s; features {e_a:.2e}; lambda {lam:.6g} vs {lam_ref:.6g}; "
          f"weights {e_w:.2e}; predictions {e_p:.2e}; updated landmarks {e_x:.2e}; solver ms {ctx.solver_timings()}")
    if name == "config4":
        import scipy.linalg
        t0 = time.time()
        G = (A_ref.T @ A_ref).astype(np.float32)
        G[np.diag_indices_from(G)] += np.float32(lam_ref)
        G[-1, -1] -= np.float32(lam_ref)
        X_lu = scipy.linalg.solve(G, (A_ref.T @ b).astype(np.float32), check_finite=False)
        print(f"{name}: float32 LAPACK partial-pivot LU (the reference's algorithm) vs float64: weights {rel_err(X_lu, X_ref):.2e}, "
              f"predictions {rel_err(A_ref.astype(np.float64) @ X_lu.astype(np.float64), A_ref.astype(np.float64) @ X_ref):.2e} ({time.time() - t0:.1f} s); "
              f"ours vs that float32 solve: weights {rel_err(X, X_lu):.2e}")
    assert e_a <= 1e-4
    assert abs(lam - lam_ref) <= 2e-5 * lam_ref
    assert e_w <= 1e-4
    assert e_p <= 1e-4
    assert e_x <= 1e-4
    del ht, sdo
    torch.cuda.empty_cache()
