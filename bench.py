#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 cascaded-regression engine.

Metric (BASELINE.json): faces/sec, RCR 22-landmark detect with the reference's pre-trained
face_landmarks_model_rcr_22.bin on 640x480 synthetic 8UC1 frames, batched, one face box per frame
(config 3, "configs[2]").  A "step" = one pass of the detect cascade (4 levels: HOG -> feature x weight
GEMM -> IED-scaled update) over one batch of B frames.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

  value        whole-job faces/s with frames + initial landmarks already resident in HBM
  e2e          the same through the reference-facing call detection_model::detect(image, facebox)
               batched with HOST (pinned) buffers: H2D of the frames and D2H of the landmarks are inside
               the timed region
  roofline     the dominant kernel (HOG, cascade level 0) against measured HBM bandwidth
  cpu_baseline the reference's own hog.c (oracle/_ref) inside the restated HogTransform/predict glue,
               timed on this box's host cores on a bounded sample
  train        (N=1 only, extra) regressor-train seconds of a reduced RCR training config

--impl reference times the CPU path alone (rank 0), same metric/config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
MODEL = os.path.join(ROOT, "tests", "golden", "face_landmarks_model_rcr_22.bin")
W_IMG, H_IMG = 640, 480


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def synth_boxes(count, seed):
    """one square face box per frame, w=h in U{200..280}, fully inside the 640x480 frame."""
    rng = np.random.Generator(np.random.PCG64(seed))
    s = rng.integers(200, 281, size=count)
    x = (rng.random(count) * (W_IMG - s - 40) + 20).astype(np.int64)
    y = (rng.random(count) * (H_IMG - s - 40) + 20).astype(np.int64)
    return np.stack([x, y, s, s], axis=1).astype(np.int32)


def synth_frames_torch(count, seed, device):
    """Low-pass filtered uniform noise (sigma = 3 px) stretched to 0..255, 8UC1, generated on the GPU."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    sigma, r = 3.0, 9
    k = torch.exp(-0.5 * (torch.arange(-r, r + 1, device=device, dtype=torch.float32) / sigma) ** 2)
    k = k / k.sum()
    out = torch.empty((count, H_IMG, W_IMG), dtype=torch.uint8, device=device)
    for i0 in range(0, count, 128):
        n = min(128, count - i0)
        x = torch.rand((n, 1, H_IMG + 2 * r, W_IMG + 2 * r), generator=g, device=device)
        x = F.conv2d(x, k.view(1, 1, -1, 1))
        x = F.conv2d(x, k.view(1, 1, 1, -1))
        lo = x.amin(dim=(2, 3), keepdim=True)
        hi = x.amax(dim=(2, 3), keepdim=True)
        out[i0:i0 + n] = ((x - lo) / (hi - lo) * 255.0).round().clamp(0, 255).to(torch.uint8)[:, 0]
    return out


def synth_frames_numpy(count, seed):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import synth
    return synth.smooth_images(count, H_IMG, W_IMG, seed)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_detect_rate(n_faces, threads, seed):
    """Reference CPU path: oracle glue + the reference's hog.c when oracle/_ref is built."""
    from oracle import oracle as O
    O.build()
    om = O.Model(MODEL)
    use_ref = O.ref_available()
    frames = synth_frames_numpy(min(n_faces, 64), seed)
    reps = (n_faces + frames.shape[0] - 1) // frames.shape[0]
    frames = np.concatenate([frames] * reps)[:n_faces]
    boxes = synth_boxes(n_faces, seed)
    om.detect_batch(frames[:threads], boxes[:threads], use_ref=use_ref, threads=threads)   # warm-up
    t0 = time.perf_counter()
    om.detect_batch(frames, boxes, use_ref=use_ref, threads=threads)
    dt = time.perf_counter() - t0
    return n_faces / dt, ("reference" if use_ref else "port"), dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    sample = max(cores * 8, 256)
    rates = []
    kind = "port"
    for s in range(args.warmup + args.steps):
        r, kind, dt = cpu_detect_rate(sample, cores, 1234 + s)
        if s >= args.warmup:
            rates.append((sample, dt))
    faces = sum(a for a, _ in rates)
    secs = sum(b for _, b in rates)
    value = faces / secs
    r1, _, _ = cpu_detect_rate(64, 1, 99)
    line = {
        "impl": "reference", "metric": "faces/sec RCR 22-landmark detect", "value": value, "unit": "faces/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(len(rates), 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[2]: RCR 22-landmark detect, face_landmarks_model_rcr_22.bin, 640x480 8UC1 synthetic frames, "
                               f"{sample} faces per step on host cores", "frames_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "faces/s", "cores": cores, "kind": kind,
                         "sample": f"{sample} faces/step x {args.steps} steps, one face per thread; single-thread (reference-faithful sequential predict): {r1:.1f} faces/s",
                         "single_thread_value": r1},
        "e2e": {"value": value, "unit": "faces/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# measured once with ncu (--set full) on hog_patch_kernel<4,5,11>, 2048 faces x 22 patches: 160.26 MB read + 51.71 MB written
HOG_L0_DRAM_BYTES_PER_FACE = (160.259584e6 + 51.714560e6) / 2048

TRAIN_CFG = {"n": 10000, "size": 128, "num_bins": 9, "cells": 5, "cell_sizes": [11, 10, 8, 6, 6], "rel": [1.0, 0.7, 0.4, 0.25, 0.25],
             "lambda_factor": 1.5}


def synth_train_set(sd, model, n_local, seed, dev):
    """128x128 8UC1 crops, box = crop shrunk by 10 %, ground truth = mean shape in a box
    jittered N(0, 0.04) in translation and N(1, 0.04) in scale (rcr-train.cpp:387-395), x0 = mean in the box."""
    import torch
    import torch.nn.functional as F
    size = TRAIN_CFG["size"]
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    sigma, r = 3.0, 9
    k = torch.exp(-0.5 * (torch.arange(-r, r + 1, device=dev, dtype=torch.float32) / sigma) ** 2)
    k = k / k.sum()
    imgs = torch.empty((n_local, size, size), dtype=torch.uint8, device=dev)
    for i0 in range(0, n_local, 1024):
        n = min(1024, n_local - i0)
        x = torch.rand((n, 1, size + 2 * r, size + 2 * r), generator=g, device=dev)
        x = F.conv2d(F.conv2d(x, k.view(1, 1, -1, 1)), k.view(1, 1, 1, -1))
        lo, hi = x.amin(dim=(2, 3), keepdim=True), x.amax(dim=(2, 3), keepdim=True)
        imgs[i0:i0 + n] = ((x - lo) / (hi - lo) * 255.0).round().clamp(0, 255).to(torch.uint8)[:, 0]
    rng = np.random.Generator(np.random.PCG64(seed))
    mean = model.get_mean()
    m = int(round(size * 0.05))
    box = (m, m, size - 2 * m, size - 2 * m)
    x0 = np.tile(sd.align_mean(mean, box), (n_local, 1)).astype(np.float32)
    x_gt = np.stack([sd.align_mean(mean, box, 1.0 + rng.normal(0, 0.04), 1.0 + rng.normal(0, 0.04), rng.normal(0, 0.04), rng.normal(0, 0.04))
                     for _ in range(n_local)]).astype(np.float32)
    return imgs, x0, x_gt


def run_train(sd, ctx, model, world, rank, dev, barrier, max_over_ranks, group):
    """Regressor-train seconds (all S levels: HOG + targets + Gram + all-reduce + solve + update), strong scaling."""
    import torch
    from superviseddescent_b200 import parallel
    cfg = TRAIN_CFG
    b, e = parallel.shard_range(cfg["n"], world, rank)
    imgs, x0, x_gt = synth_train_set(sd, model, e - b, 2024 + rank, dev)
    ids = model.landmark_ids
    norm_c = sd.NormalisationC()
    import ctypes as C
    from superviseddescent_b200 import _capi
    _capi.lib().sd_model_normalisation(model._m, C.byref(norm_c))
    right = [ids[norm_c.right_idx[i]] for i in range(norm_c.n_right)]
    left = [ids[norm_c.left_idx[i]] for i in range(norm_c.n_left)]
    hps = [sd.HoGParam(1, cfg["cells"], cs, cfg["num_bins"], rel) for cs, rel in zip(cfg["cell_sizes"], cfg["rel"])]
    ht = sd.HogTransform(imgs, hps, ids, right, left, ctx)
    D = ht.feature_length(0)

    def one_run():
        regs = [sd.LinearRegressor(sd.Regulariser(sd.RegularisationType.MatrixNorm, cfg["lambda_factor"], False), ctx) for _ in hps]
        sdo = sd.SupervisedDescentOptimiser(regs, sd.InterEyeDistanceNormalisation(ids, right, left), ctx)
        xf = sdo.train(x_gt, x0, None, ht, None, group)
        return sdo, xf

    one_run()                                  # warm-up (workspaces, tensor maps, NCCL channels)
    barrier()
    l0 = ctx.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sdo, xf = one_run()
    e1.record()
    barrier()
    secs = max_over_ranks(e0.elapsed_time(e1)) * 1e-3
    res0 = float(torch.linalg.norm(torch.from_numpy(x0).to(dev) - torch.from_numpy(x_gt).to(dev)) / torch.linalg.norm(torch.from_numpy(x_gt).to(dev)))
    res1 = float(torch.linalg.norm(xf - torch.from_numpy(x_gt).to(dev)) / torch.linalg.norm(torch.from_numpy(x_gt).to(dev)))
    S = len(hps)
    gram_flops = S * (cfg["n"] * D * (D + 1) + 2.0 * cfg["n"] * D * 44)
    chol_flops = S * (D ** 3 / 3.0 + 2.0 * D * D * 44)
    return {"metric": "regressor train sec (RCR, all cascade levels)", "value": secs, "unit": "s", "higher_is_better": False, "scaling": "strong",
            "config": {"workload": "configs[3]: RCR training, 10k synthetic 128x128 crops, 22 landmarks, 31-bin HOG (K=9), 5 cascade levels",
                       "samples_global": cfg["n"], "samples_this_rank": e - b, "feature_dim": D, "levels": S,
                       "parallelism": f"samples sharded over {world} GPU(s), one all-reduce of the upper row bands of [AtA|Atb] "
                                      f"({sum((min(b + 1024, D) - b) * (D + 44 - b) for b in range(0, D, 1024)) * 4 / 1e9:.2f} of {D * (D + 44) * 4 / 1e9:.2f} GB) per level, replicated solve"},
            "gpu_launches": int(ctx.launches() - l0),
            "algorithmic_tflop": {"gram_syrk": gram_flops / 1e12, "cholesky_and_solve": chol_flops / 1e12},
            "train_residual": {"before": res0, "after": res1},
            "last_level_solver_ms": ctx.solver_timings()}


def cpu_train_level_seconds(n_samples, threads, seed=2024):
    """Reference CPU path for ONE training level of config 4 (level 0: the most expensive one), all host threads:
    the reference's hog.c inside the restated HogTransform glue (one sample per thread, as the thread pool of
    superviseddescent.hpp:173-189), then BLAS/LAPACK (numpy/scipy sgemm, sgetrf, sgetrs) standing in for Eigen's
    A^T A and PartialPivLU (regressors.hpp:199-234) -- BASELINE.md section 3."""
    import scipy.linalg
    from oracle import oracle as O
    O.build()
    om = O.Model(MODEL)
    use_ref = O.ref_available()
    cfg = TRAIN_CFG
    size = cfg["size"]
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import synth
    base = synth.smooth_images(64, size, size, seed)
    imgs = np.concatenate([base] * ((n_samples + 63) // 64))[:n_samples]
    m = int(round(size * 0.05))
    box = (m, m, size - 2 * m, size - 2 * m)
    rng = np.random.Generator(np.random.PCG64(seed))
    x0 = np.tile(O.align_mean(om.mean, box), (n_samples, 1)).astype(np.float32)
    x_gt = np.stack([O.align_mean(om.mean, box, 1.0 + rng.normal(0, 0.04), 1.0 + rng.normal(0, 0.04), rng.normal(0, 0.04), rng.normal(0, 0.04))
                     for _ in range(n_samples)]).astype(np.float32)
    hp = O.HogParam(1, cfg["cells"], cfg["cell_sizes"][0], cfg["num_bins"], cfg["rel"][0])
    t0 = time.perf_counter()
    A = O.hog_transform_batch(imgs, x0, hp, om.right_idx, om.left_idx, use_ref=use_ref, threads=threads)
    t_hog = time.perf_counter() - t0
    ied = np.array([O.get_ied(x0[i], om.right_idx, om.left_idx) for i in range(n_samples)])
    b = ((x0 - x_gt) / ied[:, None]).astype(np.float32)
    t0 = time.perf_counter()
    G = A.T @ A
    lam = 1.5 * np.linalg.norm(G) / n_samples
    G[np.diag_indices_from(G)] += lam
    G[-1, -1] -= lam
    t_gram = time.perf_counter() - t0
    t0 = time.perf_counter()
    lu = scipy.linalg.lu_factor(G, overwrite_a=True, check_finite=False)
    X = scipy.linalg.lu_solve(lu, A.T @ b, check_finite=False)
    t_lu = time.perf_counter() - t0
    t0 = time.perf_counter()
    _ = x0 - (A @ X) * ied[:, None]
    t_upd = time.perf_counter() - t0
    return {"hog_s": t_hog, "gram_s": t_gram, "lu_solve_s": t_lu, "update_s": t_upd, "total_s": t_hog + t_gram + t_lu + t_upd,
            "kind": "reference hog.c + BLAS/LAPACK for Eigen" if use_ref else "port + BLAS/LAPACK"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    from superviseddescent_b200 import api as sd
    ctx = sd.Context(local)
    model = sd.load_detection_model(MODEL, ctx)
    B = args.batch
    L = model.num_landmarks

    frames = synth_frames_torch(B, 1234 + rank, dev)
    boxes = synth_boxes(B, 1234 + rank)
    mean = model.get_mean()
    x0 = np.stack([sd.align_mean(mean, b) for b in boxes])
    x0_dev = torch.from_numpy(x0).to(dev)
    h_frames = torch.empty((B, H_IMG, W_IMG), dtype=torch.uint8).pin_memory()
    h_frames.copy_(frames)
    torch.cuda.synchronize()
    h_np = h_frames.numpy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world > 1:
            t = torch.tensor([v], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return v

    # ---------------- device-resident throughput ----------------
    for _ in range(args.warmup):
        out = model.detect_batch_device(frames, x0_dev)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = model.detect_batch_device(frames, x0_dev)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.launches() - l0
    value = world * B * args.steps / (ms * 1e-3)

    # ---------------- end to end through the host-buffer call ----------------
    for _ in range(min(args.warmup, 2)):
        lm = model.detect_batch(h_np, boxes)
    barrier()
    e0.record()
    for _ in range(args.steps):
        lm = model.detect_batch(h_np, boxes)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None      # sampled across both timed regions (device-resident and e2e)
    e2e = world * B * args.steps / (ms_e2e * 1e-3)
    assert np.array_equal(lm, out.cpu().numpy()), "host and device paths disagree"

    # ---------------- roofline of the dominant kernel: HOG, cascade level 0 ----------------
    import ctypes as C
    from superviseddescent_b200 import _capi
    hp0 = model.hog_param(0)
    D0 = _capi.lib().sd_hog_feature_length(L, C.byref(hp0))
    ld = (D0 + 3) // 4 * 4
    A = torch.empty((B, ld), dtype=torch.float32, device=dev)
    norm = sd.NormalisationC()
    _capi.lib().sd_model_normalisation(model._m, C.byref(norm))
    ib = sd.ImageBatchC(C.c_void_p(frames.data_ptr()), W_IMG, H_IMG, frames.stride(1), frames.stride(0), B)

    def hog0():
        rc = _capi.lib().sd_hog_batch(ctx.h, C.byref(ib), None, _capi.ptr(x0_dev), C.c_int64(2 * L), B, L, C.byref(norm), C.byref(hp0), _capi.ptr(A), C.c_int64(ld))
        assert rc == 0
    for _ in range(3):
        hog0()
    torch.cuda.synchronize()
    reps = max(5, args.steps)
    e0.record()
    for _ in range(reps):
        hog0()
    e1.record()
    torch.cuda.synchronize()
    hog_ms = e0.elapsed_time(e1) / reps
    # algorithmic bytes: unique source pixels read once + the descriptor row written once
    ri = [norm.right_idx[i] for i in range(norm.n_right)]
    li = [norm.left_idx[i] for i in range(norm.n_left)]
    ied = np.hypot(x0[:, ri].mean(1) - x0[:, li].mean(1), x0[:, [i + L for i in ri]].mean(1) - x0[:, [i + L for i in li]].mean(1))
    P = 2 * np.round(hp0.relative_patch_size * ied / 2)
    alg_bytes = float(np.sum(np.minimum(L * P * P, W_IMG * H_IMG)) + B * D0 * 4)
    fs0 = hp0.num_cells * hp0.cell_size
    alg_flops = float(B * L * fs0 * fs0 * (24 + 4 * hp0.num_bins))
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (hog_ms * 1e-3) / 1e9
    roofline = {"kernel": f"hog_patch_kernel<{hp0.num_bins}> (cascade level 0, fs={fs0})", "bound": "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": HOG_L0_DRAM_BYTES_PER_FACE * B, "peak_source": peak_src, "ms_per_launch": hog_ms,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel at 2048 faces "
                                  "(profiles/r01_summary.md section 12), scaled linearly to this batch; below the algorithmic bytes because "
                                  "the 22 patches of a face overlap (SURVEY 8d counts L*P^2 source pixels)",
                "issue_slots_busy_pct": 76.4,
                "algorithmic_bytes_per_launch": alg_bytes,
                "note": "HOG is fp32-ALU/shared-memory bound (SURVEY 8d: ~40 flop/B); fp32 figure reported beside the HBM one",
                "achieved_fp32_tflops": alg_flops / (hog_ms * 1e-3) / 1e12}

    train = None
    if not args.no_train:
        try:
            train = run_train(sd, ctx, model, world, rank, dev, barrier, max_over_ranks, group)
        except Exception as ex:   # the headline line must still be printed
            train = {"error": repr(ex)[:300]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {
        "metric": "faces/sec RCR 22-landmark detect", "value": value, "unit": "faces/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[2]: RCR 22-landmark detect, pre-trained face_landmarks_model_rcr_22.bin, 640x480 8UC1 synthetic frames, batched",
                   "frames_per_gpu": B, "global_batch": world * B, "cascade_levels": model.num_levels, "landmarks": L,
                   "parallelism": f"face-batch sharded over {world} GPU(s), no collective",
                   "l2": f"inputs ({B * W_IMG * H_IMG / 1e6:.0f} MB of frames per GPU) exceed the 126 MB L2"},
        "e2e": {"value": e2e, "unit": "faces/s", "h2d_bytes_per_step": int(B * W_IMG * H_IMG + B * 2 * L * 4), "d2h_bytes_per_step": int(B * 2 * L * 4),
                "ms_per_step": ms_e2e / args.steps, "api": "detection_model.detect_batch (sd_detect_batch_host), pinned host frames",
                "note": "h2d_bytes_per_step counts the host frames handed to the call; the engine's region-of-interest route reads only "
                        "each face's window (~1/4 of a frame) over PCIe inside the timed region (DESIGN.md 4.5)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
    }
    if train is not None:
        line["train"] = train
    if world == 1 and not args.no_cpu and train is not None and "value" in train:
        try:
            cores = host_cores()
            lvl = cpu_train_level_seconds(TRAIN_CFG["n"], cores)
            train["cpu_baseline"] = {"value": lvl["total_s"] * len(TRAIN_CFG["cell_sizes"]), "unit": "s", "cores": cores, "kind": "port",
                                     "sample": f"ONE full level (level 0, all {TRAIN_CFG['n']} samples) timed on the host: HOG {lvl['hog_s']:.2f} s, "
                                               f"Gram {lvl['gram_s']:.2f} s, LU+solve {lvl['lu_solve_s']:.2f} s, update {lvl['update_s']:.2f} s; "
                                               f"x{len(TRAIN_CFG['cell_sizes'])} levels (extrapolated); {lvl['kind']}"}
        except Exception as ex:
            train["cpu_baseline"] = {"error": repr(ex)[:200]}
    if world == 1 and not args.no_cpu:
        cores = host_cores()
        n = max(256, cores * 32)
        r, kind, dt = cpu_detect_rate(n, cores, 4321)
        r1, _, _ = cpu_detect_rate(48, 1, 4322)
        line["cpu_baseline"] = {"value": r, "unit": "faces/s", "cores": cores, "kind": kind,
                                "sample": f"{n} faces of the same workload, one face per thread ({dt:.1f} s); single-thread reference-faithful predict: {r1:.1f} faces/s",
                                "single_thread_value": r1}
    if world > 1:
        dist.destroy_process_group()
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4096, help="frames per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-train", action="store_true", help="skip the extra regressor-train measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
