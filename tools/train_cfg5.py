"""Config 5 of BASELINE.json on ONE GPU: 100k synthetic 256x256 crops, 68 landmarks, K=9, 6 cascade levels
(D = 68*25*31 + 1 = 52,701).  One warm-up level, then the timed 6-level run.  Development helper: the numbers it
prints are quoted in DESIGN.md / profiles; bench.py keeps config 4 as its `train` workload.
   N=100000 SIZE=256 python tools/train_cfg5.py        (N / SIZE / LEVELS can be reduced for a quick check)"""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from superviseddescent_b200 import api as sd

N = int(os.environ.get("N", "100000")); SIZE = int(os.environ.get("SIZE", "256")); LEVELS = int(os.environ.get("LEVELS", "6"))
cell_sizes = [11, 10, 8, 6, 6, 6][:LEVELS]; rel = [1.0, 0.7, 0.4, 0.25, 0.25, 0.25][:LEVELS]
dev = torch.device("cuda", 0)
ctx = sd.Context(0)
mean = np.load(os.path.join(ROOT, "tests", "golden", "mean_ibug_lfpw_68.npy")).astype(np.float32).reshape(-1)
ids = [str(i) for i in range(1, 69)]
right, left = ["37", "40"], ["43", "46"]

class Mean68:                      # the two members synth_train_set needs from a detection model
    def get_mean(self): return mean
bench.TRAIN_CFG.update({"n": N, "size": SIZE})
t0 = time.time()
imgs, x0, x_gt = bench.synth_train_set(sd, Mean68(), N, 2025, dev)
print(f"synthetic set: {N} crops {SIZE}x{SIZE}, {time.time() - t0:.1f} s", flush=True)
hps = [sd.HoGParam(1, 5, cs, 9, r) for cs, r in zip(cell_sizes, rel)]
ht = sd.HogTransform(imgs, hps, ids, right, left, ctx)
D = ht.feature_length(0)

def run(levels):
    regs = [sd.LinearRegressor(sd.Regulariser(sd.RegularisationType.MatrixNorm, 1.5, False), ctx) for _ in range(levels)]
    sdo = sd.SupervisedDescentOptimiser(regs, sd.InterEyeDistanceNormalisation(ids, right, left), ctx)
    per_level = []
    last = [None]
    def cb(cur):
        torch.cuda.synchronize()
        now = time.time()
        per_level.append((now - last[0], ctx.solver_timings()))
        last[0] = now
    torch.cuda.synchronize(); last[0] = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    xf = sdo.train(x_gt, x0, None, ht, cb)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3, per_level, xf

secs, lv, _ = run(1)
print(f"warm-up level: {secs:.3f} s  {lv}", flush=True)
secs, lv, xf = run(LEVELS)
g = torch.from_numpy(x_gt).to(dev)
res0 = float(torch.linalg.norm(torch.from_numpy(x0).to(dev) - g) / torch.linalg.norm(g)); res1 = float(torch.linalg.norm(xf - g) / torch.linalg.norm(g))
out = {"config": f"configs[4]: {N} crops {SIZE}x{SIZE}, 68 landmarks, K=9, {LEVELS} levels, D={D}, one GPU", "train_seconds": secs,
       "per_level_wall_s": [round(t, 3) for t, _ in lv], "per_level_solver_ms": [s for _, s in lv],
       "gram_tflop_per_level": N * D * (D + 1 + 2 * 136) / 1e12, "cholesky_tflop_per_level": D ** 3 / 3e12,
       "residual": {"before": res0, "after": res1}, "max_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
print(json.dumps(out), flush=True)
