"""Times the levels of a training config once (LEVELS=n selects how many, WORKLOAD=train|train5); the command the ncu launch
lists and captures in profiles/ were taken with.  Development helper, not part of the product."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from superviseddescent_b200 import api as sd
ctx = sd.Context(0)
model = sd.load_detection_model(bench.MODEL, ctx)
cfg = dict(bench.TRAIN_CFGS[os.environ.get("WORKLOAD", "train")])
lv = int(os.environ.get("LEVELS", "1"))
cfg["cell_sizes"], cfg["rel"] = cfg["cell_sizes"][:lv], cfg["rel"][:lv]
if "N" in os.environ:
    cfg["n"] = int(os.environ["N"])
r = bench.run_train(sd, ctx, model, 1, 0, torch.device("cuda", 0), torch.cuda.synchronize, lambda v: v, None, cfg, solver=os.environ.get("SOLVER", "cholesky"))
print(r["value"], r["last_level_solver_ms"], r["gpu_launches"], r["train_residual"], r.get("roofline"))
