"""Times the levels of the config-4 training run once (LEVELS=n selects how many); the command the ncu launch lists and
captures in profiles/ were taken with.  Development helper, not part of the product."""
import sys, os, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
from superviseddescent_b200 import api as sd
ctx = sd.Context(0)
model = sd.load_detection_model(bench.MODEL, ctx)
bench.TRAIN_CFG["cell_sizes"] = bench.TRAIN_CFG["cell_sizes"][:int(os.environ.get("LEVELS","1"))]
bench.TRAIN_CFG["rel"] = bench.TRAIN_CFG["rel"][:int(os.environ.get("LEVELS","1"))]
r = bench.run_train(sd, ctx, model, 1, 0, torch.device("cuda",0), torch.cuda.synchronize, lambda v: v, None)
print(r["value"], r["last_level_solver_ms"], r["gpu_launches"])
