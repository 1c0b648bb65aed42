"""One small HOG launch through the C ABI (development helper): python tools/diag_hog.py [n_faces]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import synth
from superviseddescent_b200 import api as sd
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
m = sd.load_detection_model(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "face_landmarks_model_rcr_22.bin"))
images = synth.smooth_images(n, 240, 320, seed=1)
boxes = synth.face_boxes(n, 240, 320, seed=1)
x0 = np.stack([sd.align_mean(m.get_mean(), b) for b in boxes]).astype(np.float32)
ids = m.landmark_ids
ht = sd.HogTransform(images, [m.hog_param(0)], ids, ["37", "40"], ["43", "46"])
A = ht(x0, 0)
torch.cuda.synchronize()
print("hog ok", float(A.abs().sum()), flush=True)
out = m.detect_batch(images, boxes)
print("detect ok", float(np.abs(out).sum()), flush=True)
