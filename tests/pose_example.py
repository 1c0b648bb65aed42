"""Config 2: the 6-DOF pose estimation example of the reference (examples/pose_estimation.cpp:187-337), restated
as a host projection functor (user code, as in the reference) and shared by the oracle and the GPU tests.
The reference seeds std::mt19937 from std::random_device; SURVEY 8d fixes the seed (numpy PCG64(42) here)."""
import numpy as np

FACEMODEL = np.array([   # examples/pose_estimation.cpp:256-266, iBug points 31,34,37,40,43,46,49,52,55,58 (x, y, z, 1)
    [-0.287526, -2.0203, 3.33725, 1.0], [-0.11479, -17.2056, -13.5569, 1.0], [-46.1668, 34.7219, -35.938, 1.0],
    [-18.926, 31.5432, -29.9641, 1.0], [19.2574, 31.5767, -30.229, 1.0], [46.1914, 34.452, -36.1317, 1.0],
    [-23.7552, -35.7461, -28.2573, 1.0], [-0.0753515, -28.3064, -12.8984, 1.0], [23.7138, -35.7886, -28.5949, 1.0],
    [0.125511, -44.7427, -17.1411, 1.0]], dtype=np.float32).T       # 4 x 10


def _rot(axis, deg):
    a = np.float32(np.deg2rad(np.float32(deg)))
    c, s = np.float32(np.cos(a)), np.float32(np.sin(a))
    m = np.eye(4, dtype=np.float32)
    if axis == "x":
        m[1, 1], m[1, 2], m[2, 1], m[2, 2] = c, -s, s, c
    elif axis == "y":
        m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c
    else:
        m[0, 0], m[0, 1], m[1, 0], m[1, 1] = c, -s, s, c
    return m


def projection(params, level=0, idx=0):
    """ModelProjection::operator() (:187-240): 3-D model -> 2*10 normalised 2-D landmarks."""
    p = np.asarray(params, dtype=np.float32).ravel()
    focal = np.float32(1800.0)
    t = np.eye(4, dtype=np.float32)
    t[:3, 3] = p[3:6]
    model = t @ _rot("y", p[1]) @ _rot("x", p[0]) @ _rot("z", p[2])
    fovy = np.float32(2.0) * np.arctan(np.float32(1000.0) / (np.float32(2.0) * focal)) * np.float32(180.0 / np.pi)   # focalLengthToFovy
    rad = (fovy / np.float32(2.0)) * np.float32(np.pi) / np.float32(180.0)
    cot = np.float32(np.cos(rad) / np.sin(rad))
    n, f = np.float32(1.0), np.float32(5000.0)
    persp = np.array([[cot, 0, 0, 0], [0, cot, 0, 0], [0, 0, -(n + f) / (f - n), (-2 * n * f) / (f - n)], [0, 0, -1, 0]], dtype=np.float32)
    clip = (persp @ model @ FACEMODEL).astype(np.float32)
    clip = clip / clip[3]
    x_ss = (clip[0] + 1.0) * 500.0
    y_ss = 1000.0 - (clip[1] + 1.0) * 500.0
    out = np.concatenate([(x_ss - 500.0) / focal, (y_ss - 500.0) / focal]).astype(np.float32)
    return out


def training_set(num_samples=500, seed=42):
    rng = np.random.Generator(np.random.PCG64(seed))
    x_tr = np.zeros((num_samples, 6), dtype=np.float32)
    x_tr[:, :3] = rng.uniform(-30, 30, size=(num_samples, 3)).astype(np.float32)
    x_tr[:, 5] = -2000.0
    y_tr = np.stack([projection(x_tr[i]) for i in range(num_samples)])
    x0 = np.zeros_like(x_tr)
    x0[:, 5] = -2000.0
    return x_tr, y_tr, x0


# the hand-labelled frame of the example (:323-333): ground truth pitch 11, yaw -25, roll -10
TEST_LANDMARKS = ((np.array([498, 504, 479, 498, 529, 553, 489, 503, 527, 503, 502, 513, 457, 465, 471, 471, 522, 522, 530, 536], dtype=np.float32) - 500.0) / 1800.0).reshape(1, 20)
TEST_INIT = np.array([[0, 0, 0, 0, 0, -2000]], dtype=np.float32)
