"""ctypes binding of the CPU oracle (oracle/sd_oracle.c) and of oracle/_ref.

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never from the product
package (superviseddescent_b200/).  See oracle/sd_oracle.h for the reference
file:line each function restates.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libsd_oracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libref_vlhog.so")


def build(quiet: bool = True) -> None:
    """Compile the C restatement and (if /root/reference exists) oracle/_ref."""
    subprocess.run(["make", "-C", _HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


class HogParam(C.Structure):
    _fields_ = [("variant", C.c_int32), ("num_cells", C.c_int32), ("cell_size", C.c_int32),
                ("num_bins", C.c_int32), ("relative_patch_size", C.c_float)]


class Regulariser(C.Structure):
    _fields_ = [("type", C.c_int32), ("lambda_", C.c_float), ("regularise_last_row", C.c_int32)]


class Normalisation(C.Structure):
    _fields_ = [("kind", C.c_int32), ("right_idx", C.POINTER(C.c_int32)), ("n_right", C.c_int32),
                ("left_idx", C.POINTER(C.c_int32)), ("n_left", C.c_int32)]


class _Model(C.Structure):
    _fields_ = [("num_levels", C.c_int32), ("num_landmarks", C.c_int32),
                ("rows", C.POINTER(C.c_int32)), ("cols", C.POINTER(C.c_int32)),
                ("weights", C.POINTER(C.POINTER(C.c_float))),
                ("regularisers", C.POINTER(Regulariser)),
                ("mean", C.POINTER(C.c_float)),
                ("landmark_ids", C.POINTER(C.c_char_p)),
                ("hog_params", C.POINTER(HogParam)),
                ("n_right", C.c_int32), ("n_left", C.c_int32),
                ("right_idx", C.POINTER(C.c_int32)), ("left_idx", C.POINTER(C.c_int32)),
                ("right_ids", C.POINTER(C.c_char_p)), ("left_ids", C.POINTER(C.c_char_p))]


HOG_CORE_FN = C.CFUNCTYPE(None, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                          C.POINTER(C.c_float))
PROJECTION_FN = C.CFUNCTYPE(None, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int,
                            C.POINTER(C.c_float), C.c_void_p)
EPOCH_CB = C.CFUNCTYPE(None, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_void_p)

_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_get_ied.restype = C.c_double
        _lib.orc_test_residual.restype = C.c_double
        _lib.orc_regulariser_lambda.restype = C.c_float
        _lib.orc_model_load.restype = C.POINTER(_Model)
    return _lib


def ref_available() -> bool:
    return os.path.exists(_REF_PATH)


def ref():
    """The reference's own hog.c (oracle/_ref), or None if it was never built."""
    global _ref
    if _ref is None and ref_available():
        _ref = C.CDLL(_REF_PATH)
    return _ref


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _core_ptr(use_ref: bool):
    """Function pointer of the HOG core: the restatement or the reference's hog.c."""
    if use_ref:
        r = ref()
        if r is None:
            raise RuntimeError("oracle/_ref/libref_vlhog.so is not built")
        return C.cast(r.ref_hog_core, C.c_void_p)
    return None


def hog_dimension(variant: int, num_bins: int) -> int:
    return lib().orc_hog_dimension(variant, num_bins)


def hog_core(image: np.ndarray, cell_size: int, num_bins: int, variant: int = 1, use_ref: bool = False) -> np.ndarray:
    """HOG of a float32 (h, w) image; returns the planar [dd, ch, cw] array."""
    image = np.ascontiguousarray(image, dtype=np.float32)
    h, w = image.shape
    cw, ch = (w + cell_size // 2) // cell_size, (h + cell_size // 2) // cell_size
    dd = hog_dimension(variant, num_bins)
    out = np.zeros((dd, ch, cw), dtype=np.float32)
    if use_ref:
        ref().ref_hog_core(_fp(image), w, h, cell_size, num_bins, variant, _fp(out))
    else:
        lib().orc_hog_core(_fp(image), w, h, cell_size, num_bins, variant, _fp(out))
    return out


def hog_orientation_bins(image: np.ndarray, num_bins: int) -> np.ndarray:
    image = np.ascontiguousarray(image, dtype=np.float32)
    h, w = image.shape
    out = np.zeros((h, w), dtype=np.int32)
    lib().orc_hog_orientation_bins(_fp(image), w, h, num_bins, _ip(out))
    return out


def resize_linear_u8(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    src = np.ascontiguousarray(src, dtype=np.uint8)
    sh, sw = src.shape
    dst = np.zeros((dh, dw), dtype=np.uint8)
    lib().orc_resize_linear_u8(_u8p(src), sw, sh, sw, _u8p(dst), dw, dh, dw)
    return dst


def bgr2gray_u8(bgr: np.ndarray) -> np.ndarray:
    bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
    h, w, _ = bgr.shape
    g = np.zeros((h, w), dtype=np.uint8)
    lib().orc_bgr2gray_u8(_u8p(bgr), w, h, 3 * w, _u8p(g), w)
    return g


def perturb_box(box, tx: float, ty: float, scaling: float = 1.0):
    """apps/rcr/rcr-train.cpp:130-146."""
    b = (C.c_int32 * 4)(*[int(v) for v in box])
    o = (C.c_int32 * 4)()
    lib().orc_perturb_box(b, C.c_float(tx), C.c_float(ty), C.c_float(scaling), o)
    return tuple(int(v) for v in o)


def normalised_landmark_errors(pred: np.ndarray, gt: np.ndarray, right_idx, left_idx) -> np.ndarray:
    """apps/rcr/rcr-train.cpp:200-212: (N, L) IED-normalised per-landmark errors."""
    pred = np.ascontiguousarray(pred, dtype=np.float32)
    gt = np.ascontiguousarray(gt, dtype=np.float32)
    n, p2 = pred.shape
    r = np.asarray(right_idx, dtype=np.int32)
    l = np.asarray(left_idx, dtype=np.int32)
    out = np.zeros((n, p2 // 2), dtype=np.float32)
    lib().orc_normalised_landmark_errors(_fp(pred), _fp(gt), n, p2 // 2, _ip(r), r.size, _ip(l), l.size, _fp(out))
    return out


def cv_round(v: float) -> int:
    return lib().orc_cv_round(C.c_float(v))


def get_ied(row: np.ndarray, right_idx, left_idx) -> float:
    row = np.ascontiguousarray(row, dtype=np.float32).ravel()
    r = np.asarray(right_idx, dtype=np.int32)
    l = np.asarray(left_idx, dtype=np.int32)
    return lib().orc_get_ied(_fp(row), row.size // 2, _ip(r), r.size, _ip(l), l.size)


def crop_patch_u8(image: np.ndarray, cx: int, cy: int, half: int) -> np.ndarray:
    image = np.ascontiguousarray(image, dtype=np.uint8)
    h, w = image.shape
    out = np.zeros((2 * half, 2 * half), dtype=np.uint8)
    lib().orc_crop_patch_u8(_u8p(image), w, h, w, cx, cy, half, _u8p(out))
    return out


def patch_geometry(params: np.ndarray, p: HogParam, right_idx, left_idx):
    params = np.ascontiguousarray(params, dtype=np.float32).ravel()
    L = params.size // 2
    r = np.asarray(right_idx, dtype=np.int32)
    l = np.asarray(left_idx, dtype=np.int32)
    cx = np.zeros(L, np.int32); cy = np.zeros(L, np.int32); half = np.zeros(L, np.int32)
    lib().orc_patch_geometry(_fp(params), L, C.byref(p), _ip(r), r.size, _ip(l), l.size, _ip(cx), _ip(cy), _ip(half))
    return cx, cy, half


def feature_length(num_landmarks: int, p: HogParam) -> int:
    return lib().orc_feature_length(num_landmarks, C.byref(p))


def hog_transform(image: np.ndarray, params: np.ndarray, p: HogParam, right_idx, left_idx,
                  use_ref: bool = False) -> np.ndarray:
    """rcr::HogTransform::operator() on one 8UC1 image; returns the (D,) feature row."""
    image = np.ascontiguousarray(image, dtype=np.uint8)
    h, w = image.shape
    params = np.ascontiguousarray(params, dtype=np.float32).ravel()
    L = params.size // 2
    r = np.asarray(right_idx, dtype=np.int32)
    l = np.asarray(left_idx, dtype=np.int32)
    out = np.zeros(feature_length(L, p), dtype=np.float32)
    rc = lib().orc_hog_transform(_u8p(image), w, h, w, _fp(params), L, C.byref(p), _ip(r), r.size,
                                 _ip(l), l.size, _core_ptr(use_ref), _fp(out))
    if rc:
        raise RuntimeError(f"orc_hog_transform failed ({rc})")
    return out


def hog_transform_fixed(image: np.ndarray, params: np.ndarray, p: HogParam, use_ref: bool = False) -> np.ndarray:
    """The non-adaptive HogTransform of examples/landmark_detection.cpp:195-261 on one 8UC1 image (no bias column)."""
    image = np.ascontiguousarray(image, dtype=np.uint8)
    h, w = image.shape
    params = np.ascontiguousarray(params, dtype=np.float32).ravel()
    L = params.size // 2
    n = C.c_int(0)
    rc = lib().orc_hog_transform_fixed(_u8p(image), w, h, w, _fp(params), L, C.byref(p), _core_ptr(use_ref), None, C.byref(n))
    if rc:
        raise RuntimeError(f"orc_hog_transform_fixed failed ({rc})")
    out = np.zeros(n.value, dtype=np.float32)
    rc = lib().orc_hog_transform_fixed(_u8p(image), w, h, w, _fp(params), L, C.byref(p), _core_ptr(use_ref), _fp(out), None)
    if rc:
        raise RuntimeError(f"orc_hog_transform_fixed failed ({rc})")
    return out


def hog_transform_batch(images: np.ndarray, params: np.ndarray, p: HogParam, right_idx, left_idx,
                        use_ref: bool = False, threads: int = 1) -> np.ndarray:
    """One image per sample: images (N, h, w) u8, params (N, 2L) -> (N, D)."""
    images = np.ascontiguousarray(images, dtype=np.uint8)
    n, h, w = images.shape
    params = np.ascontiguousarray(params, dtype=np.float32)
    L = params.shape[1] // 2
    r = np.asarray(right_idx, dtype=np.int32)
    l = np.asarray(left_idx, dtype=np.int32)
    D = feature_length(L, p)
    out = np.zeros((n, D), dtype=np.float32)
    rc = lib().orc_hog_transform_batch(_u8p(images), n, w, h, w, _fp(params), L, C.byref(p), _ip(r), r.size,
                                       _ip(l), l.size, _core_ptr(use_ref), threads, _fp(out), D)
    if rc:
        raise RuntimeError(f"orc_hog_transform_batch failed ({rc})")
    return out


def align_mean(mean: np.ndarray, box, sx=1.0, sy=1.0, tx=0.0, ty=0.0) -> np.ndarray:
    mean = np.ascontiguousarray(mean, dtype=np.float32).ravel()
    out = np.zeros_like(mean)
    lib().orc_align_mean(_fp(mean), mean.size // 2, int(box[0]), int(box[1]), int(box[2]), int(box[3]),
                         C.c_float(sx), C.c_float(sy), C.c_float(tx), C.c_float(ty), _fp(out))
    return out


def gram(A: np.ndarray, precision: int = 0) -> np.ndarray:
    A = np.ascontiguousarray(A, dtype=np.float32)
    n, d = A.shape
    G = np.zeros((d, d), dtype=np.float32)
    lib().orc_gram(_fp(A), n, d, precision, _fp(G))
    return G


def regulariser_lambda(reg: Regulariser, AtA: np.ndarray, n_train: int) -> float:
    AtA = np.ascontiguousarray(AtA, dtype=np.float32)
    return lib().orc_regulariser_lambda(C.byref(reg), _fp(AtA), AtA.shape[0], n_train)


def solve(A: np.ndarray, B: np.ndarray, reg: Regulariser, precision: int = 0):
    """PartialPivLUSolver::solve.  Returns (X (D, M), lambda)."""
    A = np.ascontiguousarray(A, dtype=np.float32)
    B = np.ascontiguousarray(B, dtype=np.float32)
    n, d = A.shape
    m = B.shape[1]
    X = np.zeros((d, m), dtype=np.float32)
    lam = C.c_float(0)
    lib().orc_solve(_fp(A), _fp(B), n, d, m, C.byref(reg), precision, _fp(X), C.byref(lam))
    return X, lam.value


def predict(values: np.ndarray, X: np.ndarray) -> np.ndarray:
    values = np.ascontiguousarray(values, dtype=np.float32)
    X = np.ascontiguousarray(X, dtype=np.float32)
    n, d = values.shape
    m = X.shape[1]
    out = np.zeros((n, m), dtype=np.float32)
    lib().orc_predict(_fp(values), n, d, _fp(X), m, _fp(out))
    return out


def test_residual(data: np.ndarray, labels: np.ndarray, X: np.ndarray) -> float:
    data = np.ascontiguousarray(data, dtype=np.float32)
    labels = np.ascontiguousarray(labels, dtype=np.float32)
    X = np.ascontiguousarray(X, dtype=np.float32)
    return lib().orc_test_residual(_fp(data), _fp(labels), data.shape[0], data.shape[1], _fp(X), X.shape[1])


def _make_norm(norm):
    """norm: None (NoNormalisation) or (right_idx, left_idx) for InterEyeDistanceNormalisation."""
    if norm is None:
        return Normalisation(0, None, 0, None, 0), ()
    r = np.asarray(norm[0], dtype=np.int32)
    l = np.asarray(norm[1], dtype=np.int32)
    return Normalisation(1, _ip(r), r.size, _ip(l), l.size), (r, l)


def _wrap_projection(h, dims):
    def _cb(x_ptr, P, level, idx, out_ptr, _user):
        x = np.ctypeslib.as_array(x_ptr, shape=(P,))
        res = np.asarray(h(x.copy(), level, idx), dtype=np.float32).ravel()
        out = np.ctypeslib.as_array(out_ptr, shape=(dims[level],))
        out[:] = res
    return PROJECTION_FN(_cb)


def cascade_train(x_gt, x0, templates, regs, feat_dims, h, norm=None, precision=0, callback=None):
    """SupervisedDescentOptimiser::train.  h(x_row, level, idx) -> feature row.
    Returns (weights list, final x)."""
    x_gt = np.ascontiguousarray(x_gt, dtype=np.float32)
    x0 = np.ascontiguousarray(x0, dtype=np.float32)
    n, P = x0.shape
    S = len(regs)
    dims = (C.c_int * S)(*feat_dims)
    reg_arr = (Regulariser * S)(*regs)
    weights = [np.zeros((feat_dims[i], P), dtype=np.float32) for i in range(S)]
    wp = (C.POINTER(C.c_float) * S)(*[_fp(w) for w in weights])
    nm, keep = _make_norm(norm)
    proj = _wrap_projection(h, feat_dims)
    tp = None
    if templates is not None:
        templates = np.ascontiguousarray(templates, dtype=np.float32)
        tp = _fp(templates)
    xf = np.zeros_like(x0)
    cb = EPOCH_CB(lambda p, N, PP, lvl, u: callback(np.ctypeslib.as_array(p, shape=(N, PP)).copy(), lvl)) if callback else C.cast(None, EPOCH_CB)
    rc = lib().orc_cascade_train(_fp(x_gt), _fp(x0), tp, n, P, S, dims, reg_arr, C.byref(nm), proj, None,
                                 precision, wp, _fp(xf), cb, None)
    del keep
    return weights, xf, rc


def cascade_apply(x0, templates, weights, h, norm=None, callback=None):
    """SupervisedDescentOptimiser::test / predict."""
    x0 = np.ascontiguousarray(x0, dtype=np.float32)
    n, P = x0.shape
    S = len(weights)
    feat_dims = [w.shape[0] for w in weights]
    dims = (C.c_int * S)(*feat_dims)
    weights = [np.ascontiguousarray(w, dtype=np.float32) for w in weights]
    wp = (C.POINTER(C.c_float) * S)(*[_fp(w) for w in weights])
    nm, keep = _make_norm(norm)
    proj = _wrap_projection(h, feat_dims)
    tp = None
    if templates is not None:
        templates = np.ascontiguousarray(templates, dtype=np.float32)
        tp = _fp(templates)
    xf = np.zeros_like(x0)
    cb = EPOCH_CB(lambda p, N, PP, lvl, u: callback(np.ctypeslib.as_array(p, shape=(N, PP)).copy(), lvl)) if callback else C.cast(None, EPOCH_CB)
    lib().orc_cascade_apply(_fp(x0), tp, n, P, S, dims, wp, C.byref(nm), proj, None, _fp(xf), cb, None)
    del keep
    return xf


class Model:
    """A loaded rcr::detection_model (cereal binary)."""

    def __init__(self, path: str):
        err = C.create_string_buffer(256)
        self._m = lib().orc_model_load(path.encode(), err, 256)
        if not self._m:
            raise RuntimeError(err.value.decode() + ": " + path)
        m = self._m.contents
        self.num_levels = m.num_levels
        self.num_landmarks = m.num_landmarks
        self.weights = [np.ctypeslib.as_array(m.weights[i], shape=(m.rows[i], m.cols[i])).copy()
                        for i in range(m.num_levels)]
        self.regularisers = [(m.regularisers[i].type, m.regularisers[i].lambda_, m.regularisers[i].regularise_last_row)
                             for i in range(m.num_levels)]
        self.mean = np.ctypeslib.as_array(m.mean, shape=(2 * m.num_landmarks,)).copy()
        self.landmark_ids = [m.landmark_ids[i].decode() for i in range(m.num_landmarks)]
        self.hog_params = [HogParam(m.hog_params[i].variant, m.hog_params[i].num_cells, m.hog_params[i].cell_size,
                                    m.hog_params[i].num_bins, m.hog_params[i].relative_patch_size)
                           for i in range(m.num_levels)]
        self.right_idx = [m.right_idx[i] for i in range(m.n_right)]
        self.left_idx = [m.left_idx[i] for i in range(m.n_left)]
        self.right_ids = [m.right_ids[i].decode() for i in range(m.n_right)]
        self.left_ids = [m.left_ids[i].decode() for i in range(m.n_left)]

    def save(self, path: str) -> None:
        if lib().orc_model_save(self._m, path.encode()):
            raise RuntimeError("could not write " + path)

    def detect(self, image: np.ndarray, box, use_ref: bool = False) -> np.ndarray:
        image = np.ascontiguousarray(image, dtype=np.uint8)
        h, w = image.shape
        out = np.zeros(2 * self.num_landmarks, dtype=np.float32)
        rc = lib().orc_detect(self._m, _u8p(image), w, h, w, int(box[0]), int(box[1]), int(box[2]), int(box[3]),
                              _core_ptr(use_ref), _fp(out))
        if rc:
            raise RuntimeError(f"orc_detect failed ({rc})")
        return out

    def detect_init(self, image: np.ndarray, init: np.ndarray, use_ref: bool = False) -> np.ndarray:
        image = np.ascontiguousarray(image, dtype=np.uint8)
        h, w = image.shape
        init = np.ascontiguousarray(init, dtype=np.float32).ravel()
        out = np.zeros(2 * self.num_landmarks, dtype=np.float32)
        rc = lib().orc_detect_init(self._m, _u8p(image), w, h, w, _fp(init), _core_ptr(use_ref), _fp(out))
        if rc:
            raise RuntimeError(f"orc_detect_init failed ({rc})")
        return out

    def detect_batch(self, images: np.ndarray, boxes: np.ndarray, use_ref: bool = False, threads: int = 1) -> np.ndarray:
        images = np.ascontiguousarray(images, dtype=np.uint8)
        n, h, w = images.shape
        boxes = np.ascontiguousarray(boxes, dtype=np.int32)
        out = np.zeros((n, 2 * self.num_landmarks), dtype=np.float32)
        rc = lib().orc_detect_batch(self._m, _u8p(images), n, w, h, w, _ip(boxes), _core_ptr(use_ref), threads, _fp(out))
        if rc:
            raise RuntimeError(f"orc_detect_batch failed ({rc})")
        return out

    def __del__(self):
        try:
            if self._m:
                lib().orc_model_free(self._m)
                self._m = None
        except Exception:
            pass
