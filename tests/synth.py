"""Seeded synthetic inputs shared by the tests (numpy only, small sizes)."""
import numpy as np


def smooth_images(count, h, w, seed, sigma=3.0):
    """Low-pass filtered uniform noise stretched to 0..255, 8UC1."""
    rng = np.random.default_rng(seed)
    r = int(3 * sigma)
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2)
    k /= k.sum()
    out = np.empty((count, h, w), dtype=np.uint8)
    for i in range(count):
        img = rng.random((h + 2 * r, w + 2 * r))
        img = np.apply_along_axis(lambda v: np.convolve(v, k, mode="valid"), 0, img)
        img = np.apply_along_axis(lambda v: np.convolve(v, k, mode="valid"), 1, img)
        img = (img - img.min()) / (img.max() - img.min())
        out[i] = np.clip(np.round(img * 255.0), 0, 255).astype(np.uint8)
    return out


def face_boxes(count, h, w, seed, border_fraction=0.0):
    """One square face box per frame; `border_fraction` of them hang over the image border."""
    rng = np.random.default_rng(seed + 1)
    boxes = np.empty((count, 4), dtype=np.int32)
    for i in range(count):
        s = int(rng.integers(min(h, w) // 3, min(h, w) // 2 + 1))
        if rng.random() < border_fraction:
            x = int(rng.integers(-s // 3, w - s // 2))
            y = int(rng.integers(-s // 3, h - s // 2))
        else:
            x = int(rng.integers(s // 8, max(s // 8 + 1, w - s - s // 8)))
            y = int(rng.integers(s // 8, max(s // 8 + 1, h - s - s // 8)))
        boxes[i] = (x, y, s, s)
    return boxes
