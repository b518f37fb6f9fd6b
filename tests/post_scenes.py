"""Seeded synthetic masks for the fusion / contour parity tests (the reference ships no fixtures).

``base_mask`` draws building-like shapes chosen to hit every branch of model_fuse.py and edge_3.py:
axis-aligned and rotated rectangles, L shapes, buildings joined by bridges narrower and wider than the
21-px split element, buildings touching only at a corner (edge_3's 7-px erosion case), holes, nested
islands inside holes, thin bars (< 21 px wide: the App. D #10 drop), specks around the 100 / 500 / 1000
area thresholds, shapes touching the image border, 1-px lines and single pixels.
``five_masks`` perturbs a base mask five ways (shift, drop, add, erode/dilate, noise) like five models
that mostly agree.
"""
import cv2 as cv
import numpy as np


def _rot_rect(img, rng, cx, cy, w, h, ang):
    box = cv.boxPoints(((float(cx), float(cy)), (float(w), float(h)), float(ang)))
    cv.fillPoly(img, [np.round(box).astype(np.int32)], 255)


def base_mask(size, seed, n_objects=None):
    rng = np.random.default_rng(seed)
    H = W = size
    m = np.zeros((H, W), np.uint8)
    n = n_objects if n_objects is not None else max(6, (size * size) // 22000)
    for _ in range(n):
        kind = rng.integers(0, 12)
        cx, cy = int(rng.integers(0, W)), int(rng.integers(0, H))
        w, h = int(rng.integers(24, 90)), int(rng.integers(24, 90))
        if kind <= 2:  # axis-aligned rectangle
            m[max(cy - h // 2, 0):cy + h // 2, max(cx - w // 2, 0):cx + w // 2] = 255
        elif kind <= 4:  # rotated rectangle
            _rot_rect(m, rng, cx, cy, w, h, rng.uniform(0, 180))
        elif kind == 5:  # two buildings joined by a bridge (narrow or wide)
            bw = int(rng.choice([3, 8, 14, 19, 20, 21, 22, 30]))
            m[max(cy - 30, 0):cy + 30, max(cx - 70, 0):max(cx - 20, 0)] = 255
            m[max(cy - 30, 0):cy + 30, cx + 20:cx + 70] = 255
            m[max(cy - bw // 2, 0):cy - bw // 2 + bw, max(cx - 20, 0):cx + 20] = 255
        elif kind == 6:  # same, vertical
            bw = int(rng.choice([3, 8, 14, 19, 20, 21, 22, 30]))
            m[max(cy - 70, 0):max(cy - 20, 0), max(cx - 30, 0):cx + 30] = 255
            m[cy + 20:cy + 70, max(cx - 30, 0):cx + 30] = 255
            m[max(cy - 20, 0):cy + 20, max(cx - bw // 2, 0):cx - bw // 2 + bw] = 255
        elif kind == 7:  # two squares touching at a corner (8-connected only through the diagonal)
            s = int(rng.integers(30, 60))
            o = int(rng.integers(0, 4))
            m[max(cy - s, 0):cy, max(cx - s, 0):cx] = 255
            m[max(cy - o, 0):cy + s - o, max(cx - o, 0):cx + s - o] = 255
        elif kind == 8:  # building with a hole, sometimes an island inside the hole
            m[max(cy - 45, 0):cy + 45, max(cx - 45, 0):cx + 45] = 255
            m[max(cy - 25, 0):cy + 25, max(cx - 25, 0):cx + 25] = 0
            if rng.random() < 0.5:
                m[max(cy - 8, 0):cy + 8, max(cx - 8, 0):cx + 8] = 255
        elif kind == 9:  # thin bar (drops out of the fusion), L shape
            if rng.random() < 0.5:
                t = int(rng.integers(5, 24))
                m[max(cy - 60, 0):cy + 60, max(cx - t // 2, 0):cx - t // 2 + t] = 255
            else:
                m[max(cy - 50, 0):cy + 50, max(cx - 50, 0):max(cx - 10, 0)] = 255
                m[cy + 10:cy + 50, max(cx - 50, 0):cx + 50] = 255
        elif kind == 10:  # specks around the area thresholds
            for _k in range(3):
                s = int(rng.choice([3, 5, 9, 11, 12, 13, 22, 23, 24, 31, 32, 33]))
                px, py = int(rng.integers(0, W - 40)), int(rng.integers(0, H - 40))
                m[py:py + s, px:px + s + int(rng.integers(0, 3))] = 255
        else:  # lines and single pixels
            px, py = int(rng.integers(2, W - 40)), int(rng.integers(2, H - 40))
            m[py, px:px + int(rng.integers(1, 30))] = 255
            m[py + 4:py + 4 + int(rng.integers(1, 30)), px] = 255
            for _k in range(int(rng.integers(1, 12))):
                m[py + 8 + _k, px + 8 + _k] = 255  # diagonal chain
    return m


def five_masks(size, seed):
    rng = np.random.default_rng(seed + 7919)
    base = base_mask(size, seed)
    out = []
    for k in range(5):
        m = base.copy()
        if k == 1:
            m = np.roll(m, (int(rng.integers(-3, 4)), int(rng.integers(-3, 4))), axis=(0, 1))
        elif k == 2:
            m = cv.erode(m, np.ones((3, 3), np.uint8))
        elif k == 3:
            m = cv.dilate(m, np.ones((3, 3), np.uint8))
        # random drops / additions / salt noise, different per model
        for _ in range(max(2, size // 200)):
            y, x = int(rng.integers(0, size - 60)), int(rng.integers(0, size - 60))
            if rng.random() < 0.5:
                m[y:y + int(rng.integers(10, 60)), x:x + int(rng.integers(10, 60))] = 0
            else:
                m[y:y + int(rng.integers(10, 60)), x:x + int(rng.integers(10, 60))] = 255
        noise = rng.random(m.shape) < 0.002
        m[noise] = 255 - m[noise]
        out.append(m)
    return out


def noise_mask(size, seed, p=0.5, blur=5):
    """Unstructured mask like the argmax of a random-init network: smoothed noise thresholded at p."""
    rng = np.random.default_rng(seed)
    f = cv.blur(rng.random((size, size)).astype(np.float32), (blur, blur))
    return np.where(f > np.quantile(f, 1 - p), 255, 0).astype(np.uint8)


# (name, size, seed): the sets the golden file holds
FUSE_CASES = [("fuse_a", 512, 11), ("fuse_b", 640, 12), ("fuse_c", 384, 13)]
CONTOUR_CASES = [("cont_a", 640, 21), ("cont_b", 768, 22), ("cont_c", 300, 23), ("cont_noise", 256, 24)]


def contour_case_mask(name, size, seed):
    return noise_mask(size, seed, 0.35, 9) if name == "cont_noise" else base_mask(size, seed)
