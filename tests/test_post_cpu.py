"""CPU tests of the post-processing oracle and of the library's host-side geometry (no GPU needed).

 * oracle/post_ref.py reproduces the golden outputs made by the reference's own model_fuse.model_confuse and
   edge_3._detection (tests/golden/post.npz, tools/make_golden_post.py);
 * the library's restatements of cv::contourArea / arcLength / approxPolyDP and of the area-tiered choice of
   edge_3.py:351-378 agree exactly with cv2 on thousands of random contours (differential test);
 * numpy twins of the two non-obvious device algorithms -- the tracing-free polygon area and the border
   following -- agree with cv2.contourArea / cv2.findContours.
"""
import ctypes as C
import os

import cv2 as cv
import numpy as np
import pytest

import post_scenes as PS
from building_detection_b200 import runtime as R
from oracle import post_ref

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "post.npz"))


def unpack_polys(prefix):
    off, xs, ys, isf = (GOLD[prefix + k] for k in ("off", "xs", "ys", "isf"))
    return [(xs[off[i]:off[i + 1]], ys[off[i]:off[i + 1]], bool(isf[i])) for i in range(len(off) - 1)]


def same_as_golden(polys, prefix):
    want = unpack_polys(prefix)
    assert len(polys) == len(want)
    for (xs, ys), (wx, wy, isf) in zip(polys, want):
        assert isinstance(xs[0], np.floating) == isf
        np.testing.assert_array_equal(np.asarray(xs, np.float64), wx)
        np.testing.assert_array_equal(np.asarray(ys, np.float64), wy)


@pytest.mark.parametrize("name,size,seed", PS.FUSE_CASES)
def test_oracle_fuse_matches_reference_golden(name, size, seed):
    want = np.unpackbits(GOLD[name + "_fused"])[:size * size].reshape(size, size).astype(np.uint8) * 255
    got = post_ref.model_confuse(PS.five_masks(size, seed))
    np.testing.assert_array_equal(got, want)
    same_as_golden(post_ref.detection(got)[0], name + "_poly_")


@pytest.mark.parametrize("name,size,seed", PS.CONTOUR_CASES)
def test_oracle_contours_match_reference_golden(name, size, seed):
    polys, h = post_ref.detection(PS.contour_case_mask(name, size, seed))
    assert h == size
    same_as_golden(polys, name + "_")


def random_contours(seed, n_masks=40):
    rng = np.random.default_rng(seed)
    for _ in range(n_masks):
        size = int(rng.integers(12, 160))
        kind = rng.integers(0, 3)
        if kind == 0:
            m = PS.noise_mask(size, int(rng.integers(1 << 30)), float(rng.uniform(0.2, 0.8)), int(rng.choice([1, 3, 5, 9, 15])))
        elif kind == 1:
            m = PS.base_mask(max(size, 48), int(rng.integers(1 << 30)), n_objects=int(rng.integers(1, 6)))
        else:
            m = np.zeros((size, size), np.uint8)
            for _k in range(int(rng.integers(1, 4))):
                c = (int(rng.integers(0, size)), int(rng.integers(0, size)))
                cv.ellipse(m, c, (int(rng.integers(2, size // 2 + 3)), int(rng.integers(2, size // 2 + 3))),
                           float(rng.uniform(0, 180)), 0, 360, 255, -1)
        for c in cv.findContours(m, cv.RETR_EXTERNAL, cv.CHAIN_APPROX_NONE)[0]:
            yield c


def test_host_geometry_matches_cv2():
    L = R.lib()
    rng = np.random.default_rng(5)
    n = 0
    for c in random_contours(77, 120):
        pts = np.ascontiguousarray(c.reshape(-1, 2), np.int32)
        k = len(pts)
        assert L.bd_host_contour_area(R._ptr(pts), k) == cv.contourArea(c)
        per = cv.arcLength(c, True)
        assert L.bd_host_arc_length(R._ptr(pts), k) == per
        out = np.empty((k + 2, 2), np.int32)
        for eps in (0.0, 0.5, 1.0, 0.002 * per, 0.01 * per, 0.05 * per, float(rng.uniform(0, 6))):
            m = L.bd_host_approx_poly(R._ptr(pts), k, C.c_double(eps), R._ptr(out))
            want = cv.approxPolyDP(c, eps, True).reshape(-1, 2)
            assert m == len(want) and np.array_equal(out[:m], want), (k, eps)
        n += 1
    assert n > 300


def test_host_simplify_matches_oracle():
    """The area-tiered choice (edge_3.py:351-378), including the 11-try small_target search; kind 2 = the oracle
    fell back to minAreaRect."""
    L = R.lib()
    seen = {0: 0, 1: 0, 2: 0}
    for c in random_contours(99, 150):
        pts = np.ascontiguousarray(c.reshape(-1, 2), np.int32)
        out = np.empty((len(pts) + 2, 2), np.int32)
        m = C.c_int()
        kind = L.bd_host_simplify(R._ptr(pts), len(pts), R._ptr(out), C.byref(m))
        want = post_ref.simplify(c)
        seen[kind] += 1
        if want is None:
            assert kind == 0
        elif want.dtype == np.float32:
            assert kind == 2
        else:
            assert kind == 1 and np.array_equal(out[:m.value], want)
    assert min(seen.values()) > 5, seen


def area2_twin(comp):
    """numpy twin of ccl::polygon_area2 for one hole-free component (bool array)."""
    H, W = comp.shape
    p = np.zeros((H + 2, W + 2), bool)
    p[1:-1, 1:-1] = comp
    a, b, c, d = p[0:H + 1, 0:W + 1], p[0:H + 1, 1:W + 2], p[1:H + 2, 0:W + 1], p[1:H + 2, 1:W + 2]
    ys, xs = np.mgrid[0:H + 1, 0:W + 1]
    P = {"a": (xs - 1, ys - 1), "b": (xs, ys - 1), "c": (xs - 1, ys), "d": (xs, ys)}
    steps = [(a & b & ~c & ~d, "a", "b"), (c & d & ~a & ~b, "d", "c"), (a & c & ~b & ~d, "c", "a"),
             (b & d & ~a & ~c, "b", "d"), (a & b & c & ~d, "c", "b"), (a & b & d & ~c, "a", "d"),
             (a & c & d & ~b, "d", "a"), (b & c & d & ~a, "b", "c")]
    tot = 0
    for mask, s, e in steps:
        tot += int((P[s][0] * P[e][1] - P[e][0] * P[s][1])[mask].sum())
    return tot


DX, DY = [1, 1, 0, -1, -1, -1, 0, 1], [0, -1, -1, -1, 0, 1, 1, 1]


def trace_twin(img, x0, y0):
    """python twin of cont::trace_one."""
    H, W = img.shape
    at = lambda x, y: 0 <= x < W and 0 <= y < H and img[y, x] != 0  # noqa: E731
    s = 4
    while True:
        s = (s - 1) & 7
        x1, y1 = x0 + DX[s], y0 + DY[s]
        if at(x1, y1) or s == 4:
            break
    if not at(x1, y1):
        return [(x0, y0)]
    pts, x3, y3 = [], x0, y0
    while True:
        while True:
            s += 1
            x4, y4 = x3 + DX[s & 7], y3 + DY[s & 7]
            if at(x4, y4):
                break
        s &= 7
        pts.append((x3, y3))
        if (x4, y4) == (x0, y0) and (x3, y3) == (x1, y1):
            break
        x3, y3, s = x4, y4, (s + 4) & 7
    return pts


def test_device_algorithm_twins_match_cv2():
    rng = np.random.default_rng(3)
    n = 0
    for _ in range(60):
        size = int(rng.integers(6, 70))
        m = PS.noise_mask(size, int(rng.integers(1 << 30)), float(rng.uniform(0.2, 0.8)), int(rng.choice([1, 3, 5])))
        cs = cv.findContours(m.copy(), cv.RETR_EXTERNAL, cv.CHAIN_APPROX_NONE)[0]
        filled = np.zeros_like(m)
        for c in cs:
            cv.drawContours(filled, [c], 0, 255, cv.FILLED)
        prev = None
        for c in cs:
            one = np.zeros_like(m)
            cv.drawContours(one, [c], 0, 255, cv.FILLED)
            assert abs(area2_twin(one > 0)) / 2 == cv.contourArea(c)
            x0, y0 = (int(v) for v in c[0, 0])
            assert prev is None or (y0, x0) < prev  # findContours lists contours by descending start pixel
            prev = (y0, x0)
            assert np.array_equal(np.array(trace_twin(filled, x0, y0)).reshape(-1, 1, 2), c)
            n += 1
    assert n > 500


def test_contour_length_is_bounded_by_boundary_cracks():
    """The one-walk tracer of csrc/contours.cu sizes every contour's slot by the component's number of boundary
    cracks (pixel edges towards background or the frame); cv2's CHAIN_APPROX_NONE external contour must never be
    longer than that (holes only add cracks)."""
    import cv2
    rng = np.random.default_rng(0)
    worst = 0.0
    for trial in range(120):
        H, W = rng.integers(8, 64, 2)
        m = (rng.random((H, W)) < rng.choice([0.3, 0.5, 0.6, 0.7, 0.9])).astype(np.uint8)
        if trial % 3 == 0:
            m = cv2.dilate(m, np.ones((3, 3), np.uint8))
        n, lab = cv2.connectedComponents(m, connectivity=8)
        for k in range(1, n):
            comp = (lab == k).astype(np.uint8)
            cs, _ = cv2.findContours(comp, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
            pad = np.pad(comp, 1)
            c = pad[1:-1, 1:-1] == 1
            cracks = int((c & (pad[1:-1, :-2] == 0)).sum() + (c & (pad[1:-1, 2:] == 0)).sum()
                         + (c & (pad[:-2, 1:-1] == 0)).sum() + (c & (pad[2:, 1:-1] == 0)).sum())
            assert len(cs) == 1 and len(cs[0]) <= cracks
            worst = max(worst, len(cs[0]) / cracks)
    assert worst > 0.9  # the bound is tight enough to be worth using
