"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement of the reference's post-processing, model_fuse.py (3-of-5 mask fusion with per-mask
clean-up) and edge_3.py (contour extraction and polygon simplification).  The arithmetic of both lives
in OpenCV (un-vendored third-party dependency, un-pinned by the reference; written against <= 4.5.3 list
semantics).  This restatement calls the same cv2 primitives -- opencv-python-headless 4.13.0 in this image,
which IS the de-facto pin -- in the order the reference does, but works on in-memory arrays instead of
PNG files and has none of the reference's side effects (plt.imshow, gray.png, rectangles drawn for
debugging).  Every function cites the reference lines it follows.

Pinned: tools/make_golden_post.py runs the reference's own model_fuse.model_confuse and
edge_3._detection (imported from /root/reference under the SURVEY App. E stubs) on the seeded scenes of
tests/post_scenes.py, asserts this restatement reproduces them exactly, and stores the outputs under
tests/golden/ (the reference tree is absent on the GPU box).
"""
from __future__ import annotations

import cv2 as cv
import numpy as np


def _contours(img):
    """cv.findContours(RETR_EXTERNAL, CHAIN_APPROX_NONE) as a list (cv2 >= 4.5.4 returns a tuple)."""
    res = cv.findContours(img, mode=cv.RETR_EXTERNAL, method=cv.CHAIN_APPROX_NONE)
    return list(res[0] if len(res) == 2 else res[1])


# =============================================================================== model_fuse.py
def fill_and_delete(mask, min_area=1000):
    """model_fuse.py:9-32.  mask: (H,W) u8.  Every external contour is filled (holes vanish); polygons of
    area <= min_area are erased again.  Returns (cleaned mask, its external contours)."""
    g = mask.copy()
    cs = _contours(g)
    for i, c in enumerate(cs):
        area = cv.contourArea(c)
        cv.fillPoly(g, [c], (255, 255, 255))
        if area <= min_area:
            cv.drawContours(g, [c], 0, 0, cv.FILLED)  # (= drawContours(g, cs, i, ...): O(1) instead of O(len(cs)) per call)
    return g, _contours(g)


def _split_one_direction(obj, kernel, iters=5, frag_area=500):
    """model_fuse.py:65-117 (erode_process / erode_process1 differ only in the kernel orientation), with
    fill_small_target (:52-62) and dilate_process (:35-49) inlined.
    Returns None (one fragment: no split), False (fragments existed but all were <= frag_area) or the list
    of contours of the surviving fragments, each dilated back on its own."""
    er = cv.erode(obj, kernel, iterations=iters)
    frags = _contours(er)
    if len(frags) == 1:
        return None
    erased = False
    for i, c in enumerate(frags):
        area = cv.contourArea(c)
        cv.fillPoly(er, [c], (255, 255, 255))
        if area <= frag_area:
            erased = True
            cv.drawContours(er, [c], 0, 0, cv.FILLED)
    if erased:
        frags = _contours(er)
        if len(frags) == 0:
            return False
    out = []
    for j in range(len(frags)):
        one = np.zeros_like(obj)
        cv.drawContours(one, frags, j, 255, cv.FILLED)
        out.append(_contours(cv.dilate(one, kernel, iterations=iters))[0])
    return out


def erode_dilate_split(shape, contours):
    """model_fuse.py:173-218 (eroede_dilate_process): per object, try to split it at bridges narrower than
    21 px horizontally (1x5 kernel, 5 iterations) and vertically; combine as the reference does, including
    the quirk that an empty list is neither None nor False."""
    h, w = shape
    kept = []
    kh, kv = np.ones((1, 5), np.uint8), np.ones((5, 1), np.uint8)
    for i in range(len(contours)):
        obj = np.zeros((h, w), np.uint8)
        cv.drawContours(obj, contours, i, 255, cv.FILLED)
        a = _split_one_direction(obj, kh)
        b = _split_one_direction(obj, kv)
        if a is False or b is False:
            continue
        if a is None and b is None:
            kept.append(contours[i])
        else:
            kept.extend(a or [])
            kept.extend(b or [])
    return kept


def clean_mask(mask):
    """fill_and_delete -> eroede_dilate_process -> only_plt (model_fuse.py:285-289): (H,W) u8 {0,255}."""
    g, cs = fill_and_delete(mask)
    pieces = erode_dilate_split(g.shape, cs)
    out = np.zeros(g.shape, np.uint8)
    for j in range(len(pieces)):
        cv.drawContours(out, pieces, j, 255, cv.FILLED)
    return out


def model_confuse(masks5):
    """model_fuse.py:271-350 on arrays: masks5 = five (H,W) u8 masks (any order: the vote is symmetric).
    Returns the fused (H,W) u8 {0,255} mask the reference writes to ``<path>\\<name>_result.png``."""
    assert len(masks5) == 5
    votes = sum((clean_mask(m) // 255).astype(np.int32) for m in masks5)  # :315
    fused = np.where(votes >= 3, 255, 0).astype(np.uint8)  # :323-324
    return clean_mask(fused)  # :339-346


# =============================================================================== edge_3.py
def _bbox(c, idx):
    x, y, w, h = cv.boundingRect(c)
    return [x, y, x + w, y + h, idx]


def _best_iou(box, others):
    """edge_3.py:26-47: index of the best-overlapping box of ``others`` if its IoU exceeds 0.5, else None."""
    box = np.array(box)
    others = np.array(others)
    lt = np.maximum(box[:2], others[:, :2])
    rb = np.minimum(box[2:4], others[:, 2:4])
    wh = np.maximum(rb - lt, 0)
    inter = wh[:, 0] * wh[:, 1]
    union = (box[2] - box[0]) * (box[3] - box[1]) + (others[:, 2] - others[:, 0]) * (others[:, 3] - others[:, 1]) - inter
    v = inter / union
    return int(np.argmax(v)) if np.any(v > 0.5) else None


def _match(initial, eroded):
    """edge_3.py:50-85 / :88-121 (process_td / process_rl; the latter tolerates None entries):
    -> (boxes of initial contours with no eroded counterpart, boxes of eroded contours nobody claimed)."""
    ib = [[0, 0, 0, 0, j] if c is None else _bbox(c, j) for j, c in enumerate(initial)]
    eb = [_bbox(c, j) for j, c in enumerate(eroded)]
    lost, claimed = [], []
    for b in ib:
        r = _best_iou(b, eb)
        if r is None:
            lost.append(b)
        else:
            claimed.append(r)
    new = [eb[j] for j in range(len(eroded)) if j not in claimed]
    return lost, new


def _eroded_contours(img3, kernel):
    """edge_3.py:172-185 + erode_images_process (:124-144): erode once, erase fragments of area < 50."""
    er = cv.erode(img3, kernel, iterations=1)
    cs = _contours(cv.cvtColor(er, cv.COLOR_BGR2GRAY))
    for i, c in enumerate(cs):
        if cv.contourArea(c) < 50:
            cv.drawContours(er, cs, i, 0, cv.FILLED)
    return _contours(er[:, :, 0].copy())


def split_corner_touching(img3, k=7):
    """edge_3.py:159-262 (detction_overlap_building): buildings that only touch at a corner fall apart under
    a 1xk / kx1 erosion; the originals that lost their counterpart become None and the unclaimed eroded
    contours are appended as they are (not dilated back)."""
    res = _contours(cv.cvtColor(img3, cv.COLOR_BGR2GRAY))
    n = len(res)
    td = _eroded_contours(img3, np.ones((1, k), np.uint8))
    rl = _eroded_contours(img3, np.ones((k, 1), np.uint8))
    if len(td) == n and len(rl) == n:
        return res
    lost_td = new_td = lost_rl = new_rl = None
    if len(td) != n:
        lost_td, new_td = _match(res, td)
    if len(rl) != n:
        lost_rl, new_rl = _match(res, rl)  # :218 runs before any entry is set to None
    for lost in (lost_td, lost_rl):
        if lost is not None:
            for b in lost:
                res[b[4]] = None
    if new_td is not None and new_rl is not None:
        if len(new_td) >= 1 and len(new_rl) >= 1:
            dup = []
            for b in new_td:
                r = _best_iou(b, new_rl)
                res.append(td[b[4]])
                if r is not None:
                    dup.append(r)
            for i, b in enumerate(new_rl):
                if i not in dup:
                    res.append(rl[b[4]])
        elif len(new_td) >= 1:
            res.extend(td[b[4]] for b in new_td)
        else:
            res.extend(rl[b[4]] for b in new_rl)
    elif new_td is not None:
        res.extend(td[b[4]] for b in new_td)
    else:
        res.extend(rl[b[4]] for b in new_rl)
    return res


def _small_target(c, eps):
    """edge_3.py:265-286: up to 11 retries with eps = 0.002k * perimeter until the polygon has 4 vertices,
    otherwise the minimum-area rectangle."""
    pts = cv.approxPolyDP(c, eps, True).reshape((-1, 2))
    rate, tries = 0.002, 0
    while len(pts) != 4:
        eps = rate * cv.arcLength(c, True)
        rate = rate + 0.002
        pts = cv.approxPolyDP(c, eps, True).reshape((-1, 2))
        tries += 1
        if tries > 10:
            break
    if len(pts) != 4:
        pts = cv.boxPoints(cv.minAreaRect(c))
    return pts


def simplify(c):
    """edge_3.py:351-378: area-tiered polygon approximation of one contour, or None when it is skipped."""
    area = cv.contourArea(c)
    per = cv.arcLength(c, True)
    eps = 0.01 * per
    if cv.moments(c)["m00"] <= 10:
        return None
    if area < 150:
        return _small_target(c, eps)
    if 150 < area < 300:
        return cv.approxPolyDP(c, 5 * eps, True).reshape((-1, 2))
    if 3000 < area < 8000:
        return cv.approxPolyDP(c, 0.005 * per, True).reshape((-1, 2))
    if 8000 < area <= 15000:
        return cv.approxPolyDP(c, 0.004 * per, True).reshape((-1, 2))
    if area > 15000:
        return cv.approxPolyDP(c, 0.002 * per, True).reshape((-1, 2))
    return cv.approxPolyDP(c, eps, True).reshape((-1, 2))


def detection(mask):
    """edge_3.py:310-387 (_detection) on an (H,W) u8 mask instead of a PNG path.
    Returns (polygons, H); polygons[i] = [xs, ys], closed by repeating the first vertex; coordinates are
    np.int32, or np.float32 for the minAreaRect fallback."""
    img3 = cv.cvtColor(mask, cv.COLOR_GRAY2BGR)  # what cv.imread gives for a grey PNG
    cs = _contours(mask.copy())
    for i, c in enumerate(cs):
        area = cv.contourArea(c)
        cv.fillPoly(img3, [c], (255, 255, 255))
        if area <= 100:
            cv.drawContours(img3, cs, i, 0, cv.FILLED)
    polys = []
    for c in split_corner_touching(img3, 7):
        if c is None:
            continue
        pts = simplify(c)
        if pts is None:
            continue
        polys.append([list(pts[:, 0]) + [pts[0, 0]], list(pts[:, 1]) + [pts[0, 1]]])
    return polys, mask.shape[0]
