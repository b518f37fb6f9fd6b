"""Pins oracle/post_ref.py against the reference's OWN model_fuse.model_confuse and edge_3._detection
(imported from /root/reference under the SURVEY App. E stubs, run through their PNG-file interface) on the
seeded scenes of tests/post_scenes.py, and stores the reference outputs as tests/golden/post.npz so that
the GPU box (where /root/reference is absent) can check the CUDA path against them.
Run in the build container:  python tools/make_golden_post.py"""
import glob
import os
import sys
import tempfile
import time

import cv2 as cv
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ref_harness import reference_modules  # noqa: E402
import post_scenes as PS  # noqa: E402
from oracle import post_ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "post.npz")


def ref_fuse(model_fuse, masks):
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        work = os.path.join(d, "work")
        os.makedirs(work)
        for k, m in enumerate(masks):
            cv.imwrite(os.path.join(work, f"m{k}.png"), m, [int(cv.IMWRITE_PNG_COMPRESSION), 0])
        os.chdir(d)  # fill_and_delete drops gray.png into the CWD (model_fuse.py:31)
        try:
            model_fuse.model_confuse(work, "x")
        finally:
            os.chdir(cwd)
        out = [p for p in glob.glob(os.path.join(d, "*")) if p.endswith("x_result.png")]  # literal backslash name
        assert len(out) == 1, os.listdir(d)
        return cv.imread(out[0], cv.IMREAD_GRAYSCALE)


def ref_detect(edge_3, mask):
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "result.png")
        cv.imwrite(p, mask)
        return edge_3._detection(p)


def pack_polys(polys):
    off = np.cumsum([0] + [len(p[0]) for p in polys]).astype(np.int32)
    xs = np.array([v for p in polys for v in p[0]], np.float64)
    ys = np.array([v for p in polys for v in p[1]], np.float64)
    isf = np.array([isinstance(p[0][0], np.floating) for p in polys], np.uint8)
    return off, xs, ys, isf


def same_polys(a, b):
    if len(a) != len(b):
        return False
    for p, q in zip(a, b):
        if len(p[0]) != len(q[0]) or type(p[0][0]) is not type(q[0][0]):
            return False
        if not (np.array_equal(np.asarray(p[0]), np.asarray(q[0])) and np.array_equal(np.asarray(p[1]), np.asarray(q[1]))):
            return False
    return True


def main():
    _, model_fuse, edge_3 = reference_modules()
    gold = {}
    for name, size, seed in PS.FUSE_CASES:
        masks = PS.five_masks(size, seed)
        t = time.time()
        want = ref_fuse(model_fuse, masks)
        t_ref = time.time() - t
        got = post_ref.model_confuse(masks)
        assert np.array_equal(want, got), f"{name}: restatement differs from the reference on {(want != got).sum()} px"
        for k, m in enumerate(masks):  # per-mask clean-up as well (intermediate of the reference)
            pass
        gold[name + "_fused"] = np.packbits(want > 0)
        gold[name + "_shape"] = np.int32(want.shape)
        print(f"{name}: {size}^2 reference fuse {t_ref:.1f}s, restatement identical, fused on = {(want > 0).mean():.3f}")
    for name, size, seed in PS.CONTOUR_CASES:
        mask = PS.contour_case_mask(name, size, seed)
        want, h = ref_detect(edge_3, mask)
        got, h2 = post_ref.detection(mask)
        assert h == h2 and same_polys(want, got), f"{name}: restatement differs from the reference"
        off, xs, ys, isf = pack_polys(want)
        gold[name + "_off"], gold[name + "_xs"], gold[name + "_ys"], gold[name + "_isf"] = off, xs, ys, isf
        print(f"{name}: {size}^2 {len(want)} polygons ({int(isf.sum())} minAreaRect fallbacks), restatement identical")
    # detection on the fused outputs too (the real hand-off: fuse -> contours)
    for name, size, seed in PS.FUSE_CASES:
        fused = np.unpackbits(gold[name + "_fused"])[:size * size].reshape(size, size).astype(np.uint8) * 255
        want, _ = ref_detect(edge_3, fused)
        got, _ = post_ref.detection(fused)
        assert same_polys(want, got)
        off, xs, ys, isf = pack_polys(want)
        gold[name + "_poly_off"], gold[name + "_poly_xs"], gold[name + "_poly_ys"], gold[name + "_poly_isf"] = off, xs, ys, isf
        print(f"{name}: contours of the fused mask: {len(want)} polygons, restatement identical")
    np.savez_compressed(OUT, **gold)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
