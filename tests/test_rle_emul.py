"""CPU tests of the run-based fusion kernels: building_detection_b200/csrc/rle.cuh compiled for the host
(tests/emul) and compared with OpenCV / the oracle on seeded masks -- labelling, hole fill, polygon areas, line
morphology, one clean-up pass stage by stage, and the whole model_confuse against the reference's golden outputs.
The same kernels run on the GPU in tests/test_post_gpu.py; this file catches logic errors without a GPU."""
import ctypes as C
import os
import subprocess

import cv2 as cv
import numpy as np
import pytest

import post_scenes as PS
from oracle import post_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "post.npz"))


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emul") / "librle_emul.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas",
                           "-I" + os.path.join(ROOT, "tests", "emul"), "-I" + os.path.join(ROOT, "building_detection_b200", "csrc"),
                           os.path.join(ROOT, "tests", "emul", "rle_emul.cpp"), "-o", so])
    return C.CDLL(so)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def cleanup(lib, m, stage=-1):
    m = np.ascontiguousarray(m, np.uint8)
    out = np.empty_like(m)
    lib.emul_cleanup(_p(m), m.shape[0], m.shape[1], stage, _p(out))
    return out


def labels(lib, m, fg, conn8):
    m = np.ascontiguousarray(m, np.uint8)
    out = np.empty(m.shape, np.int32)
    lib.emul_labels(_p(m), m.shape[0], m.shape[1], fg, conn8, _p(out))
    return out


def canon(lab):
    """component labels -> raster index of each component's first pixel (-1 for background label 0)"""
    flat = lab.ravel()
    first = np.full(lab.max() + 1, -1, np.int64)
    idx = np.arange(flat.size)
    order = np.argsort(flat, kind="stable")
    vals, start = np.unique(flat[order], return_index=True)
    first[vals] = idx[order][start]
    out = first[flat].reshape(lab.shape)
    out[lab == 0] = -1
    return out.astype(np.int32)


MASKS = [("base", 333, 1), ("base", 512, 2), ("base", 700, 3), ("noise", 257, 4), ("noise", 384, 5), ("noise1", 200, 6),
         ("base", 96, 7), ("tiny", 31, 8), ("tiny", 33, 9), ("tiny", 1, 10)]


def make(kind, size, seed):
    if kind == "base":
        return PS.base_mask(size, seed, n_objects=max(3, size * size // 9000))
    if kind == "tiny":
        return PS.noise_mask(size, seed, 0.6, 3) if size > 1 else np.full((1, 1), 255, np.uint8)
    if kind == "noise1":
        return PS.noise_mask(size, seed, 0.5, 1)
    return PS.noise_mask(size, seed, 0.45, 5)


@pytest.mark.parametrize("kind,size,seed", MASKS)
def test_labels_match_opencv(emul, kind, size, seed):
    m = make(kind, size, seed)
    m = m[:, :size - 5] if seed % 2 else m  # non-square, width not a multiple of 32
    for fg, conn8 in ((1, 1), (0, 0), (1, 0), (0, 1)):
        src = m if fg else 255 - m
        _, want = cv.connectedComponents(src, connectivity=8 if conn8 else 4)
        np.testing.assert_array_equal(labels(emul, m, fg, conn8), canon(want))


@pytest.mark.parametrize("kind,size,seed", MASKS)
def test_polygon_area_matches_contour_area(emul, kind, size, seed):
    m = make(kind, size, seed)
    filled, cs = post_ref.fill_and_delete(m, min_area=-1)  # hole-free components
    got = np.empty(m.shape, np.int64)
    emul.emul_area2(_p(np.ascontiguousarray(filled)), m.shape[0], m.shape[1], _p(got))
    for c in post_ref._contours(filled):
        x, y = c[0, 0]
        assert abs(got[y, x]) == round(2 * cv.contourArea(c)), (x, y)


@pytest.mark.parametrize("half", [1, 3, 10, 16])
@pytest.mark.parametrize("w", [40, 64, 333])
def test_line_morphology_matches_opencv(emul, half, w):
    rng = np.random.default_rng(half * 100 + w)
    m = cv.blur(rng.random((90, w)).astype(np.float32), (7, 3))
    m = np.where(m > 0.5, 255, 0).astype(np.uint8)
    m[:, :3] = 255  # objects touching the frame are not eroded from outside
    m[:4, :] = 255
    k = 2 * half + 1
    for vertical in (0, 1):
        ker = np.ones((k, 1) if vertical else (1, k), np.uint8)
        for erode in (0, 1):
            out = np.empty_like(m)
            emul.emul_morph(_p(m), m.shape[0], m.shape[1], half, vertical, erode, _p(out))
            want = cv.erode(m, ker) if erode else cv.dilate(m, ker)
            np.testing.assert_array_equal(out, want, err_msg=f"vertical={vertical} erode={erode}")


@pytest.mark.parametrize("kind,size,seed", MASKS[:6])
def test_cleanup_stages_and_result_match_oracle(emul, kind, size, seed):
    m = make(kind, size, seed)
    filled, _ = post_ref.fill_and_delete(m, min_area=-1)
    np.testing.assert_array_equal(cleanup(emul, m, 0), filled)                       # holes filled
    keep, _ = post_ref.fill_and_delete(m)
    np.testing.assert_array_equal(cleanup(emul, m, 1), keep)                         # polygon area <= 1000 dropped
    np.testing.assert_array_equal(cleanup(emul, m, 2), cv.erode(keep, np.ones((1, 21), np.uint8)))
    np.testing.assert_array_equal(cleanup(emul, m, 3), cv.erode(keep, np.ones((21, 1), np.uint8)))
    np.testing.assert_array_equal(cleanup(emul, m), post_ref.clean_mask(m))


@pytest.mark.parametrize("name,size,seed", PS.FUSE_CASES)
def test_fuse_matches_reference_golden(emul, name, size, seed):
    want = np.unpackbits(GOLD[name + "_fused"])[:size * size].reshape(size, size).astype(np.uint8) * 255
    masks = np.ascontiguousarray(np.stack(PS.five_masks(size, seed)))
    out = np.empty((size, size), np.uint8)
    emul.emul_fuse(_p(masks), size, size, _p(out))
    assert (out != want).sum() == 0, f"{(out != want).sum()} px differ"


@pytest.mark.parametrize("kind,size,seed", MASKS + [("noise1", 300, 11), ("base", 900, 12)])
def test_parallel_border_following_matches_findcontours(emul, kind, size, seed):
    """crack numbering + successor rule + pointer jumping + scatter == cv2.findContours(RETR_EXTERNAL, CHAIN_APPROX_NONE),
    contour order and points, on hole-free masks (what the contour stage traces) and on raw ones (holes are skipped)"""
    m0 = make(kind, size, seed)
    if seed % 2:
        m0 = np.ascontiguousarray(m0[:, :max(1, m0.shape[1] - 5)])
    filled, _ = post_ref.fill_and_delete(m0, min_area=-1)
    for m in (filled, m0):
        h, w = m.shape
        npts = np.zeros(h * w + 1, np.int32)
        xy = np.zeros(2 * (4 * h * w + 4), np.int32)
        n = emul.emul_contours(_p(np.ascontiguousarray(m)), h, w, _p(npts), npts.size, _p(xy), xy.size // 2)
        want, _ = cv.findContours(m, cv.RETR_EXTERNAL, cv.CHAIN_APPROX_NONE)
        offs = np.concatenate([[0], np.cumsum(npts[:n])])
        mine = [xy[2 * offs[i]:2 * offs[i + 1]].reshape(-1, 2) for i in range(n)]
        if m is filled:  # same contours in the same order
            assert n == len(want)
            for a, c in zip(mine, want):
                np.testing.assert_array_equal(a, c[:, 0, :])
        else:  # RETR_EXTERNAL skips islands inside holes; every contour it returns must be among ours, outer border only
            by_start = {tuple(a[0]): a for a in mine}
            assert n >= len(want)
            for c in want:
                np.testing.assert_array_equal(by_start[tuple(c[0, 0])], c[:, 0, :])
        bb = np.zeros((h, w, 4), np.int32)
        emul.emul_bboxes(_p(np.ascontiguousarray(m)), h, w, _p(bb))
        for c in want:
            x, y, bw, bh = cv.boundingRect(c)
            assert tuple(bb[c[0, 0, 1], c[0, 0, 0]]) == (x, y, x + bw, y + bh)
