"""Fusion (model_fuse.py) parity on the GPU, through bd_fuse / bd_mask_cleanup:
 * against the golden outputs of the reference's own model_confuse (tests/golden/post.npz, made by
   tools/make_golden_post.py in the build container);
 * against oracle/post_ref.py (the cv2-based restatement, pinned to the reference by the same script) on further
   seeded scenes, including unstructured noise masks and a scene too large for the reference to finish quickly.
Integer work: the bar is bit-exact masks."""
import os

import numpy as np
import pytest

import post_scenes as PS
from oracle import post_ref

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "post.npz"))


@pytest.mark.parametrize("name,size,seed", PS.FUSE_CASES)
def test_fuse_matches_reference_golden(gpu, name, size, seed):
    from building_detection_b200 import model_fuse
    want = np.unpackbits(GOLD[name + "_fused"])[:size * size].reshape(size, size).astype(np.uint8) * 255
    got = model_fuse.fuse(PS.five_masks(size, seed))
    assert got.shape == want.shape and set(np.unique(got)) <= {0, 255}
    assert (got != want).sum() == 0, f"{(got != want).sum()} px differ"


@pytest.mark.parametrize("size,seed", [(256, 101), (333, 102), (500, 103), (700, 104), (1024, 105)])
def test_cleanup_matches_oracle_structured(gpu, size, seed):
    import torch
    from building_detection_b200 import model_fuse
    for m in (PS.base_mask(size, seed), PS.five_masks(size, seed)[4]):
        want = post_ref.clean_mask(m)
        got = model_fuse.cleanup_device(torch.from_numpy(m).cuda()).cpu().numpy()
        assert (got != want).sum() == 0, f"{(got != want).sum()} px differ"


@pytest.mark.parametrize("size,seed,p,blur", [(256, 201, 0.5, 5), (400, 202, 0.35, 9), (400, 203, 0.6, 13),
                                              (512, 204, 0.45, 21), (300, 205, 0.5, 1)])
def test_cleanup_matches_oracle_noise(gpu, size, seed, p, blur):
    """Random-init networks produce unstructured masks; every shape pathology shows up here."""
    import torch
    from building_detection_b200 import model_fuse
    m = PS.noise_mask(size, seed, p, blur)
    want = post_ref.clean_mask(m)
    got = model_fuse.cleanup_device(torch.from_numpy(m).cuda()).cpu().numpy()
    assert (got != want).sum() == 0, f"{(got != want).sum()} of {want.size} px differ"


def _canon(lab):
    flat = lab.ravel()
    first = np.full(lab.max() + 1, -1, np.int64)
    order = np.argsort(flat, kind="stable")
    vals, start = np.unique(flat[order], return_index=True)
    first[vals] = np.arange(flat.size)[order][start]
    out = first[flat].reshape(lab.shape)
    out[lab == 0] = -1
    return out.astype(np.int32)


@pytest.mark.parametrize("size,seed,kind", [(333, 1, "base"), (700, 3, "base"), (1024, 4, "noise"), (257, 5, "noise1"),
                                            (2048, 6, "base"), (31, 8, "noise"), (1, 9, "one")])
def test_run_labels_and_stages_match_opencv(gpu, size, seed, kind):
    """the run-based labelling (4- and 8-connected, set and clear pixels) against cv2.connectedComponents, and the
    intermediate planes of a clean-up pass against the OpenCV steps of model_fuse.py -- on the GPU, where the unions
    race (the same checks run on the host-compiled kernels in tests/test_rle_emul.py)"""
    import cv2 as cv
    from building_detection_b200 import model_fuse
    if kind == "base":
        m = PS.base_mask(size, seed, n_objects=max(3, size * size // 9000))
    elif kind == "one":
        m = np.full((1, 1), 255, np.uint8)
    else:
        m = PS.noise_mask(size, seed, 0.5, 1 if kind == "noise1" else 5)
    if seed % 2:
        m = np.ascontiguousarray(m[:, :max(1, size - 5)])
    for fg, conn8 in ((1, 1), (0, 0), (1, 0), (0, 1)):
        _, want = cv.connectedComponents(m if fg else 255 - m, connectivity=8 if conn8 else 4)
        np.testing.assert_array_equal(model_fuse.debug_labels(m, fg, conn8), _canon(want))
    filled, _ = post_ref.fill_and_delete(m, min_area=-1)
    np.testing.assert_array_equal(model_fuse.debug_stage(m, 0), filled)
    keep, _ = post_ref.fill_and_delete(m)
    np.testing.assert_array_equal(model_fuse.debug_stage(m, 1), keep)
    np.testing.assert_array_equal(model_fuse.debug_stage(m, 2), cv.erode(keep, np.ones((1, 21), np.uint8)))
    np.testing.assert_array_equal(model_fuse.debug_stage(m, 3), cv.erode(keep, np.ones((21, 1), np.uint8)))
    if size <= 1024:
        np.testing.assert_array_equal(model_fuse.debug_stage(m, 7), post_ref.clean_mask(m))


def test_packed_planes_round_trip_and_fuse(gpu):
    """bd_mask_pack / bd_mask_unpack / bd_fuse_planes: what the multi-GPU gather ships; bands pack like the whole"""
    import torch
    from building_detection_b200 import model_fuse
    for size, cut in ((640, 0), (333, 4)):  # 16-byte aligned rows (vector path) and ragged ones
        masks = np.stack(PS.five_masks(size, 12))[:, :, :size - cut]
        w = masks.shape[2]
        d = torch.from_numpy(np.ascontiguousarray(masks)).cuda()
        planes = torch.stack([model_fuse.pack_device(d[k]) for k in range(5)])
        assert planes.shape == (5, size, model_fuse.plane_words(w))
        for k in range(5):
            assert torch.equal(model_fuse.unpack_device(planes[k], w), d[k])
        band = model_fuse.pack_device(d[2, 100:300])
        assert torch.equal(band, planes[2, 100:300])
        bits = np.unpackbits(planes[0].cpu().numpy().view(np.uint8), axis=1, bitorder="little")[:, :w]
        np.testing.assert_array_equal(bits * 255, masks[0])
        assert torch.equal(model_fuse.fuse_planes_device(planes, w), model_fuse.fuse_device(d))
        np.testing.assert_array_equal(model_fuse.fuse_device(d).cpu().numpy(), post_ref.model_confuse(list(masks)))


def test_fuse_edge_cases(gpu):
    from building_detection_b200 import model_fuse
    z = np.zeros((64, 80), np.uint8)
    assert model_fuse.fuse([z] * 5).sum() == 0                      # empty masks
    full = np.full((64, 80), 255, np.uint8)
    np.testing.assert_array_equal(model_fuse.fuse([full] * 5), post_ref.model_confuse([full] * 5))  # one object = the frame
    a = PS.base_mask(300, 5)
    np.testing.assert_array_equal(model_fuse.fuse([a, a, a, np.zeros_like(a), np.zeros_like(a)]),
                                  post_ref.model_confuse([a, a, a, np.zeros_like(a), np.zeros_like(a)]))  # 3 of 5
    np.testing.assert_array_equal(model_fuse.fuse([a, a, np.zeros_like(a), np.zeros_like(a), np.zeros_like(a)]),
                                  np.zeros_like(a))  # 2 of 5 -> nothing


def test_fuse_idempotent_at_scale(gpu):
    """Size-independent property at a scene the reference cannot finish in test time (4096^2, ~700 objects):
    clean-up output is hole-free and made of kept objects, so cleaning it again with the oracle's rules on a crop
    agrees, and fusing five copies of a cleaned mask returns that mask cleaned once more."""
    import torch
    from building_detection_b200 import model_fuse
    m = torch.from_numpy(PS.base_mask(4096, 7, n_objects=700)).cuda()
    c1 = model_fuse.cleanup_device(m)
    f = model_fuse.fuse_device(torch.stack([c1] * 5))
    c2 = model_fuse.cleanup_device(model_fuse.cleanup_device(c1))
    assert torch.equal(f, c2)
    crop = c1[1000:1700, 2000:2700].cpu().numpy()
    np.testing.assert_array_equal(model_fuse.cleanup_device(torch.from_numpy(crop).cuda()).cpu().numpy(),
                                  post_ref.clean_mask(crop))


# ------------------------------------------------------------------------------------------ contours (edge_3.py)
def _same(polys, want):
    assert len(polys) == len(want), (len(polys), len(want))
    for p, q in zip(polys, want):
        assert type(p[0][0]) is type(q[0][0])
        assert np.array_equal(np.asarray(p[0]), np.asarray(q[0])) and np.array_equal(np.asarray(p[1]), np.asarray(q[1]))


def _golden_polys(prefix):
    off, xs, ys, isf = (GOLD[prefix + k] for k in ("off", "xs", "ys", "isf"))
    return [[list(xs[off[i]:off[i + 1]].astype(np.float32 if isf[i] else np.int32)),
             list(ys[off[i]:off[i + 1]].astype(np.float32 if isf[i] else np.int32))] for i in range(len(off) - 1)]


@pytest.mark.parametrize("name,size,seed", PS.CONTOUR_CASES)
def test_contours_match_reference_golden(gpu, name, size, seed):
    """bit-exact polygons vs the reference's own edge_3._detection (golden)."""
    from building_detection_b200 import edge_3
    polys, h = edge_3.detect(PS.contour_case_mask(name, size, seed))
    assert h == size
    _same(polys, _golden_polys(name + "_"))


@pytest.mark.parametrize("name,size,seed", PS.FUSE_CASES)
def test_fuse_then_contours_match_reference_golden(gpu, name, size, seed):
    """the real hand-off: fused mask (device) -> polygons, both against the reference's outputs."""
    import torch
    from building_detection_b200 import edge_3, model_fuse
    masks = torch.from_numpy(np.stack(PS.five_masks(size, seed))).cuda()
    polys, _ = edge_3.contours_device(model_fuse.fuse_device(masks))
    _same(polys, _golden_polys(name + "_poly_"))


@pytest.mark.parametrize("size,seed,kind", [(300, 301, "base"), (512, 302, "base"), (900, 303, "base"), (1400, 304, "base"),
                                            (256, 305, "noise"), (384, 306, "noise"), (200, 307, "noise5")])
def test_contours_match_oracle(gpu, size, seed, kind):
    from building_detection_b200 import edge_3
    m = PS.base_mask(size, seed) if kind == "base" else PS.noise_mask(size, seed, 0.35 if kind == "noise" else 0.5, 9 if kind == "noise" else 5)
    try:
        want, _ = post_ref.detection(m)
    except IndexError:
        with pytest.raises(IndexError):
            edge_3.detect(m)
        return
    polys, _ = edge_3.detect(m)
    _same(polys, want)


def test_contours_edge_cases(gpu):
    from building_detection_b200 import edge_3
    assert edge_3.detect(np.zeros((40, 50), np.uint8)) == ([], 40)
    one = np.zeros((64, 64), np.uint8)
    one[10:40, 12:50] = 255
    _same(edge_3.detect(one)[0], post_ref.detection(one)[0])
    full = np.full((64, 64), 255, np.uint8)
    _same(edge_3.detect(full)[0], post_ref.detection(full)[0])
    # two squares joined only through a corner: the 7-px erosion splits them (edge_3.py:159-262)
    m = np.zeros((200, 200), np.uint8)
    m[20:80, 20:80] = 255
    m[79:140, 79:140] = 255
    _same(edge_3.detect(m)[0], post_ref.detection(m)[0])
