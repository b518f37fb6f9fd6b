"""Tiler / stitcher parity against golden vectors made by the reference's own predict.detection
(tools/make_golden_tiler.py -> tests/golden/tiler.npz)."""
import os

import numpy as np
import pytest

from building_detection_b200 import scene as S
from fake_model import fake_probs, scene_image

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "tiler.npz"))
CASES = sorted(k[:-6] for k in GOLD.files if k.endswith("_shape"))


def golden(case):
    h, w = (int(v) for v in GOLD[case + "_shape"])
    mask = np.unpackbits(GOLD[case + "_maskbits"])[:h * w].reshape(h, w) * 255
    return h, w, mask.astype(np.uint8), int(GOLD[case + "_ntiles"])


@pytest.mark.parametrize("case", CASES)
def test_tile_origins_count_matches_reference(case):
    h, w, _, ntiles = golden(case)
    assert len(S.tile_origins(h, w)) == ntiles


def test_tile_origins_geometry():
    o = S.tile_origins(20000, 20000)
    assert len(o) == 56 * 56 and o[0] == (0, 0) and o[1] == (0, 360) and o[-1] == (19800, 19800)
    with pytest.raises(ValueError):  # predict.py:106 column loop over new_h
        S.tile_origins(1233, 1232)
    assert len(S.tile_origins(1233, 1232, bug_compatible=False)) == 12
    assert S.tile_origins(130, 130) == []


def test_best_batch_pads_least():
    assert S.best_batch(3136) == 32 and S.best_batch(1568) == 32      # 1 and 2 GPUs on the 20 000^2 scene: no padding
    assert S.best_batch(784) == 28 and S.best_batch(392) == 28        # 4 and 8 GPUs: 28 divides, 32 would pad
    assert S.best_batch(36) == 20 and S.best_batch(25) == 28 and S.best_batch(16) == 16
    assert S.best_batch(9) == 16 and S.best_batch(4) == 4 and S.best_batch(1) == 1 and S.best_batch(3) == 4


def test_row_band_shards_partition_the_tiles():
    o = S.tile_origins(20000, 20000)
    for world in (1, 2, 4, 8):
        parts = [S.shard_rows(o, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == sorted(o)
        assert all(len(p) == len(o) // world for p in parts)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_device_gather_and_stitch_match_reference(gpu, case):
    import torch
    from building_detection_b200 import runtime as R
    h, w, want, _ = golden(case)
    img = scene_image(h, w)
    origins = S.tile_origins(h, w)
    L, ctx = R.lib(), R.context(0)
    scene = torch.from_numpy(img).cuda()
    out = torch.zeros((h, w), dtype=torch.uint8, device="cuda")
    from oracle import plan_interp
    for b0 in range(0, len(origins), 5):
        chunk = origins[b0:b0 + 5]
        n = len(chunk)
        ys = np.asarray([c[0] for c in chunk], np.int32)
        xs = np.asarray([c[1] for c in chunk], np.int32)
        # what predict.py:91-104 feeds the model: BGR->RGB, /127.5-1 in float64, zero pad, float32 cast
        pad = np.zeros((n, 512, 512, 3))
        for k, (i, j) in enumerate(chunk):
            sub = img[i:i + 512, j:j + 512, ::-1] / 127.5 - 1
            pad[k, :sub.shape[0], :sub.shape[1]] = sub
        xh = pad.astype(np.float32)
        for stride in (1, 2):  # plan input layout: the tile im2col'ed for the 3x3 stem, 255*x exact in fp16
            o = 512 // stride
            x = torch.empty((n, o, o, 32), dtype=torch.float16, device="cuda")
            R.check(L.bd_tiles_gather(ctx, scene.data_ptr(), h, w, R._ptr(ys), R._ptr(xs), n, x.data_ptr(), stride, None))
            torch.cuda.synchronize()
            want_x = plan_interp.im2col_input(xh, stride, emulate_h16=False).numpy()
            np.testing.assert_array_equal(x.cpu().numpy().astype(np.float32), np.round(want_x))
            assert np.abs(want_x - np.round(want_x)).max() < 1e-4  # 255*x is an integer up to float32 rounding
        tile_mask = torch.from_numpy(fake_probs(xh).argmax(-1).astype(np.uint8)).cuda()
        R.check(L.bd_stitch_or(ctx, tile_mask.data_ptr(), R._ptr(ys), R._ptr(xs), n, out.data_ptr(), h, w, None))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy(), want)
