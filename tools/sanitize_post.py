"""Small fusion + contour workload for compute-sanitizer (memcheck / initcheck / racecheck / synccheck over the
bit-plane and run-based kernels of csrc/rle.cuh, post.cu, contours.cu) -- structured, noise, all-foreground and empty
masks at sizes that are not multiples of 32, results checked against the oracle as in the parity tests.
usage: compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_post.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import post_scenes as PS  # noqa: E402
from building_detection_b200 import edge_3, model_fuse  # noqa: E402
from oracle import post_ref  # noqa: E402

cases = [np.stack(PS.five_masks(333, 3))[:, :, :301], np.stack([PS.noise_mask(257, 10 + k, 0.5, 5) for k in range(5)]),
         np.full((5, 97, 130), 255, np.uint8), np.zeros((5, 64, 64), np.uint8)]
for masks in cases:
    masks = np.ascontiguousarray(masks)
    fused = model_fuse.fuse(list(masks))
    assert np.array_equal(fused, post_ref.model_confuse(list(masks)))
    try:
        want = post_ref.detection(fused)[0]
    except IndexError:
        want = None
    try:
        got = edge_3.detect(fused)[0]
    except IndexError:
        got = None
    assert (want is None) == (got is None) and (want is None or len(want) == len(got))
    print(masks.shape, "fused on", float((fused > 0).mean()), "polygons", None if got is None else len(got), flush=True)
print("sanitize_post: ok")
