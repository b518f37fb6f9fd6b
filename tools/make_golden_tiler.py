"""Generates tests/golden/tiler.npz by running the reference's own predict.detection
(/root/reference/predict.py:90-116) with a deterministic fake model, so that the device tiler /
stitcher (bd_tiles_gather + bd_stitch_or) can be checked on the GPU box where the reference is absent.

The fake model's answer depends on the pixel value AND on the position inside the tile, so the stitched
mask exercises the overlap OR, the zero padding in normalised space and the tile order."""
import os
import sys
import tempfile

import cv2 as cv
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from ref_harness import reference_modules  # noqa: E402
from fake_model import FakeModel, scene_image  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "tiler.npz")


def main():
    predict, _, _ = reference_modules()
    cases = {}
    for name, (h, w) in {"s400x300": (400, 300), "s512": (512, 512), "s600": (600, 600), "s872": (872, 872),
                         "s1000": (1000, 1000), "s130": (130, 130), "s1233": (1233, 1233)}.items():
        img = scene_image(h, w)
        with tempfile.TemporaryDirectory() as d:
            p = os.path.join(d, "in.png")
            cv.imwrite(p, img)
            model = FakeModel()
            predict.detection(p, d, model, save_name="out")
            mask = cv.imread(os.path.join(d, "out.png"), cv.IMREAD_GRAYSCALE)
        assert mask.shape == (h, w)
        cases[name + "_shape"] = np.int32([h, w])  # the image is regenerated from the seed by scene_image()
        cases[name + "_maskbits"] = np.packbits(mask > 0)
        assert set(np.unique(mask)) <= {0, 255}
        cases[name + "_ntiles"] = np.int32(model.calls)
        print(name, "tiles", model.calls, "mask on", float((mask > 0).mean()))
    np.savez_compressed(OUT, **cases)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
