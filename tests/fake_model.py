"""Deterministic stand-in for a Keras model, used on both sides of the tiler parity test."""
import numpy as np


def fake_probs(x):
    """x: (N,512,512,3) float in [-1,1] RGB -> (N,512,512,2).  Class 1 where the red channel beats the blue one
    AND a position-in-tile pattern holds; zero-padded pixels (r == b == 0) are class 0."""
    x = np.asarray(x, dtype=np.float32)
    ty, tx = np.mgrid[0:512, 0:512]
    pattern = ((ty * 7 + tx * 3) % 11) < 6
    p1 = ((x[..., 0] > x[..., 2]) & pattern[None]).astype(np.float32)
    return np.stack([1.0 - p1, p1], axis=-1)


class FakeModel:
    def __init__(self):
        self.calls = 0

    def predict(self, x):
        self.calls += 1
        assert x.shape[1:] == (512, 512, 3), x.shape
        return fake_probs(x)


def scene_image(h, w):
    """Seeded BGR u8 test scene (regenerated on both sides instead of being stored in the fixture)."""
    return np.random.default_rng(1000003 * h + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
