"""The literals of the reference's post-processing in one place, each with the reference line
it comes from.  The CUDA / C++ side holds the same values in csrc/post_ws.cuh (PostConstants); ``from_library()``
reads them back through the C ABI so that a test can assert the two never drift apart."""
from __future__ import annotations

from dataclasses import dataclass, fields


@dataclass(frozen=True)
class PostConstants:
    # model_fuse.py
    fuse_min_area: int = 1000        # :22   fill_and_delete erases polygons of area <= 1000
    fuse_min_fragment: int = 500     # :57   fill_small_target erases fragments of area <= 500
    fuse_split_width: int = 21       # :180-181 (erode_process(img, 5, 5)), :68/:93   1x5 / 5x1 kernel, 5 iterations
    fuse_votes: int = 3              # :323  sum of the five masks >= 3
    # edge_3.py
    edge_min_area: int = 100         # :326
    edge_min_fragment: int = 50      # :126  (strict: < 50 erased)
    edge_split_width: int = 7        # :172-199 (1x7 / 7x1 kernels, one iteration)
    edge_iou: float = 0.5            # :42
    edge_min_moment: float = 10.0    # :360  m00 <= 10 skipped
    tier_small: float = 150.0        # :364  area < 150 -> small_target
    tier_mid: float = 300.0          # :366  150 < area < 300 -> 5 x epsilon
    tier_big0: float = 3000.0        # :370
    tier_big1: float = 8000.0        # :372
    tier_big2: float = 15000.0       # :374
    eps_default: float = 0.01        # :357  epsilon = 0.01 x perimeter
    eps_mid_mult: float = 5.0        # :367
    eps_big0: float = 0.005          # :290
    eps_big1: float = 0.004          # :297
    eps_big2: float = 0.002          # :304

    @classmethod
    def from_library(cls, device=None):
        import ctypes as C

        from . import runtime as R
        k = R.PostConstants()
        R.check(R.lib().bd_post_constants(R.context(device), C.byref(k)))
        return cls(**{f.name: getattr(k, f.name) for f in fields(cls)})


# tiler (predict.py:98-107)
TILE, STRIDE, OVERLAP = 512, 360, 152
DEFAULT = PostConstants()
