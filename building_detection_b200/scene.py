"""Device-resident scene pipeline: the reference's tiler (predict.py:90-116) driven through the C ABI.

A scene (H,W,3 BGR u8) is uploaded once; tiles are gathered on the device (normalise + zero pad fused
into the gather), pushed through a model's native plan in batches, and the per-tile argmax masks are
OR-stitched into a u8 scene mask that never leaves the GPU until fusion / contour extraction are done.
torch is used only for device memory and streams.
"""
from __future__ import annotations

import math

import numpy as np

from . import runtime as R

TILE, STRIDE, OVERLAP = 512, 360, 152  # predict.py:98-107


def tile_origins(h, w, bug_compatible=True):
    """Top-left corners of the 512x512 tiles in the reference's order (predict.py:98-106).

    The reference iterates the *column* loop over ``new_h`` as well; with
    ``bug_compatible`` that is reproduced, and a scene for which it would slice a tile narrower than
    512 px (Keras rejects the shape) raises ValueError.  ``bug_compatible=False`` tiles the columns
    over ``new_w``."""
    h_num = math.ceil((h - OVERLAP) / STRIDE)
    w_num = math.ceil((w - OVERLAP) / STRIDE)
    new_h, new_w = h_num * STRIDE + OVERLAP, w_num * STRIDE + OVERLAP
    pad_h, pad_w = max(new_h, TILE), max(new_w, TILE)
    col_extent = new_h if bug_compatible else new_w
    out = []
    for i in range(0, new_h - OVERLAP, STRIDE):
        for j in range(0, col_extent - OVERLAP, STRIDE):
            if i + TILE > pad_h or j + TILE > pad_w:
                raise ValueError(f"tile at ({i},{j}) leaves the padded {pad_h}x{pad_w} scene: the reference slices a "
                                 "short tile there and model.predict rejects it (predict.py:106 iterates columns "
                                 "over new_h); pass bug_compatible=False for non-square scenes")
            out.append((i, j))
    return out


def shard_rows(origins, rank, world):
    """Contiguous row-band shard of the tile list for ``rank`` of ``world``: whole
    tile rows stay on one GPU so that each rank's writes cover a contiguous band of the scene mask."""
    rows = sorted({i for i, _ in origins})
    per = math.ceil(len(rows) / world) if rows else 0
    mine = set(rows[rank * per:(rank + 1) * per])
    return [o for o in origins if o[0] in mine]


BATCH_CANDIDATES = (32, 28, 24, 20, 16)


def best_batch(n_tiles, max_batch=32):
    """Tiles per plan launch for a shard of ``n_tiles``: the candidate that pads the shard the least (a ragged last
    batch runs a whole plan), the largest on ties -- 3136 tiles -> 32 (98 launches), 392 tiles (one of 8 GPUs on a
    20 000^2 scene) -> 28 (14 launches, no padding; 32 would pad 392 to 416).  Fewer than 16 tiles: the next power of
    two (engine.Model.plan_batch_for)."""
    cands = [b for b in BATCH_CANDIDATES if b <= max_batch]
    if n_tiles < min(cands):
        b = 1
        while b < n_tiles:
            b *= 2
        return b
    return min(cands, key=lambda b: (-(-n_tiles // b) * b, -b))


class SceneRunner:
    """Runs models over one scene on one GPU.  ``models``: engine.Model objects.  ``batch``: tiles per plan launch, or
    None to choose per scene with ``best_batch``."""

    def __init__(self, models, batch=None, device=None):
        import torch
        if not torch.cuda.is_available():
            raise R.NativeError("building_detection_b200 needs a CUDA device (B200); there is no CPU path")
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.models = list(models)
        self.batch = batch
        self.ctx = R.context(self.device.index)
        self.lib = R.lib()

    def upload(self, scene_bgr):
        """numpy (H,W,3) u8 BGR (what cv.imread returns) -> device tensor."""
        t = self.torch
        a = np.ascontiguousarray(scene_bgr, dtype=np.uint8)
        if a.ndim != 3 or a.shape[2] != 3:
            raise ValueError(f"expected (H,W,3) uint8 BGR, got {a.shape}")
        return t.from_numpy(a).to(self.device, non_blocking=True)

    def run_average(self, scene_dev, origins=None, bug_compatible=True):
        """Opt-in fusion by probability averaging (bd_scene_run_average): ONE (H,W) u8 mask, 255 where the mean over the
        models of P(building) exceeds 0.5 in any covering tile.  Not the reference's behaviour (it votes 3 of 5 on
        argmax masks, model_fuse.py:315-324); the caller applies the final clean-up and the contour stage."""
        t = self.torch
        h, w = int(scene_dev.shape[0]), int(scene_dev.shape[1])
        if origins is None:
            origins = tile_origins(h, w, bug_compatible)
        out = t.zeros((h, w), dtype=t.uint8, device=self.device)
        if not len(origins):
            return out
        ys = np.ascontiguousarray([o[0] for o in origins], np.int32)
        xs = np.ascontiguousarray([o[1] for o in origins], np.int32)
        pb = best_batch(len(origins)) if self.batch is None else (
            self.batch if len(origins) >= self.batch else min(self.batch, self.models[0].plan_batch_for(len(origins))))
        self.last_batch = pb
        plans = [m.native_plan(pb, device=self.device.index) for m in self.models]
        handles = (R.C.c_void_p * len(plans))(*[p.h for p in plans])
        R.check(self.lib.bd_scene_run_average(self.ctx, handles, len(plans), scene_dev.data_ptr(), h, w, R._ptr(ys), R._ptr(xs),
                                              len(origins), out.data_ptr(), t.cuda.current_stream(self.device).cuda_stream))
        return out

    def run(self, scene_dev, origins=None, out=None, bug_compatible=True):
        """scene_dev: (H,W,3) u8 cuda tensor.  Returns masks (len(models),H,W) u8 {0,255} on the device
        (predict.py:113-114).  ``origins`` restricts the work to a shard of the tile list."""
        t = self.torch
        h, w = int(scene_dev.shape[0]), int(scene_dev.shape[1])
        if origins is None:
            origins = tile_origins(h, w, bug_compatible)
        if out is None:
            out = t.zeros((len(self.models), h, w), dtype=t.uint8, device=self.device)
        stream = t.cuda.current_stream(self.device).cuda_stream
        L, C = self.lib, R.C
        if not len(origins):
            return out
        # One call for the whole scene (bd_scene_run: the batch loop, CUDA-graph replay of every plan).  A ragged last
        # batch runs through the SAME batch-sized plan; scenes with fewer tiles than the batch use the next power of two
        # (engine.Model.plan_batch_for), so at most five plan sizes per model ever exist.
        ys = np.ascontiguousarray([o[0] for o in origins], np.int32)
        xs = np.ascontiguousarray([o[1] for o in origins], np.int32)
        if self.batch is None:
            pb = best_batch(len(origins))
        else:
            pb = self.batch if len(origins) >= self.batch else min(self.batch, self.models[0].plan_batch_for(len(origins)))
        self.last_batch = pb
        plans = [m.native_plan(pb, device=self.device.index) for m in self.models]
        handles = (C.c_void_p * len(plans))(*[p.h for p in plans])
        R.check(L.bd_scene_run(self.ctx, handles, len(plans), scene_dev.data_ptr(), h, w, R._ptr(ys), R._ptr(xs),
                               len(origins), out.data_ptr(), stream))
        return out
