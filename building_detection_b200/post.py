"""Whole-scene job: tiles -> five stitched masks (sharded over GPUs) -> gather to rank 0 -> 3-of-5 fusion
-> contours.  This is run_model + model_confuse + _detection of the reference (predict.py:148-152) as one
device-resident pipeline; bench.py and predict.predict() drive it.

Multi-GPU: every rank owns a contiguous band of tile rows and all five models' weights;
no collective per tile.  Band masks (5 x rows x W u8) are sent to rank 0 over NCCL and OR-ed into its scene
masks (adjacent bands overlap by 152 rows; OR is idempotent), then fusion and contour extraction -- which
need whole connected components -- run on rank 0.
"""
from __future__ import annotations

from . import scene as S


def band_of(origins, h):
    """Rows [r0, r1) of the scene that a shard of tile origins can write."""
    if not origins:
        return (0, 0)
    return (min(o[0] for o in origins), min(h, max(o[0] for o in origins) + S.TILE))


def gather_bands(masks, bands, rank, world, stage=None):
    """Band masks -> rank 0 (send/recv on the default process group: NCCL on GPUs, gloo in the CPU tests), OR-ed
    into rank 0's scene masks.  ``masks``: (M,H,W) u8 tensor on every rank; ``bands[r]`` = rows rank r owns."""
    if world == 1:
        return masks
    import torch
    import torch.distributed as dist
    if rank == 0:
        for r in range(1, world):
            r0, r1 = bands[r]
            if r1 <= r0:
                continue
            if stage is not None and stage.shape[1] >= r1 - r0:
                buf = stage[:, :r1 - r0]
                buf = buf if buf.is_contiguous() else torch.empty_like(masks[:, r0:r1])
            else:
                buf = torch.empty_like(masks[:, r0:r1])
            dist.recv(buf, src=r)
            masks[:, r0:r1].bitwise_or_(buf)
    else:
        r0, r1 = bands[rank]
        if r1 > r0:
            dist.send(masks[:, r0:r1].contiguous(), dst=0)
    return masks


def model_owner(k, world):
    """Rank that cleans model k's scene mask: the five per-model clean-ups of model_fuse.py:285-313 are independent,
    so with several GPUs they run on different ranks."""
    return k % world


def exchange_by_model(masks, bands, rank, world):
    """Every rank holds its band of all M model masks; afterwards the owner of model k holds the complete mask k
    (band rows OR-ed in place into its own ``masks[k]``).  Plain send/recv, every rank walks k in the same order and
    for a given k there is a single receiver, so there is no circular wait."""
    if world == 1:
        return masks
    import torch
    import torch.distributed as dist
    for k in range(masks.shape[0]):
        owner = model_owner(k, world)
        if rank == owner:
            for r in range(world):
                r0, r1 = bands[r]
                if r == rank or r1 <= r0:
                    continue
                buf = torch.empty_like(masks[k, r0:r1])
                dist.recv(buf, src=r)
                masks[k, r0:r1].bitwise_or_(buf)
        else:
            r0, r1 = bands[rank]
            if r1 > r0:
                dist.send(masks[k, r0:r1].contiguous(), dst=owner)
    return masks


def collect_cleaned(cleaned, n_models, rank, world, like):
    """cleaned: {k: (H,W) u8 tensor} for the models this rank owns -> (M,H,W) on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist
    if rank == 0:
        out = torch.empty((n_models,) + tuple(like.shape[-2:]), dtype=torch.uint8, device=like.device)
        for k in range(n_models):
            owner = model_owner(k, world)
            if owner == 0:
                out[k].copy_(cleaned[k])
            else:
                dist.recv(out[k], src=owner)
        return out
    for k in sorted(cleaned):
        dist.send(cleaned[k].contiguous(), dst=0)
    return None


class SceneJob:
    def __init__(self, runner, h, w, origins, rank=0, world=1, do_post=True):
        import torch
        self.t = torch
        self.runner, self.h, self.w = runner, h, w
        self.origins, self.rank, self.world, self.do_post = origins, rank, world, do_post
        dev = runner.device
        self.masks = torch.zeros((len(runner.models), h, w), dtype=torch.uint8, device=dev)
        self.all_origins = S.tile_origins(h, w)
        self.bands = [band_of(S.shard_rows(self.all_origins, r, world), h) for r in range(world)]
        self.scene_dev = None
        self.stage = None
        if world > 1 and rank == 0 and not do_post:
            rows = max(b[1] - b[0] for b in self.bands[1:])
            self.stage = torch.empty((len(runner.models), rows, w), dtype=torch.uint8, device=dev)
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.host_mask = None
        self.result = None

    # ------------------------------------------------------------------ stages
    def _forward(self, scene_dev):
        self.masks.zero_()
        self.runner.run(scene_dev, origins=self.origins, out=self.masks)

    def _gather(self):
        """N > 1 with post-processing: model k's complete mask goes to its owner rank (the five clean-ups run in
        parallel on different GPUs); without post-processing all bands go to rank 0."""
        if self.do_post:
            exchange_by_model(self.masks, self.bands, self.rank, self.world)
        else:
            gather_bands(self.masks, self.bands, self.rank, self.world, self.stage)

    def _post(self):
        if not self.do_post:
            return None
        from . import edge_3, model_fuse
        if self.world == 1:
            fused = model_fuse.fuse_device(self.masks)
        else:
            mine = {k: model_fuse.cleanup_device(self.masks[k]) for k in range(self.masks.shape[0])
                    if model_owner(k, self.world) == self.rank}
            cleaned = collect_cleaned(mine, self.masks.shape[0], self.rank, self.world, self.masks)
            if self.rank != 0:
                return None
            fused = model_fuse.fuse_cleaned_device(cleaned)
        polys = edge_3.contours_device(fused)
        return fused, polys

    # ------------------------------------------------------------------ entry points
    def run_resident(self, scene_dev):
        """Scene already in HBM.  Returns (fused mask tensor, polygons) on rank 0 (None elsewhere / without post)."""
        self._forward(scene_dev)
        self._gather()
        self.result = self._post()
        return self.result

    def run_e2e(self, scene_host):
        """Scene in pinned host memory: every rank uploads the rows its band needs, rank 0 reads the fused mask
        (or, without post-processing, the five stitched masks) back to the host."""
        t = self.t
        dev = self.runner.device
        if self.scene_dev is None:
            self.scene_dev = t.empty((self.h, self.w, 3), dtype=t.uint8, device=dev)
        r0, r1 = self.bands[self.rank]
        self.scene_dev[r0:r1].copy_(scene_host[r0:r1], non_blocking=True)
        self.h2d_bytes = (r1 - r0) * self.w * 3
        res = self.run_resident(self.scene_dev)
        if self.rank == 0:
            src = res[0] if res is not None else self.masks
            if self.host_mask is None or self.host_mask.shape != src.shape:
                self.host_mask = t.empty(src.shape, dtype=t.uint8).pin_memory()
            self.host_mask.copy_(src, non_blocking=True)
            self.d2h_bytes = src.numel()
            if res is not None:
                self.d2h_bytes += res[1].nbytes if hasattr(res[1], "nbytes") else 0
        t.cuda.current_stream(dev).synchronize()
        return self.host_mask, (res[1] if res is not None else None)
