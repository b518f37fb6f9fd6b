"""Whole-scene job: tiles -> five stitched masks (sharded over GPUs) -> gather to rank 0 -> 3-of-5 fusion
-> contours.  This is run_model + model_confuse + _detection of the reference (predict.py:148-152) as one
device-resident pipeline; bench.py and predict.predict() drive it.

Multi-GPU: every rank owns a contiguous band of tile rows and all five models' weights;
no collective per tile.  Every rank packs its band of the five masks to 1 bit per pixel (bd_mask_pack: rows are
independent) and sends it to rank 0 in ONE message (grouped NCCL send/recv, 1/8 of the u8 bytes); rank 0 ORs the
bands into five scene planes (adjacent bands overlap by 152 rows; OR is idempotent) and runs fusion and contour
extraction -- which need whole connected components -- on them (bd_fuse_planes, bd_contours: ~20 ms at 20 000^2,
so sharding them further would cost more in exchange than it saves).
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


def gather_packed(band_planes, bands, rank, world, h):
    """band_planes: (M, rows, wp) int32 -- this rank's band of the M masks, bit-packed -> (M, h, wp) planes of the
    whole scene on rank 0 (None elsewhere).  One message per rank, posted as one group."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return band_planes
    m, _, wp = band_planes.shape
    if rank != 0:
        if band_planes.shape[1]:
            for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, band_planes.contiguous(), 0)]):
                req.wait()
        return None
    planes = torch.zeros((m, h, wp), dtype=band_planes.dtype, device=band_planes.device)
    r0, r1 = bands[0]
    planes[:, r0:r1] = band_planes
    stage, ops = {}, []
    for r in range(1, world):
        r0, r1 = bands[r]
        if r1 > r0:
            stage[r] = torch.empty((m, r1 - r0, wp), dtype=band_planes.dtype, device=band_planes.device)
            ops.append(dist.P2POp(dist.irecv, stage[r], r))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for r, buf in stage.items():
        r0, r1 = bands[r]
        planes[:, r0:r1].bitwise_or_(buf)
    return planes


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
        self.planes = None
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
        """N > 1 with post-processing: every rank's bit-packed band goes to rank 0 in one message; without
        post-processing the u8 bands go to rank 0."""
        self.planes = None
        if self.world == 1:
            return
        if self.do_post:
            from . import model_fuse
            r0, r1 = self.bands[self.rank]
            band = self.t.stack([model_fuse.pack_device(self.masks[k, r0:r1]) for k in range(self.masks.shape[0])]) \
                if r1 > r0 else self.t.empty((self.masks.shape[0], 0, model_fuse.plane_words(self.w)), dtype=self.t.int32,
                                             device=self.masks.device)
            self.planes = gather_packed(band, self.bands, self.rank, self.world, self.h)
        else:
            gather_bands(self.masks, self.bands, self.rank, self.world, self.stage)

    def _post(self):
        if not self.do_post:
            return None
        from . import edge_3, model_fuse
        if self.world == 1:
            fused = model_fuse.fuse_device(self.masks)
        else:
            if self.rank != 0:
                return None
            fused = model_fuse.fuse_planes_device(self.planes, self.w)
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
