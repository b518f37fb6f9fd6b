"""N > 1 host logic on the CPU (gloo, world_size 2): tile-row band sharding + gather-to-rank-0 with OR of the
overlapping band rows must reproduce the single-process stitched masks exactly."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from building_detection_b200 import post, scene as S
from fake_model import fake_probs, scene_image

H = W = 1233  # 4 x 4 tiles -> 2 tile rows per rank at world 2; bands overlap by 152 rows


def stitched(origins, img):
    """CPU stand-in for SceneRunner.run with the fake model: OR-stitch of per-tile argmax masks."""
    out = np.zeros((2, H, W), np.uint8)
    pad = np.zeros((max(o[0] for o in S.tile_origins(H, W)) + 512, max(o[1] for o in S.tile_origins(H, W)) + 512, 3))
    pad[:H, :W] = img[:, :, ::-1] / 127.5 - 1
    for (i, j) in origins:
        m = fake_probs(pad[None, i:i + 512, j:j + 512]).argmax(-1)[0].astype(np.uint8) * 255
        for k in range(2):  # two "models": the mask and its transpose pattern
            t = m if k == 0 else m[::-1]
            out[k, i:i + 512, j:j + 512][:max(0, min(512, H - i)), :max(0, min(512, W - j))] |= t[:max(0, min(512, H - i)), :max(0, min(512, W - j))]
    return out


def fake_cleanup(mask):
    """stand-in for bd_mask_cleanup in the CPU test: any deterministic whole-mask transform (3x3 dilation)"""
    import cv2 as cv
    return torch.from_numpy(cv.dilate(mask.numpy(), np.ones((3, 3), np.uint8)))


def worker_by_model(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    img = scene_image(H, W)
    origins = S.tile_origins(H, W)
    bands = [post.band_of(S.shard_rows(origins, r, world), H) for r in range(world)]
    two = stitched(S.shard_rows(origins, rank, world), img)
    masks = torch.from_numpy(np.concatenate([two, two[:1] // 255 * 255, two[1:], two[:1]], axis=0))  # 5 "models"
    post.exchange_by_model(masks, bands, rank, world)
    mine = {k: fake_cleanup(masks[k]) for k in range(5) if post.model_owner(k, world) == rank}
    out = post.collect_cleaned(mine, 5, rank, world, masks)
    if rank == 0:
        q.put(out.numpy())
    dist.barrier()
    dist.destroy_process_group()


def worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    img = scene_image(H, W)
    origins = S.tile_origins(H, W)
    mine = S.shard_rows(origins, rank, world)
    bands = [post.band_of(S.shard_rows(origins, r, world), H) for r in range(world)]
    masks = torch.from_numpy(stitched(mine, img))
    post.gather_bands(masks, bands, rank, world)
    if rank == 0:
        q.put(masks.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_band_shard_gather_equals_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = stitched(S.tile_origins(H, W), scene_image(H, W))
    np.testing.assert_array_equal(got, want)
    bands = [post.band_of(S.shard_rows(S.tile_origins(H, W), r, 2), H) for r in range(2)]
    assert bands[0][1] - bands[1][0] == 152  # the overlap the OR has to absorb


def test_per_model_cleanup_exchange_equals_single_process():
    """multi-GPU post-processing: model k's complete mask is assembled on rank k % world, cleaned there, and the
    cleaned masks are collected on rank 0 -- must equal cleaning the single-process masks."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker_by_model, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    two = stitched(S.tile_origins(H, W), scene_image(H, W))
    full = np.concatenate([two, two[:1] // 255 * 255, two[1:], two[:1]], axis=0)
    want = np.stack([fake_cleanup(torch.from_numpy(m)).numpy() for m in full])
    np.testing.assert_array_equal(got, want)
