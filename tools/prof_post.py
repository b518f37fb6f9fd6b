"""Fusion + contour stage timings on scene-sized synthetic masks (structured: ~N rectangles/polygons per SURVEY 8d
config 5; degenerate: what random-init networks produce).  usage: python tools/prof_post.py [size] [objects]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import post_scenes as PS  # noqa: E402
from building_detection_b200 import edge_3, model_fuse  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
nobj = int(sys.argv[2]) if len(sys.argv) > 2 else (size * size) // 20000


def clock(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return best * 1e3, r


t0 = time.time()
base = PS.base_mask(size, 5, n_objects=nobj)
rng = np.random.default_rng(6)
masks = []
for k in range(5):
    m = np.roll(base, (int(rng.integers(-2, 3)), int(rng.integers(-2, 3))), axis=(0, 1))
    masks.append(m)
print(f"built structured masks {size}^2, {nobj} objects, fill {float((base > 0).mean()):.3f} in {time.time() - t0:.1f}s", flush=True)
d = torch.from_numpy(np.stack(masks)).cuda()
ms, fused = clock(lambda: model_fuse.fuse_device(d))
print(f"structured : fuse {ms:8.2f} ms  ({19 * size * size / ms / 1e6:7.1f} GB/s algorithmic)")
ms, res = clock(lambda: edge_3.contours_device(fused))
print(f"structured : contours {ms:8.2f} ms, {len(res[0])} polygons")
full = torch.full((5, size, size), 255, dtype=torch.uint8, device="cuda")
full[:, ::97, ::89] = 0  # degenerate: almost everything set, sparse pin holes
ms, fused = clock(lambda: model_fuse.fuse_device(full))
print(f"degenerate : fuse {ms:8.2f} ms")
ms, res = clock(lambda: edge_3.contours_device(fused))
print(f"degenerate : contours {ms:8.2f} ms, {len(res[0])} polygons")
noise = torch.from_numpy(np.stack([PS.noise_mask(min(size, 4096), 30 + k, 0.5, 5) for k in range(5)])).cuda()
ms, fused = clock(lambda: model_fuse.fuse_device(noise))
print(f"noise {noise.shape[1]}^2: fuse {ms:8.2f} ms")
ms, res = clock(lambda: edge_3.contours_device(fused))
print(f"noise      : contours {ms:8.2f} ms, {len(res[0])} polygons")
