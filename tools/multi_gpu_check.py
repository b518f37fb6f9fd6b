"""N-GPU bit-identity check on hardware (launched by tests/test_multi_gpu_hw.py under torchrun, one rank per GPU):
post.SceneJob with the tile rows sharded over the ranks must give exactly the masks, the fused mask and the polygons
of the single-GPU job on the same scene.  Rank 0 also runs the world-1 job on its own GPU.
usage: torchrun --nproc-per-node N tools/multi_gpu_check.py [scene_edge]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from building_detection_b200 import post, scene as S  # noqa: E402
from building_detection_b200.predict_model import CTORS, MODEL_NAMES  # noqa: E402
from oracle import nets  # noqa: E402


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 2312  # 6 x 6 tiles
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from test_scene_gpu import blob_scene
    img = blob_scene(size, size, 1234)
    models = []
    for n in MODEL_NAMES:  # parity weights except HRNet's calibration-free variant would cost a CPU forward per rank:
        m = CTORS[n]()     # He-scaled scse, Keras-default + randomised BN for the rest (identical on every rank)
        if n == "scse":
            m.set_weights(nets.he_scaled_weights(m.spec, 2))
        models.append(m)
    runner = S.SceneRunner(models, batch=16, device=local)
    origins = S.tile_origins(size, size)
    scene = runner.upload(img)
    ok = True
    for do_post in (False, True):
        job = post.SceneJob(runner, size, size, S.shard_rows(origins, rank, world), rank, world, do_post=do_post)
        res = job.run_resident(scene)
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            one = post.SceneJob(runner, size, size, origins, 0, 1, do_post=do_post)
            ref = one.run_resident(scene)
            torch.cuda.synchronize()
            if not do_post:
                same = torch.equal(job.masks, one.masks)
                print(f"[world {world}] stitched masks identical to world 1: {same}; class-1 fractions "
                      f"{[round(float((m > 0).float().mean()), 3) for m in one.masks]}", flush=True)
                ok &= same
            else:
                same_f = torch.equal(res[0], ref[0])
                pa, pb = res[1][0], ref[1][0]
                same_p = len(pa) == len(pb) and all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) for a, b in zip(pa, pb))
                print(f"[world {world}] fused mask identical: {same_f} (on fraction {float((ref[0] > 0).float().mean()):.3f}); "
                      f"{len(pb)} polygons identical: {same_p}", flush=True)
                ok &= same_f and same_p
        dist.barrier()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL", flush=True)
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
