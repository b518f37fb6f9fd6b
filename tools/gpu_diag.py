"""GPU bring-up diagnostics (run on the B200 box): per-op and per-model comparisons of the native
kernels against the CPU plan interpreter (oracle/plan_interp.py), printed as a table.  Not a test --
tests/ holds the pass/fail versions -- this keeps going after a mismatch so that one GPU trip tells as
much as possible.   usage: python tools/gpu_diag.py <group> [...]
groups: direct, umma_basic, umma_more, ops, models_direct, models_umma, perf"""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from building_detection_b200 import graph as G  # noqa: E402
from util import build_two_pass, rand_map, run_interp, run_native, h16_ulp  # noqa: E402


if os.environ.get("BD_DRY"):  # CPU dry run of the harness itself: "native" = a second interpreter run
    class _Fake:
        def __init__(self, plan, inputs):
            self.it = run_interp(plan, inputs)
        def read_buffer(self, b):
            return self.it.get(b)
        def close(self):
            pass
    def run_native(plan, inputs):  # noqa: F811
        return _Fake(plan, inputs)


def report(tag, got, ref):
    err = np.abs(got - ref)
    ulp = h16_ulp(ref)
    rel = (err / np.maximum(ulp, 1e-6)).max()
    bad = int((err > 2 * ulp + 1e-5).sum())
    status = "OK " if bad == 0 else "BAD"
    print(f"[{status}] {tag:58s} max|d|={err.max():.3e} max_ulp={rel:7.2f} bad={bad}/{err.size} ref_rms={np.sqrt((ref**2).mean()):.3f}",
          flush=True)
    return bad == 0


def conv_case(tag, N, H, W, Cin, Cout, k=3, s=1, d=1, res=False, res_after_act=False, act="relu", umma=True, bn=True,
              in_slice=None, out_slice=None, seed=0):
    def builder(g):
        if in_slice:
            xb = g.buf(H, W, in_slice[1])
            x = G.T(xb, in_slice[0], Cin)
        else:
            x = g.new(H, W, Cin)
        Ho, Wo = -(-H // s), -(-W // s)
        r = g.new(Ho, Wo, Cout) if res else None
        out = None
        if out_slice:
            ob = g.buf(Ho, Wo, out_slice[1])
            out = G.T(ob, out_slice[0], Cout)
        y = g.conv(x, "c", Cout, k=k, s=s, d=d, bn=bn, act=act, res=r, res_after_act=res_after_act, out=out)
        return x, r, y

    try:
        plan, (x, r, y), _ = build_two_pass(builder, N, seed=seed, umma=umma)
        path = plan.ops[0]["path"]
        rng = np.random.default_rng(seed + 1)
        inputs = {x.buf.id: rand_map(rng, plan, x.buf.id)}
        if r is not None:
            inputs[r.buf.id] = rand_map(rng, plan, r.buf.id)
        ref = run_interp(plan, inputs).get(y.buf.id)
        nat = run_native(plan, inputs)
        got = nat.read_buffer(y.buf.id)
        nat.close()
        return report(f"{tag} [{path}]", got[..., y.c0:y.c0 + y.C], ref[..., y.c0:y.c0 + y.C])
    except Exception as e:  # keep going
        print(f"[ERR] {tag}: {type(e).__name__}: {e}", flush=True)
        traceback.print_exc()
        return False


def convT_case(tag, N, H, W, Cin, Cout, k, act, umma=True):
    def builder(g):
        x = g.new(H, W, Cin)
        y = g.conv_transpose(x, "t", Cout, k, act=act)
        return x, y

    try:
        plan, (x, y), _ = build_two_pass(builder, N, umma=umma)
        rng = np.random.default_rng(5)
        inputs = {x.buf.id: rand_map(rng, plan, x.buf.id)}
        ref = run_interp(plan, inputs).get(y.buf.id)
        nat = run_native(plan, inputs)
        got = nat.read_buffer(y.buf.id)
        nat.close()
        return report(f"{tag} [{plan.ops[0]['path']}]", got, ref)
    except Exception as e:
        print(f"[ERR] {tag}: {type(e).__name__}: {e}", flush=True)
        traceback.print_exc()
        return False


def group_direct():
    conv_case("direct 1x1 16ch", 1, 16, 16, 16, 16, k=1, umma=False)
    conv_case("direct 3x3 cin3 f32->64", 1, 32, 32, 8, 64, k=3, umma=False)
    conv_case("direct 3x3 s2", 2, 32, 32, 16, 24, k=3, s=2, umma=False)
    conv_case("direct 3x3 d4 c4", 1, 32, 32, 4, 4, k=3, d=4, umma=False)
    conv_case("direct 3x3 res after act", 1, 16, 16, 16, 16, k=3, res=True, res_after_act=True, umma=False)
    convT_case("direct convT k3", 1, 8, 8, 16, 8, 3, "relu", umma=False)
    convT_case("direct convT k2", 1, 8, 8, 16, 8, 2, None, umma=False)


def group_umma_basic():
    conv_case("umma 1x1 64->64 32x32", 1, 32, 32, 64, 64, k=1, act=None, bn=False)
    conv_case("umma 1x1 64->64 32x32 bn relu", 1, 32, 32, 64, 64, k=1)
    conv_case("umma 3x3 64->64 32x32", 1, 32, 32, 64, 64, k=3)
    conv_case("umma 1x1 128->128 (2 k-chunks)", 1, 32, 32, 128, 128, k=1)
    conv_case("umma 3x3 256->256", 1, 32, 32, 256, 256, k=3)
    conv_case("umma 3x3 32->32 (K OOB fill)", 1, 32, 32, 32, 32, k=3)
    conv_case("umma 1x1 64->16", 1, 32, 32, 64, 16, k=1)
    conv_case("umma 3x3 64->64 batch 3 64x64", 3, 64, 64, 64, 64, k=3)


def group_umma_more():
    conv_case("umma 1x1 728->728 (N tail, K tail)", 2, 32, 32, 728, 728, k=1)
    conv_case("umma 1x1 1536->2048", 1, 32, 32, 1536, 2048, k=1)
    conv_case("umma 3x3 d6 256->256", 1, 32, 32, 256, 256, k=3, d=6)
    conv_case("umma 3x3 d12 256->256", 1, 32, 32, 256, 256, k=3, d=12)
    conv_case("umma 3x3 d18 2048->256", 1, 32, 32, 2048, 256, k=3, d=18)
    conv_case("umma 3x3 s2 64->128", 2, 64, 64, 64, 128, k=3, s=2)
    conv_case("umma 1x1 s2 64->128", 2, 64, 64, 64, 128, k=1, s=2, act=None, bn=False)
    conv_case("umma 3x3 res (act after add)", 1, 32, 32, 64, 64, k=3, res=True)
    conv_case("umma 3x3 res_after_act", 1, 32, 32, 64, 64, k=3, res=True, res_after_act=True)
    conv_case("umma 3x3 res no act", 1, 32, 32, 64, 64, k=3, res=True, act=None)
    conv_case("umma 3x3 in/out slices", 1, 32, 32, 64, 64, k=3, in_slice=(64, 192), out_slice=(32, 128))
    conv_case("umma 3x3 16x16 batch 2", 2, 16, 16, 64, 64, k=3)
    conv_case("umma 3x3 8x8 batch 4", 4, 8, 8, 64, 64, k=3)
    conv_case("umma 3x3 40x24 (ragged tiles)", 1, 40, 24, 64, 64, k=3)
    conv_case("umma 3x3 384->384 64x64", 1, 64, 64, 384, 384, k=3)
    conv_case("umma 3x3 640->640 64x64", 1, 64, 64, 640, 640, k=3)
    conv_case("umma 3x3 64->64 256x256", 1, 256, 256, 64, 64, k=3)
    convT_case("umma convT k3 128->64", 1, 32, 32, 128, 64, 3, "relu")
    convT_case("umma convT k2 128->64", 1, 32, 32, 128, 64, 2, "relu")


def group_ops():
    from oracle import plan_interp  # noqa: F401

    def run(tag, builder, n=2, outs=None, seed=0):
        try:
            plan, res, _ = build_two_pass(builder, n, seed=seed)
            ins, outs_t = res
            rng = np.random.default_rng(seed + 7)
            inputs = {}
            for t in ins:
                bid = t.buf.id
                if bid not in inputs:
                    inputs[bid] = rand_map(rng, plan, bid)
            it = run_interp(plan, inputs)
            nat = run_native(plan, inputs)
            for j, t in enumerate(outs_t):
                report(f"{tag} out{j}", nat.read_buffer(t.buf.id), it.get(t.buf.id))
            nat.close()
        except Exception as e:
            print(f"[ERR] {tag}: {type(e).__name__}: {e}", flush=True)
            traceback.print_exc()

    def dw(s, relu_in, C=728, H=32):
        def b(g):
            x = g.new(H, H, C)
            y = g.sepconv(x, "s", C, s=s, relu_in=relu_in, act="relu")
            return [x], [y]
        return b

    run("sepconv 728 s1 relu_in", dw(1, True))
    run("sepconv 64 s2 64x64", dw(2, False, 64, 64))

    def pools(g):
        x = g.new(64, 64, 64)
        return [x], [g.maxpool(x, 2, 2), g.maxpool(x, 2, 4), g.maxpool(x, 3, 2, same=True)]
    run("maxpool k2s2 / k2s4 / k3s2same", pools)

    def addn(g):
        a, b, c = g.new(64, 64, 32), g.new(32, 32, 32), g.new(16, 16, 32)
        cat = g.buf(64, 64, 64)
        y = g.addn([(a, 1), (b, 2), (c, 4)], out=G.T(cat, 32, 32))
        z = g.upsample(c, 4, out=G.T(cat, 0, 32))
        return [a, b, c], [y]
    run("addn + upsample into slices", addn)

    def se(g):
        x = g.new(32, 32, 64)
        v = g.gap(x)
        v1 = g.dense([v], "fc1", 32, bn="bn1", act="relu")
        v2 = g.dense([v1], "fc2", 64, bn="bn2", act="sigmoid")
        return [x], [g.gate_se(x, v2)]
    run("gap + dense + SE gate", se)

    def scse(C, H):
        def b(g):
            x = g.new(H, H, C)
            return [x], [g.scse(x, "q")]
        return b
    run("scSE 64ch 64x64", scse(64, 64))
    run("scSE 256ch 32x32", scse(256, 32))
    run("scSE 512ch 16x16", scse(512, 16))

    def bam(g):
        from building_detection_b200.predict_model._xception import bam_attention
        x = g.new(32, 32, 128)
        return [x], [bam_attention(g, x, "bam")]
    run("BAM block 128ch", bam)

    def neck(g):
        from building_detection_b200.predict_model._xception import neck as nk
        x = g.new(32, 32, 2048)
        return [x], [nk(g, x)]
    run("SK + ASPP neck", neck, n=1)


def models(umma, names):
    from building_detection_b200.predict_model.res34 import ResNetFamily
    from building_detection_b200.predict_model.hrnet import HRNet
    from building_detection_b200.predict_model.v3plus import Xception_DeepLabV3_Plus
    from building_detection_b200.predict_model.scse import UNet
    from building_detection_b200.predict_model.bam import Xception_DeepLabV3_Plus_bam
    from building_detection_b200.runtime import NativePlan
    from oracle import plan_interp
    ctors = {"res34": lambda: ResNetFamily().run_model("res34"), "hrnet": HRNet, "v3plus": Xception_DeepLabV3_Plus,
             "scse": lambda: UNet(2), "bam": Xception_DeepLabV3_Plus_bam}
    rng = np.random.default_rng(0)
    x = (rng.integers(0, 256, (1, 512, 512, 3), dtype=np.uint8) / 127.5 - 1).astype(np.float32)
    for name in names:
        try:
            m = ctors[name]()
            m.set_weights(G.init_weights(m.spec, seed=1, randomize_bn=True))
            plan = m.build_plan(1, umma=umma)
            t = time.time()
            it = plan_interp.Interp(plan, True)
            import torch
            with torch.no_grad():
                ref = it.run(x)
            t_cpu = time.time() - t
            nat = NativePlan(plan)
            t = time.time()
            got = nat.run_host(x)
            t_gpu = time.time() - t
            d = np.abs(got - ref)
            agree = (got.argmax(-1) == ref.argmax(-1)).mean()
            print(f"[{'OK ' if d.max() < 5e-2 else 'BAD'}] model {name} umma={umma}: probs max|d|={d.max():.3e} mean|d|={d.mean():.3e} "
                  f"argmax agree={agree:.6f} launches={nat.num_launches} arena={nat.arena_bytes/2**20:.0f}MiB cpu={t_cpu:.1f}s gpu={t_gpu:.3f}s",
                  flush=True)
            # first divergent buffer (helps localise a broken op)
            worst = []
            for b in plan.bufs:
                g_ = nat.read_buffer(b.id)
                r_ = it.get(b.id)
                e = np.abs(g_ - r_).max() / (np.abs(r_).max() + 1e-6)
                worst.append((b.id, e))
            bad = [(i, e) for i, e in worst if e > 0.05]
            print(f"      buffers with rel err > 5%: {bad[:10]}", flush=True)
            ms, kinds, flops = nat.time_ops()
            ms, kinds, flops = nat.time_ops()
            for kc, nm in ((0, "conv_umma"), (1, "conv_direct"), (2, "other")):
                sel = kinds == kc
                tf = flops[sel].sum() / max(ms[sel].sum(), 1e-9) / 1e9
                print(f"      {nm:12s} ops={sel.sum():4d} time={ms[sel].sum():9.3f} ms flops={flops[sel].sum()/1e9:9.2f} G -> {tf:8.1f} TFLOP/s",
                      flush=True)
            nat.close()
        except Exception as e:
            print(f"[ERR] model {name}: {type(e).__name__}: {e}", flush=True)
            traceback.print_exc()


def group_perf():
    """Per-layer timing of the dominant conv shapes at batch 16."""
    from building_detection_b200.runtime import NativePlan
    shapes = [("3x3 64->64 @512", 16, 512, 64, 64, 3, 1), ("3x3 128->128 @256", 16, 256, 128, 128, 3, 1),
              ("3x3 256->256 @128", 16, 128, 256, 256, 3, 1), ("3x3 512->512 @64", 16, 64, 512, 512, 3, 1),
              ("3x3 1024->1024 @32", 16, 32, 1024, 1024, 3, 1), ("1x1 728->728 @32", 16, 32, 728, 728, 1, 1),
              ("3x3 d12 2048->256 @32", 16, 32, 2048, 256, 3, 12), ("3x3 32->32 @256", 16, 256, 32, 32, 3, 1),
              ("3x3 384->384 @128", 16, 128, 384, 384, 3, 1), ("3x3 640->640 @64", 16, 64, 640, 640, 3, 1)]
    for tag, N, H, Cin, Cout, k, d in shapes:
        def builder(g):
            x = g.new(H, H, Cin)
            return x, g.conv(x, "c", Cout, k=k, d=d, bn=True, act="relu")
        try:
            plan, (x, y), _ = build_two_pass(builder, N)
            nat = NativePlan(plan)
            best = 1e9
            for _ in range(5):
                ms, kinds, flops = nat.time_ops()
                best = min(best, ms[0])
            print(f"[perf] {tag:26s} {plan.ops[0]['path']:6s} {best:8.3f} ms  {flops[0]/best/1e9:8.1f} TFLOP/s", flush=True)
            nat.close()
        except Exception as e:
            print(f"[ERR] perf {tag}: {e}", flush=True)


if __name__ == "__main__":
    all_models = ["hrnet", "scse", "res34", "v3plus", "bam"]
    for grp in sys.argv[1:]:
        print(f"==== {grp}", flush=True)
        if grp == "direct":
            group_direct()
        elif grp == "umma_basic":
            group_umma_basic()
        elif grp == "umma_more":
            group_umma_more()
        elif grp == "ops":
            group_ops()
        elif grp == "models_direct":
            models(False, all_models)
        elif grp == "models_umma":
            models(True, all_models)
        elif grp == "perf":
            group_perf()
        elif grp.startswith("model:"):
            models(True, [grp.split(":")[1]])
        else:
            print("unknown group", grp)
