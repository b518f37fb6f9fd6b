"""Which HRNet convolutions need more than fp16 weight precision?  CPU study with the fp16-faithful interpreter:
activations stay fp16, the weights of the selected layer groups are kept in fp32 (what a hi/lo fp16 weight split
as extra taps would give), error = max|prob - fp32 oracle| on SURVEY 8d config 4 (seed 3).
usage: python tools/hrnet_split_study.py [tiles]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from building_detection_b200.predict_model import CTORS  # noqa: E402
from oracle import nets, plan_interp  # noqa: E402

ntile = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rng = np.random.default_rng(3)
x = (rng.integers(0, 256, (2, 512, 512, 3), dtype=np.uint8) / 127.5 - 1).astype(np.float32)[:ntile]
m = CTORS["hrnet"]()
rng99 = np.random.default_rng(99)
cal = (rng99.integers(0, 256, (1, 512, 512, 3), dtype=np.uint8) / 127.5 - 1).astype(np.float32)
w = nets.calibrated_weights("hrnet", m.spec, 1, cal)
m.set_weights(w)
with torch.no_grad():
    ref = nets.FORWARD["hrnet"](w, x)
plan = m.build_plan(ntile, keep_f32=True)
names = sorted({op["name"].split("[")[0] for op in plan.ops if op.get("name")})
print(len(plan.ops), "ops;", len(names), "named convs")


def run(pred, label):
    for op in plan.ops:
        op["w_exact"] = bool(op.get("name")) and pred(op["name"])
    t0 = time.time()
    got = plan_interp.run_plan(plan, x, emulate_h16=True)
    n = sum(1 for op in plan.ops if op.get("w_exact"))
    print(f"{label:40s} exact={n:3d}  max|dp|={np.abs(got - ref).max():.4e}  p99.99={np.quantile(np.abs(got - ref), 0.9999):.3e}  ({time.time() - t0:.0f}s)",
          flush=True)


run(lambda n: False, "fp16 weights everywhere")
run(lambda n: True, "fp32 weights everywhere")
run(lambda n: n.startswith(("stem", "l1_")), "stem + layer1")
run(lambda n: n.startswith(("stem", "l1_", "t1_", "b1_", "f1_")), "... + stage 1")
run(lambda n: n.startswith(("stem", "l1_", "t1_", "b1_", "f1_", "t2_", "b2_", "f2_")), "... + stage 2")
run(lambda n: n.startswith(("b3_", "f3_", "t3_", "head")), "stage 3 + head only")
run(lambda n: "_0" in n or n.startswith(("stem", "l1_", "head")), "branch 0 chain + stem + l1 + head")
