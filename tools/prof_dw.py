"""One depthwise 3x3 (+ its pointwise 1x1) at batch 16 for ncu.  usage: python tools/prof_dw.py [H C stride]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from util import build_two_pass  # noqa: E402
from building_detection_b200.runtime import NativePlan  # noqa: E402

H, C, S = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (32, 728, 1)


def builder(g):
    x = g.new(H, H, C)
    return x, g.sepconv(x, "s", C, s=S, relu_in=True)


plan, (x, y), _ = build_two_pass(builder, 16)
nat = NativePlan(plan)
nat.write_buffer(x.buf.id, np.random.default_rng(0).standard_normal((16, H, H, C)).astype(np.float32))
best = [1e9, 1e9]
for _ in range(5):
    ms, kinds, flops = nat.time_ops()
    best = [min(best[0], ms[0]), min(best[1], ms[1])]
byt = 16 * H * H * C * 2 * (1 + 1.0 / (S * S))
print(f"dwconv C{C} @{H}^2 s{S} batch 16: {best[0] * 1e3:.1f} us ({byt / best[0] / 1e6:.0f} GB/s), pointwise {best[1] * 1e3:.1f} us")
nat.close()
