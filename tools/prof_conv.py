"""One tcgen05 conv layer at batch 16 (BATCH=n in the environment: another batch), launched a few times: the short
command ncu --set full replays.  usage: python tools/prof_conv.py [H Cin Cout k]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from util import build_two_pass  # noqa: E402
from building_detection_b200.runtime import NativePlan  # noqa: E402

H, Cin, Cout, k = (int(a) for a in sys.argv[1:5]) if len(sys.argv) >= 5 else (128, 256, 256, 3)
B = int(os.environ.get("BATCH", "16"))


def builder(g):
    x = g.new(H, H, Cin)
    return x, g.conv(x, "c", Cout, k=k, bn=True, act="relu")


plan, (x, y), _ = build_two_pass(builder, B)
nat = NativePlan(plan)
nat.write_buffer(x.buf.id, np.random.default_rng(0).standard_normal((B, H, H, Cin)).astype(np.float32))
best = 1e9
for _ in range(5):
    ms, kinds, flops = nat.time_ops()
    best = min(best, ms[0])
print(f"conv {k}x{k} {Cin}->{Cout} @{H}^2 batch {B}: {best:.3f} ms, {flops[0] / best / 1e9:.1f} TFLOP/s")
nat.close()
