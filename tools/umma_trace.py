"""Event trace of CTA 0 of one tcgen05 conv (debug aid): when did the producer get a free stage, when did the MMA
warp see the data, when did the epilogue start / finish a tile.  usage: BD_UMMA_TRACE=1 python tools/umma_trace.py H Cin Cout k"""
import os
import sys

os.environ["BD_UMMA_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from util import build_two_pass  # noqa: E402
from building_detection_b200 import runtime as R  # noqa: E402

H, Cin, Cout, k = (int(a) for a in sys.argv[1:5]) if len(sys.argv) >= 5 else (512, 64, 64, 3)


def builder(g):
    x = g.new(H, H, Cin)
    if k == 0:  # fused separable convolution
        return x, g.sepconv(x, "s", Cout, relu_in=True, act="relu")
    return x, g.conv(x, "c", Cout, k=k, bn=True, act="relu")


plan, (x, y), _ = build_two_pass(builder, 16)
nat = R.NativePlan(plan)
nat.run_device(0, 0, 0)
buf = np.zeros(4 * 4096, np.int64)
R.check(R.lib().bd_debug_read_trace(nat.ctx, R._ptr(buf), 0))
nat.run_device(0, 0, 0)
R.check(R.lib().bd_debug_read_trace(nat.ctx, R._ptr(buf), 0))
names = {0: "prod: stage free ", 1: "mma : data ready ", 2: "epi : acc ready  ", 3: "epi : tile done  "}
evs = []
for role in range(4):
    b = buf[role * 4096:(role + 1) * 4096]
    n = int(b[0])
    for a_, b_, c_ in b[1:1 + 3 * n].reshape(n, 3):
        evs.append((int(c_), role, int(a_), int(b_)))
evs.sort()
t0 = evs[0][0]
print(f"{len(evs)} events; conv {k}x{k} {Cin}->{Cout} @{H}")
last = {}
skip = int(os.environ.get("TRACE_SKIP", "0"))
for c, role, a_, b_ in evs[skip:skip + int(os.environ.get("TRACE_LINES", "120"))]:
    d = c - last.get(role, c)
    last[role] = c
    if role == 3 and b_:
        nm = {1: "epi/dw: wait", 2: "epi : tmem loaded", 3: "epi/dw: go", 4: "epi : fence+bar", 5: "epi/dw: stored"}[b_]
        print(f"{c - t0:9d} (+{d:6d})  {nm} tile {a_:6d}")
        continue
    if role == 1 and b_ < 0:
        nm = {-1: "committed", -2: "mmas issued", -10: "loop top", -11: "tempty ok", -12: "fence1", -13: "full ok", -14: "fence2"}[b_]
        print(f"{c - t0:9d} (+{d:6d})  mma : {nm} tile {a_:6d}")
        continue
    print(f"{c - t0:9d} (+{d:6d})  {names[role]} tile {a_:6d} kb {b_}")
for role in range(4):
    cs = np.array(sorted(c for c, r, _, _ in evs if r == role))
    if len(cs) > 40:
        print(f"role {role}: median gap {np.median(np.diff(cs[20:])):.0f} clk, mean {np.mean(np.diff(cs[20:])):.0f}, events {len(cs)}")
