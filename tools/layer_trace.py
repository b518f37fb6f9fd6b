"""Per-layer divergence table: the GPU's plan buffers (NativePlan.read_buffer) against the fp16-faithful CPU
interpreter of the same plan, op by op -- shows whether an end-to-end gap enters at one kernel (a bug) or grows
smoothly (amplified summation-order noise).  Needs a B200.
usage: python tools/layer_trace.py [model] [undamped]      (output is committed under profiles/)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from building_detection_b200 import graph as G  # noqa: E402
from building_detection_b200.predict_model import CTORS  # noqa: E402
from building_detection_b200.runtime import NativePlan  # noqa: E402
from oracle import nets, plan_interp  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "hrnet"
damped = not (len(sys.argv) > 2 and sys.argv[2] == "undamped")
seed = {"res34": 0, "v3plus": 1, "scse": 2, "bam": 2, "hrnet": 3}[name]
rng = np.random.default_rng(seed)
x = (rng.integers(0, 256, (2, 512, 512, 3), dtype=np.uint8) / 127.5 - 1).astype(np.float32)[:1]
m = CTORS[name]()
m.set_weights(nets.parity_weights(name, m.spec, hrnet_damped=damped))
plan = m.build_plan(1)
nat = NativePlan(plan)
got = nat.run_host(x)
it = plan_interp.Interp(plan, True)
with torch.no_grad():
    want = it.run(x)
KIND = {G.OP_CONV: "conv", G.OP_DWCONV: "dwconv", G.OP_MAXPOOL: "maxpool", G.OP_ADDN: "addn", G.OP_GAP: "gap",
        G.OP_DENSE: "dense", G.OP_GATE: "gate", G.OP_SKFUSE: "skfuse", G.OP_BCAST: "bcast", G.OP_SOFTMAX2: "softmax"}
print(f"# {name} ({'parity' if damped else 'undamped'} recipe), 1 tile: GPU vs fp16 interpreter, per op output")
print(f"# {'op':>4} {'kind':8} {'name':22} {'shape':>16} {'rms':>10} {'max|d|':>10} {'max|d|/rms':>10} {'frac != ':>9} {'max ulp':>8}")
seen = {}
for i, op in enumerate(plan.ops):
    k = op["op"]
    if k == G.OP_SOFTMAX2:
        continue
    y = op["y"]
    bid, c0, c = (y, 0, None) if isinstance(y, int) else y
    a = nat.read_buffer(bid)
    b = it.get(bid)
    if c is not None:
        a, b = a[..., c0:c0 + c], b[..., c0:c0 + c]
    if op.get("out_scale", 1) > 1:
        s, oy, ox = op["out_scale"], op["out_oy"], op["out_ox"]
        a, b = a[:, oy::s, ox::s], b[:, oy::s, ox::s]
    d = np.abs(a - b)
    rms = float(np.sqrt((b.astype(np.float64) ** 2).mean())) + 1e-30
    ulp = 2.0 ** (np.floor(np.log2(np.maximum(np.abs(b), 2.0 ** -14))) - 10)
    print(f"{i:6d} {KIND[k]:8} {op.get('name', ''):22} {str(tuple(a.shape[1:])):>16} {rms:10.3e} {d.max():10.3e} "
          f"{d.max() / rms:10.3e} {float((d > 0).mean()):9.2e} {float((d / ulp).max()):8.1f}")
d = np.abs(got - want)
print(f"# probabilities: max|d| {d.max():.3e}  p99.99 {np.quantile(d, 0.9999):.3e}  mean {d.mean():.3e}")
with torch.no_grad():
    ref = nets.FORWARD[name](m.get_weights(), x)
print(f"# vs fp32 oracle: GPU max|dp| {np.abs(got - ref).max():.3e}, interpreter max|dp| {np.abs(want - ref).max():.3e}")
