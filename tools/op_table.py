"""Per-op device times of the five batch-16 plans (CUDA events around every op, bd_plan_time_ops), with layer
names and shapes: the table that says where the forward's time goes.  usage: python tools/op_table.py [batch] [top]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from building_detection_b200 import graph as G  # noqa: E402
from building_detection_b200.predict_model import CTORS, MODEL_NAMES  # noqa: E402

_args = [a for a in sys.argv[1:] if a not in ("--batch", "--top")]  # "32 30" and "--batch 32 --top 30" both work
batch = int(_args[0]) if len(_args) > 0 else 16
top = int(_args[1]) if len(_args) > 1 else 14
KIND = {G.OP_CONV: "conv", G.OP_DWCONV: "dwconv", G.OP_MAXPOOL: "maxpool", G.OP_ADDN: "addn", G.OP_GAP: "gap",
        G.OP_DENSE: "dense", G.OP_GATE: "gate", G.OP_SKFUSE: "skfuse", G.OP_BCAST: "bcast", G.OP_SOFTMAX2: "softmax"}
grand = {}
for name in MODEL_NAMES:
    m = CTORS[name]()
    nat = m.native_plan(batch)
    plan = nat.plan
    best = None
    for _ in range(3):
        ms, kinds, flops = nat.time_ops()
        best = ms if best is None else np.minimum(best, ms)
    ms = best
    assert len(ms) == len(plan.ops), (len(ms), len(plan.ops))
    # the whole plan back to back without per-op events (what the scene loop sees)
    import torch
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        nat.run_device(0, 0, 0, st)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(5):
        nat.run_device(0, 0, 0, st)
    ev1.record()
    torch.cuda.synchronize()
    whole = ev0.elapsed_time(ev1) / 5
    whole_all = whole_all + whole if "whole_all" in dir() else whole
    rows = []
    for i, op in enumerate(plan.ops):
        k = KIND[op["op"]]
        desc = ""
        if op["op"] == G.OP_CONV:
            cin, cout = op["x"][2], op["y"][2]
            desc = f"{op['path']} {len(op['taps'])}tap s{op['stride']} {cin}->{cout} @{op['Ho']}x{op['Wo']}"
            k = "conv_" + op["path"]
        elif "y" in op and isinstance(op["y"], tuple):
            b = plan.bufs[op["y"][0]]
            desc = f"C{op['y'][2]} @{b.H}x{b.W}"
        rows.append((ms[i], k, op.get("name", ""), desc, flops[i]))
        grand.setdefault(k, [0.0, 0.0]); grand[k][0] += ms[i]; grand[k][1] += flops[i]
    tot = ms.sum()
    print(f"==== {name}: {tot:.2f} ms / batch {batch}  ({m.flops_per_tile * batch / tot / 1e9:.0f} TFLOP/s algorithmic), {len(ms)} ops; "
          f"whole plan back to back {whole:.2f} ms")
    agg = {}
    for r in rows:
        key = (r[1], r[3])
        a = agg.setdefault(key, [0.0, 0, 0.0]); a[0] += r[0]; a[1] += 1; a[2] += r[4]
    for (k, desc), (t, n, fl) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"   {t:8.3f} ms {100 * t / tot:5.1f}%  x{n:<3d} {k:12s} {desc:40s} {fl / max(t, 1e-9) / 1e9:8.1f} TFLOP/s")
tot = sum(v[0] for v in grand.values())
print(f"==== all five: {tot:.2f} ms / batch {batch} -> {batch / tot * 1e3:.1f} tiles/s; whole plans back to back {whole_all:.2f} ms "
      f"-> {batch / whole_all * 1e3:.1f} tiles/s")
for k, (t, fl) in sorted(grand.items(), key=lambda kv: -kv[1][0]):
    print(f"   {k:12s} {t:8.2f} ms {100 * t / tot:5.1f}%  {fl / max(t, 1e-9) / 1e9:8.1f} TFLOP/s")
