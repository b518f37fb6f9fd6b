"""HBM roofline of the memory-bound plan kernels (gate, gate_scse, maxpool, gap, addn, depthwise, skfuse, bcast):
algorithmic bytes = 2 B x (fp16 elements read + written) (+ 4 B per fp32 element) per op, over the per-op device time
of bd_plan_time_ops (CUDA events), against the measured HBM copy peak.  usage: python tools/mem_table.py [batch]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from building_detection_b200 import graph as G  # noqa: E402
from building_detection_b200.predict_model import CTORS, MODEL_NAMES  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
peak = 6536.4
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["hbm_gbs"]
KIND = {G.OP_DWCONV: "dwconv", G.OP_MAXPOOL: "maxpool", G.OP_ADDN: "addn", G.OP_GAP: "gap", G.OP_GATE: "gate",
        G.OP_SKFUSE: "skfuse", G.OP_BCAST: "bcast"}


def ref_bytes(plan, r, n_up=1):
    b = plan.bufs[r[0]]
    return batch * b.H * b.W * r[2] * (4 if b.dtype == "f32" else 2)


agg = {}
for name in MODEL_NAMES:
    nat = CTORS[name]().native_plan(batch)
    plan = nat.plan
    best = None
    for _ in range(3):
        ms, _k, _f = nat.time_ops()
        best = ms if best is None else np.minimum(best, ms)
    for i, op in enumerate(plan.ops):
        k = op["op"]
        if k not in KIND or best[i] <= 0:
            continue
        kind = KIND[k]
        if k == G.OP_GATE:
            kind = {G.GATE_SE: "gate_se", G.GATE_SCSE: "gate_scse", G.GATE_BAM: "gate_bam"}[op["mode"]]
        by = 0
        if k == G.OP_ADDN:
            by = sum(ref_bytes(plan, r) for r in op["xs"]) + ref_bytes(plan, op["y"])
        elif k == G.OP_GAP:
            by = ref_bytes(plan, op["x"])
        elif k == G.OP_SKFUSE:
            by = sum(ref_bytes(plan, r) for r in op["xs"]) + ref_bytes(plan, op["y"])
        elif k == G.OP_BCAST:
            by = ref_bytes(plan, op["y"])
        else:
            by = ref_bytes(plan, op["x"]) + ref_bytes(plan, op["y"])
            if k == G.OP_GATE and op["s"] is not None:
                by += ref_bytes(plan, op["s"])
        a = agg.setdefault(kind, [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += best[i]; a[2] += by
        a[3] = max(a[3], by / best[i] / 1e6)
print(f"# memory-bound plan kernels, five batch-{batch} plans, HBM copy peak {peak:.0f} GB/s")
print("| kernel | ops | ms / batch | algorithmic GB | GB/s (all ops) | frac of peak | best op GB/s |")
print("|---|---|---|---|---|---|---|")
for kind, (n, ms, by, bst) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    gbs = by / ms / 1e6
    print(f"| {kind} | {n} | {ms:.3f} | {by / 1e9:.2f} | {gbs:.0f} | {gbs / peak:.2f} | {bst:.0f} |")
