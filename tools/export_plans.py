"""Write plan files for a non-Python host: one `<model>_b<batch>.bdplan` per network (bd_plan_save) from the models'
current weights -- Keras .h5 checkpoints if given, else the seeded random initialisation.  Needs a B200 (plans are
finalized on the device before they are saved).
usage: python tools/export_plans.py OUT_DIR [batch] [name=weights.h5 ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from building_detection_b200.predict_model import CTORS, MODEL_NAMES  # noqa: E402


def export(out_dir, batch=16, weights=None, models=None):
    os.makedirs(out_dir, exist_ok=True)
    paths = []
    for name in MODEL_NAMES:
        m = models[name] if models and name in models else CTORS[name]()
        if weights and name in weights:
            m.load_weights(weights[name])
        p = os.path.join(out_dir, f"{name}_b{batch}.bdplan")
        m.native_plan(batch).save(p)
        paths.append(p)
    return paths


if __name__ == "__main__":
    out = sys.argv[1]
    batch = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 16
    w = dict(a.split("=", 1) for a in sys.argv[2:] if "=" in a)
    for p in export(out, batch, w):
        print(p, os.path.getsize(p) // 1024, "KB")
