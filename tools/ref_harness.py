"""Imports the reference's own predict.py / model_fuse.py / edge_3.py from /root/reference under the
stubs SURVEY.md Appendix E documents (tensorflow, matplotlib and the five predict_model modules are
absent or unnecessary; cv2 >= 4.5.4 returns contours as a tuple, which edge_3.py mutates).

Only the golden-vector generators (tools/make_golden_*.py) use this, in the build container; nothing
under tests/, bench.py or the product imports it, because /root/reference does not exist on the GPU box.
"""
import os
import sys
from types import ModuleType

import numpy as np

REF = os.environ.get("BD_REFERENCE", "/root/reference")


def install_stubs():
    import cv2 as cv
    if "tensorflow" not in sys.modules:
        tf = ModuleType("tensorflow")
        tf.newaxis = None
        tf.argmax = lambda x, axis=-1: np.argmax(x, axis=axis)
        tf.squeeze = np.squeeze
        sys.modules["tensorflow"] = tf
    if "matplotlib" not in sys.modules:
        plt = ModuleType("matplotlib.pyplot")
        plt.imshow = plt.show = plt.cla = lambda *a, **k: None
        mpl = ModuleType("matplotlib")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    pm = ModuleType("predict_model")
    sys.modules.setdefault("predict_model", pm)
    for mod, sym in (("res34", "ResNetFamily"), ("hrnet", "HRNet"), ("v3plus", "Xception_DeepLabV3_Plus"),
                     ("scse", "UNet"), ("bam", "Xception_DeepLabV3_Plus_bam")):
        m = ModuleType("predict_model." + mod)
        setattr(m, sym, object)
        sys.modules.setdefault("predict_model." + mod, m)
    if not getattr(cv.findContours, "_bd_list_shim", False):
        _fc = cv.findContours

        def find_contours(*a, **k):
            r = _fc(*a, **k)
            return (list(r[0]), r[1]) if len(r) == 2 else (r[0], list(r[1]), r[2])
        find_contours._bd_list_shim = True
        cv.findContours = find_contours
    if REF not in sys.path:
        sys.path.insert(0, REF)


def reference_modules():
    """(predict, model_fuse, edge_3) modules of the reference."""
    install_stubs()
    import edge_3
    import model_fuse
    import predict
    assert os.path.dirname(os.path.abspath(predict.__file__)) == os.path.abspath(REF)
    return predict, model_fuse, edge_3
