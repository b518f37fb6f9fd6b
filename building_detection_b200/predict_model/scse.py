"""SCSE U-Net (no BatchNorm) -- B200 plan for the graph of reference predict_model/scse.py:49-97."""
from ..engine import Model
from ..graph import Net, T

WIDTHS = (64, 128, 256, 512, 1024)


def build(g: Net):
    x = g.input()

    def cr(t, name, ch, out=None):  # Conv2D(ch, 3, 'same', activation='relu'), glorot, no BN
        return g.conv(t, name, ch, k=3, act="relu", out=out)

    # decoder concat buffers [convT(up) | skip] (scse.py:72,78,84,90): skips land in their slice
    cats = [g.buf(512 >> lvl, 512 >> lvl, 2 * WIDTHS[lvl]) for lvl in range(4)]
    t = x
    for lvl in range(4):  # scse.py:52-66
        ch = WIDTHS[lvl]
        t = cr(t, f"enc{lvl + 1}_a", ch)
        skip = cr(t, f"enc{lvl + 1}_b", ch, out=T(cats[lvl], ch, ch))
        t = g.maxpool(skip, 2, 2)
    t = cr(cr(t, "enc5_a", 1024), "enc5_b", 1024)  # scse.py:68-69
    for i, lvl in enumerate((3, 2, 1, 0), start=1):  # scse.py:71-93
        ch = WIDTHS[lvl]
        g.conv_transpose(t, f"dec{i}_up", ch, 3, act="relu", out=T(cats[lvl], 0, ch))
        t = cr(T(cats[lvl], 0, 2 * ch), f"dec{i}_a", ch)
        t = cr(t, f"dec{i}_b", ch)
        t = g.scse(t, f"dec{i}_scse")
    logits = g.conv(t, "head_out", 2, k=1, f32_out=True)  # scse.py:95
    g.softmax_head(logits)


def UNet(num_classes=2, input_shape=(512, 512, 3)):
    """Drop-in for reference predict_model/scse.py:49."""
    if num_classes != 2:
        raise ValueError("the B200 head kernel is the reference's 2-class softmax")
    return Model("scse", build, tuple(input_shape))
