"""DeepLabv3+ (Xception backbone, ASPP, SK block, scSE U-decoder) -- B200 plan for the graph of
reference predict_model/v3plus.py:170-350."""
from ..engine import Model
from ..graph import Net, T
from . import _xception as xc


def build(g: Net):
    x = g.input()
    cat1 = g.buf(64, 64, 512)    # [up2(neck) | c2]   v3plus.py:323
    cat2 = g.buf(128, 128, 256)  # [convT(dec1) | c1] v3plus.py:330
    cat3 = g.buf(256, 256, 128)  # [c | convT(dec2)]  v3plus.py:336
    _c, _c1, _c2, c5 = xc.backbone(g, x, with_bam=False, c_out=T(cat3, 0, 64),
                                   c1_out=T(cat2, 128, 128), c2_out=T(cat1, 256, 256))
    n = xc.neck(g, c5)
    g.upsample(n, 2, out=T(cat1, 0, 256))  # v3plus.py:321

    def stage(cat, name, ch):  # two conv_bn_relu + scSE, v3plus.py:324-326 etc.
        t = g.conv(T(cat, 0, cat.C), name + "_a", ch, k=3, bn=True, act="relu")
        t = g.conv(t, name + "_b", ch, k=3, bn=True, act="relu")
        return g.scse(t, name + "_scse")

    t = stage(cat1, "dec1", 256)
    g.conv_transpose(t, "dec2_up", 128, 3, out=T(cat2, 0, 128))  # no activation, v3plus.py:328
    t = stage(cat2, "dec2", 128)
    g.conv_transpose(t, "dec3_up", 64, 3, out=T(cat3, 64, 64))  # v3plus.py:335
    t = stage(cat3, "dec3", 64)
    # UpSampling2D(2) + conv3x3 + BN + ReLU (v3plus.py:341-342) as four sub-pixel 2x2 convolutions of the 256^2 map
    o = g.conv_up2(t, "head_a", 32, bn=True, act="relu")
    o = g.conv(o, "head_b", 32, k=3, bn=True, act="relu")
    logits = g.conv(o, "head_out", 2, k=1, f32_out=True)  # v3plus.py:345
    g.softmax_head(logits)


def Xception_DeepLabV3_Plus(shape=(512, 512, 3), num_classes=2):
    """Drop-in for reference predict_model/v3plus.py:170."""
    if num_classes != 2:
        raise ValueError("the B200 head kernel is the reference's 2-class softmax")
    return Model("v3plus", build, tuple(shape))
