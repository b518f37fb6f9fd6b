"""DeepLabv3+ with BAM attention in the backbone and a nearest-upsample decoder -- B200 plan
for the graph of reference predict_model/bam.py:170-338."""
from ..engine import Model
from ..graph import Net, T
from . import _xception as xc


def build(g: Net):
    x = g.input()
    cat1 = g.buf(64, 64, 512)    # [c2 | up2(neck)]  bam.py:320-321
    cat2 = g.buf(128, 128, 256)  # [c1 | up2(dec1)]  bam.py:327
    _c, _c1, _c2, c5 = xc.backbone(g, x, with_bam=True, c1_out=T(cat2, 0, 128), c2_out=T(cat1, 0, 256))
    n = xc.neck(g, c5)
    g.upsample(n, 2, out=T(cat1, 256, 256))
    t = g.conv(T(cat1, 0, 512), "dec1_a", 128, k=3, bn=True, act="relu")
    t = g.conv(t, "dec1_b", 128, k=3, bn=True, act="relu")
    t = g.scse(t, "dec1_scse")
    g.upsample(t, 2, out=T(cat2, 128, 128))  # bam.py:325
    t = g.conv(T(cat2, 0, 256), "dec2_a", 64, k=3, bn=True, act="relu")
    t = g.conv(t, "dec2_b", 64, k=3, bn=True, act="relu")
    t = g.scse(t, "dec2_scse")
    # bam.py:332-333: UpSampling2D(4) then 1x1 conv + softmax.  A 1x1 conv and a per-pixel softmax
    # commute with nearest replication, so the logits are computed at 128x128 and replicated 4x4.
    logits = g.conv(t, "head_out", 2, k=1, f32_out=True)
    g.softmax_head(logits, up=4)


def Xception_DeepLabV3_Plus_bam(shape=(512, 512, 3), num_classes=2):
    """Drop-in for reference predict_model/bam.py:170."""
    if num_classes != 2:
        raise ValueError("the B200 head kernel is the reference's 2-class softmax")
    return Model("bam", build, tuple(shape))
