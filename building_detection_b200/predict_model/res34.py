"""res34 U-Net -- B200 plan for the graph of reference predict_model/res34.py:27-170.

Same call surface as the reference: ``ResNetFamily(input_shape).run_model('res34')`` returns a
model object with ``.predict`` / ``.load_weights`` (see ``engine.Model``).  The graph below is
expressed in fused plan ops (graph.py), not Keras layers.
"""
from ..engine import Model
from ..graph import Net, T

STAGES = ((2, 64, 3), (3, 128, 4), (4, 256, 6), (5, 512, 3))  # (index, channels, blocks): res34.py:54-68


def build(g: Net):
    x = g.input()

    def cbr(t, name, out=None):  # bn_conv_a, res34.py:32-38 (he_normal, BN named '<name>_BN')
        return g.conv(t, name, 64 if t.buf.id == g.plan.input else t.C, k=3, bn=name + "_BN", act="relu", he=True, out=out)

    def res_block(t, name, out=None):  # res_block1, res34.py:40-45: relu(x + relu(bn(conv(relu(bn(conv x))))))
        a = cbr(t, name + "_1")
        return g.conv(a, name + "_2", t.C, k=3, bn=name + "_2_BN", act="relu", he=True,
                      res=t, res_after_act=True, out=out)

    # encoder, res34.py:47-69.  Feature maps that later feed a concat are produced straight into
    # their channel slice of the concat buffer (concat elision).
    #   l2h_a mid  = [conv2(64)  | mp2(conv1)(64)]                      @256 -> conv2' (128)
    #   l2h_a high = [conv3(128) | mp2(conv2)(64) | mp2s4(conv1)(64)]   @128 -> conv3' (256)
    #   l2h_b mid  = [conv3'(256) | mp2(conv2')(128)]                   @128 -> conv3'' (384)
    #   l2h_b high = [conv4(256) | mp2(conv3')(256) | mp2s4(conv2')(128)] @64 -> conv4' (640)
    cat_a_mid = g.buf(256, 256, 128)
    cat_a_high = g.buf(128, 128, 256)
    cat_b_mid = g.buf(128, 128, 384)
    cat_b_high = g.buf(64, 64, 640)
    conv1 = cbr(cbr(cbr(x, "conv1_1"), "conv1_2"), "conv1_3")
    feats = [conv1]
    finals = {2: T(cat_a_mid, 0, 64), 3: T(cat_a_high, 0, 128), 4: T(cat_b_high, 0, 256), 5: None}
    t = conv1
    for idx, ch, blocks in STAGES:
        t = g.conv(t, f"pool{idx - 1}", ch, k=1, s=2)  # Conv2D(1, strides=2): bias, no BN, no act
        for i in range(blocks):
            t = res_block(t, f"conv{idx}_{i}", out=finals[idx] if i == blocks - 1 else None)
        feats.append(t)
    conv1, conv2, conv3, conv4, conv5 = feats

    # low_to_high_feature #1 (res34.py:151-159, call at :74)
    g.maxpool(conv1, 2, 2, out=T(cat_a_mid, 64, 64))
    g.maxpool(conv2, 2, 2, out=T(cat_a_high, 128, 64))
    g.maxpool(conv1, 2, 4, out=T(cat_a_high, 192, 64))  # MaxPool2D(strides=4): 2x2 window, valid
    conv3p = g.conv(T(cat_a_high, 0, 256), "l2h_a_high", 256, act="relu", he=True, out=T(cat_b_mid, 0, 256))
    conv2p = g.conv(T(cat_a_mid, 0, 128), "l2h_a_mid", 128, act="relu", he=True)
    # low_to_high_feature #2 (call at :75)
    g.maxpool(conv2p, 2, 2, out=T(cat_b_mid, 256, 128))
    g.maxpool(conv3p, 2, 2, out=T(cat_b_high, 256, 256))
    g.maxpool(conv2p, 2, 4, out=T(cat_b_high, 512, 128))
    conv4p = g.conv(T(cat_b_high, 0, 640), "l2h_b_high", 640, act="relu", he=True)
    conv3pp = g.conv(T(cat_b_mid, 0, 384), "l2h_b_mid", 384, act="relu", he=True)

    def attention(t, name, out=None):  # attention_demo, res34.py:90-105
        v = g.gap(t)
        v = g.dense([v], name + "_fc1", t.C // 2, bn=name + "_bn1", act="relu")
        v = g.dense([v], name + "_fc2", t.C, bn=name + "_bn2", act="sigmoid")
        return g.gate_se(t, v, out=out)

    # decoder concat buffers [low | ConvT(high)], res34.py:143-149
    lows = {"4": conv4p, "3": conv3pp, "2": conv2p, "1": conv1}
    att_names = {"1": "att1", "2": "att2", "3": "att3", "4": "att4"}
    cats = {k: g.buf(v.H, v.W, 2 * v.C) for k, v in lows.items()}
    # attention outputs in reference order conv1..conv5 (res34.py:76-80)
    for k in ("1", "2", "3", "4"):
        attention(lows[k], att_names[k], out=T(cats[k], 0, lows[k].C))
    high = attention(conv5, "att5")
    for k in ("4", "3", "2", "1"):  # upsame_feature x4, res34.py:82-85
        c = lows[k].C
        g.conv_transpose(high, f"up{k}_convT", c, 2, act="relu", out=T(cats[k], c, c))
        mix = g.conv(T(cats[k], 0, 2 * c), f"up{k}_mix", c, act="relu", he=True)
        high = res_block(mix, f"upsame_{k}")
    o = g.conv(high, "head_conv", 64, k=3, act="relu", he=True)  # res34.py:86
    logits = g.conv(o, "head_out", 2, k=3, he=True, f32_out=True)  # res34.py:87 (softmax below)
    g.softmax_head(logits)


class ResNetFamily:
    """Drop-in for reference predict_model/res34.py:27 (``ResNetFamily().run_model('res34')``)."""

    def __init__(self, input_shape=(512, 512, 3)):
        self.input_shape = tuple(input_shape)
        self.f_size = 64

    def run_model(self, name):
        if name != "res34":
            raise ValueError("This network does not exist.")  # res34.py:167
        return Model("res34", build, self.input_shape)
