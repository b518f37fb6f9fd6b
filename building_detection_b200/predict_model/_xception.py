"""Pieces shared by the two DeepLabv3+ plans (reference predict_model/v3plus.py and bam.py, whose
backbone / SK block / ASPP / scSE code is identical): Xception-65-style backbone at output
stride 16, selective-kernel block, ASPP, and the 512-channel neck."""
from ..graph import Net, T


def bam_attention(g: Net, x, name):
    """BAM block, bam.py:20-71: x * (1 + sigmoid(channel_gate(x) + spatial_gate(x))).
    The C/16-channel maps of the spatial gate (4..45 channels, dilation 4) are zero-padded to a multiple of 16
    channels so that these convolutions are tensor-core tiles too; the 1-channel gate logits stay fp32."""
    r = x.C // 16
    rp = -(-r // 16) * 16
    v = g.gap(x)
    v = g.dense([v], name + "_cg1", r, bn=name + "_cg1_bn", act="relu")
    v = g.dense([v], name + "_cg2", r, bn=name + "_cg2_bn", act="relu")
    cg = g.dense([v], name + "_cg3", x.C)
    s = g.conv(x, name + "_sg1", r, k=1, bn=True, act="relu", cout_pad=rp)
    s = g.conv(s, name + "_sg2", r, k=3, d=4, bn=True, act="relu", cout_pad=rp)
    s = g.conv(s, name + "_sg3", r, k=3, d=4, bn=True, act="relu", cout_pad=rp)
    s = g.conv(s, name + "_sg4", 1, k=1, f32_out=True)
    return g.gate_bam(x, cg, s)


def backbone(g: Net, x, with_bam, c_out=None, c1_out=None, c2_out=None):
    """v3plus.py:173-280 / bam.py:173-277.  Returns (c, c1, c2, c5); the skip tensors can be
    produced directly into decoder concat slices."""
    t = g.conv(x, "entry1", 32, k=3, s=2, bn=True, act="relu")
    t = g.conv(t, "entry2", 64, k=3, bn=True, act="relu", out=None if with_bam else c_out)
    if with_bam:
        t = bam_attention(g, t, "bam1")  # bam.py:182 (the bam decoder never uses c)
    c = t

    # block 1 (v3plus.py:185-194): sep+BN+ReLU, sep+BN, MaxPool(3,2,same), + strided 1x1 residual
    g.declare_conv(t.C, "b1_res", 128, k=1, bn=True)  # the reference builds the residual conv + BN first
    m = g.sepconv(t, "b1_s1", 128, act="relu")
    m = g.sepconv(m, "b1_s2", 128)
    m = g.maxpool(m, 3, 2, same=True)
    c1 = g.conv(t, "b1_res", 128, k=1, s=2, bn=True, res=m, out=c1_out)
    t = bam_attention(g, c1, "bam2") if with_bam else c1  # bam.py:194-198: c1 is taken before BAM

    def entry_block(t, name, ch, out=None):  # v3plus.py:202-236: [ReLU, sep, BN] x3 (last strided) + residual
        r = g.conv(t, name + "_res", ch, k=1, s=2, bn=True)
        m = g.sepconv(t, name + "_s1", ch, relu_in=True)
        m = g.sepconv(m, name + "_s2", ch, relu_in=True)
        return g.sepconv(m, name + "_s3", ch, s=2, relu_in=True, res=r, out=out)

    c2 = entry_block(t, "b2", 256, out=c2_out)
    t = bam_attention(g, c2, "bam3") if with_bam else c2
    t = entry_block(t, "b3", 728)

    for i in range(16):  # middle flow, v3plus.py:239-252
        r = t
        for j in range(3):
            t = g.sepconv(t, f"mid{i}_s{j}", 728, relu_in=True, res=r if j == 2 else None)
    if with_bam:
        t = bam_attention(g, t, "bam4")  # bam.py:253

    # exit flow, v3plus.py:258-280 (stride 1: output stride stays 16)
    r = g.conv(t, "exit_res", 1024, k=1, bn=True)
    m = g.sepconv(t, "exit_s1", 728, relu_in=True)
    m = g.sepconv(m, "exit_s2", 1024, relu_in=True)
    t = g.sepconv(m, "exit_s3", 1024, relu_in=True, res=r)
    t = g.sepconv(t, "exit_s4", 1536, act="relu")
    t = g.sepconv(t, "exit_s5", 1536, act="relu")
    c5 = g.sepconv(t, "exit_s6", 2048, act="relu")
    return c, c1, c2, c5


def neck(g: Net, c5, out=None):
    """SK block || ASPP -> 1x1 proj -> concat -> 2x conv3x3 -> scSE (v3plus.py:74-138, 295-316)."""
    H, W = c5.H, c5.W

    def cbr(t, name, cout, k, d=1, out=None):  # inner conv_bn_relu, v3plus.py:288-293
        return g.conv(t, name, cout, k=k, d=d, bn=True, act="relu", out=out)

    cat = g.buf(H, W, 512)  # [aspp_proj | sk]  (v3plus.py:313)

    # --- SKNet_block, v3plus.py:74-138
    conv = cbr(c5, "sk_in", 256, 3)
    ds = [cbr(conv, "sk_d1", 256, 1), cbr(conv, "sk_d6", 256, 3, 6),
          cbr(conv, "sk_d12", 256, 3, 12), cbr(conv, "sk_d18", 256, 3, 18)]
    gv = g.dense([g.gap(conv)], "sk_gap", 256, bn="sk_gap_bn", act="relu", conv_kernel=True)
    # GAP(d1+d6+d12+d18+broadcast(gv)) = sum of the pooled vectors (mean is linear)
    pooled = [g.gap(d) for d in ds] + [gv]
    sq = g.dense(pooled, "sk_squeeze", 16, bn="sk_squeeze_bn", act="relu", conv_kernel=True)
    logits = [g.dense([sq], f"sk_w{i}", 256, conv_kernel=True) for i in range(5)]
    g.skfuse(ds, gv, logits, "sk_out_bn", out=T(cat, 256, 256))

    # --- ASPP, v3plus.py:295-307: concat [1x1 | d6 | d12 | d18 | pooled] = 1280 channels
    acat = g.buf(H, W, 1280)
    cbr(c5, "aspp_1x1", 256, 1, out=T(acat, 0, 256))
    cbr(c5, "aspp_d6", 256, 3, 6, out=T(acat, 256, 256))
    cbr(c5, "aspp_d12", 256, 3, 12, out=T(acat, 512, 256))
    cbr(c5, "aspp_d18", 256, 3, 18, out=T(acat, 768, 256))
    pv = g.dense([g.gap(c5)], "aspp_pool", 256, bn="aspp_pool_bn", act="relu", conv_kernel=True)
    g.bcast(pv, T(acat, 1024, 256))  # AveragePooling2D(32) on 32x32 + UpSampling2D(32)
    cbr(T(acat, 0, 1280), "aspp_proj", 256, 1, out=T(cat, 0, 256))

    t = cbr(T(cat, 0, 512), "neck1", 256, 3)
    t = cbr(t, "neck2", 256, 3)
    return g.scse(t, "neck_scse", out=out)
