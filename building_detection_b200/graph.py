"""Host-side graph builder: lowers one of the five segmentation networks to a *plan*.

A plan is plain data -- a buffer table plus an ordered list of fused ops with their
BatchNorm-folded, fp16-quantised parameters -- that ``engine.py`` uploads once through the
C ABI (``include/bd_b200.h``) where it becomes a native launch list of sm_100a kernels.
Nothing in here computes activations; there is no CPU execution path in the product.

Fusions decided here:
  * Conv2D + BatchNormalization (+ReLU) (+residual add) (+ReLU)  -> one CONV op
  * channel concat                                              -> producers write channel slices
  * Conv2DTranspose k2/k3 stride 2                              -> 4 sub-pixel CONV ops (tap subsets)
  * SeparableConv2D                                             -> DWCONV + 1x1 CONV (BN folded)
  * GlobalAveragePooling of a sum                               -> sum of pooled vectors
All feature maps are NHWC fp16 (fp32 for the network input and the 2-channel logits); pooled
vectors and the tiny attention MLPs stay fp32.

TF/Keras semantics honoured: 'same' padding incl. the asymmetric stride-2 case, BN eps 1e-3,
nearest UpSampling2D, Conv2DTranspose crop.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

BN_EPS = 1e-3  # Keras BatchNormalization default
INPUT_C = 32         # the network input is stored im2col'ed for its 3x3 stem conv: 27 = 9 taps x RGB, padded to 32
INPUT_SCALE = 255.0  # ... holding 255 * x (exact integers for 8-bit imagery)

# op codes shared with csrc/bd_api.cu (enum bd_op_kind in include/bd_b200.h)
MAX_TAPS = 18  # bd_b200.h BD_MAX_TAPS
SMALL_MACS = 600  # multiply-adds per pixel up to which a <= 16-output-channel conv runs on CUDA cores (path "small")
OP_CONV, OP_DWCONV, OP_MAXPOOL, OP_ADDN, OP_GAP, OP_DENSE, OP_GATE, OP_SKFUSE, OP_BCAST, OP_SOFTMAX2 = range(10)
GATE_SE, GATE_SCSE, GATE_BAM = range(3)
ACT_NONE, ACT_RELU, ACT_SIGMOID = range(3)


def same_pad(size, k, s, d=1):
    """TF 'same' padding (before, after) -- asymmetric when the total is odd."""
    k_eff = (k - 1) * d + 1
    out = math.ceil(size / s)
    total = max((out - 1) * s + k_eff - size, 0)
    return total // 2, total - total // 2


H16_MAX = 65504.0


def to_h16(a):
    """fp32 -> IEEE fp16 (round to nearest even, saturating like the device stores), as uint16 bits."""
    a = np.clip(np.ascontiguousarray(a, dtype=np.float32), -H16_MAX, H16_MAX)
    return a.astype(np.float16).view(np.uint16)


def h16_to_f32(u16):
    return np.ascontiguousarray(u16, dtype=np.uint16).view(np.float16).astype(np.float32)


@dataclass
class Buf:
    id: int
    H: int
    W: int
    C: int
    dtype: str = "f16"  # 'f16' | 'f32'
    kind: str = "map"  # 'map' (N,H,W,C) | 'vec' (N,C) fp32


@dataclass
class T:
    """Channel-slice view [c0, c0+C) of a map buffer.  ``ctrue`` < C marks zero-padded channels (the network
    input is stored as 8 channels, BAM's C/16-channel maps as multiples of 16, so that every convolution is a
    tensor-core tile); ``wscale`` is folded into the weights of whatever convolution consumes the tensor (the
    input buffer holds 255*x, exact integers for 8-bit imagery)."""
    buf: Buf
    c0: int
    C: int
    ctrue: int = -1
    wscale: float = 1.0

    @property
    def cin(self):
        return self.C if self.ctrue < 0 else self.ctrue

    @property
    def H(self):
        return self.buf.H

    @property
    def W(self):
        return self.buf.W

    def ref(self):
        return (self.buf.id, self.c0, self.C)


@dataclass
class V:
    """fp32 (N, C) vector buffer."""
    buf: Buf

    @property
    def C(self):
        return self.buf.C


@dataclass
class Plan:
    model: str
    batch: int
    bufs: list = field(default_factory=list)
    ops: list = field(default_factory=list)
    input: int = -1
    input_stride: int = 1  # stride of the stem conv the im2col'ed input buffer was laid out for
    logits: int = -1
    logits_up: int = 1
    flops: int = 0  # algorithmic 2*MAC of conv/convT/depthwise/dense per batch


class Net:
    """Builder.  With ``weights=None`` it only records the weight spec (name -> shape/init)."""

    def __init__(self, model, batch, weights=None, umma=True, keep_f32=False):
        self.plan = Plan(model, batch)
        self.split_weights = False  # see _conv_op
        self.w = weights
        self.spec = {}  # name -> (shape, init)
        self.keras_layers = []  # (Keras class, [weight keys in Keras' get_weights() order]) in construction order
        self._declared = set()
        self._declared_layers = set()
        self.umma = umma
        self.keep_f32 = keep_f32  # tests only: also keep the unquantised conv weights in the op

    # ------------------------------------------------------------------ weights
    def _get(self, name, shape, init):
        shape = tuple(int(s) for s in shape)
        if name in self._declared:  # registered ahead of use by declare_conv()
            assert self.spec[name] == (shape, init), (name, self.spec[name], shape, init)
            self._declared.discard(name)
        else:
            assert name not in self.spec, f"duplicate weight {name}"
            self.spec[name] = (shape, init)
        if self.w is None:
            return np.zeros(shape, np.float32)
        a = np.asarray(self.w[name], dtype=np.float32)
        if a.shape != shape:
            raise ValueError(f"weight {name}: expected {shape}, got {a.shape}")
        return a

    def _bn(self, name, c):
        g = self._get(name + "/gamma", (c,), "bn_gamma")
        b = self._get(name + "/beta", (c,), "bn_beta")
        m = self._get(name + "/mean", (c,), "bn_mean")
        v = self._get(name + "/var", (c,), "bn_var")
        self._keras("BatchNormalization", [name + "/gamma", name + "/beta", name + "/mean", name + "/var"])
        scale = g / np.sqrt(v + BN_EPS)
        return scale.astype(np.float32), (b - m * scale).astype(np.float32)

    def _keras(self, cls, keys):
        """Record a Keras layer in construction order (unless declare_conv() already placed it)."""
        if keys[0] in self._declared_layers:
            self._declared_layers.discard(keys[0])
        else:
            self.keras_layers.append((cls, keys))

    def declare_conv(self, cin, name, cout, k=1, bn=False, he=False):
        """Register the weights of a Conv2D (+ BatchNormalization) NOW, although the fused op that uses them is
        emitted later: the weight spec / Keras layer list then follow the order in which the reference's code constructs
        its layers (tools/keras_trace.py), which is the order a Keras checkpoint is matched by.  Needed where a
        convolution carries its residual input and therefore has to be emitted after the branch the reference builds
        second (hrnet.py:28-38 shortcut, v3plus.py:185-194 block 1)."""
        for key, shape, init in ((name + "/k", (k, k, cin, cout), "he_normal" if he else "glorot_uniform"),
                                 (name + "/b", (cout,), "zeros")):
            self._get(key, shape, init)
            self._declared.add(key)
        self.keras_layers.append(("Conv2D", [name + "/k", name + "/b"]))
        self._declared_layers.add(name + "/k")
        if bn:
            bname = bn if isinstance(bn, str) else name + "_bn"
            keys = [bname + "/gamma", bname + "/beta", bname + "/mean", bname + "/var"]
            for key, init in zip(keys, ("bn_gamma", "bn_beta", "bn_mean", "bn_var")):
                self._get(key, (cout,), init)
                self._declared.add(key)
            self.keras_layers.append(("BatchNormalization", keys))
            self._declared_layers.add(keys[0])

    # ------------------------------------------------------------------ buffers
    def buf(self, H, W, C, dtype="f16", kind="map"):
        # fp16 maps wider than 64 channels get a pixel pitch that is a multiple of 64 channels (128 bytes), so that
        # every 64-channel TMA row is one aligned 128-byte line (728 -> 768: the Xception middle flow); slices stay C
        # wide, the padding channels are never written and stay zero.
        if kind == "map" and dtype == "f16" and C > 64 and C % 64:
            C = -(-C // 64) * 64
        b = Buf(len(self.plan.bufs), H, W, C, dtype, kind)
        self.plan.bufs.append(b)
        return b

    def new(self, H, W, C, dtype="f16"):
        return T(self.buf(H, W, C, dtype), 0, C)

    def vec(self, C):
        return V(self.buf(1, 1, C, "f32", "vec"))

    def input(self, H=512, W=512, C=3):
        """Network input.  All five networks start with a 3x3 convolution on the RGB tile (res34.py:50, scse.py:52
        stride 1; hrnet.py:168, v3plus.py:173 stride 2), so the input buffer holds the tile already im2col'ed for
        that convolution: fp16, at the stem's OUTPUT resolution, channel (kh*3+kw)*3+c = 255 * x[oy*s+kh-pt,
        ox*s+kw-pl, c] (zero outside the tile: TF 'same' padding), 27 channels padded to 32.  255*x = 2*pixel-255
        is exact in fp16 for 8-bit imagery normalised as predict.py:93 does; the stem weights carry the 1/255.
        The stem then is a 1x1 tensor-core convolution with K = 32.  The buffer is laid out when the stem conv is
        built (``conv`` sees it is fed by the plan input)."""
        assert C == 3
        t = T(self.buf(H, W, INPUT_C, "f16"), 0, INPUT_C, ctrue=27, wscale=1.0 / INPUT_SCALE)
        self.plan.input = t.buf.id
        self._input_hw = (H, W)
        self._input_used = False
        return t

    def _stem(self, x, name, cout, s, bn, act, out, he):
        """3x3 stem conv on the plan input, lowered to a 1x1 conv over the im2col'ed input buffer."""
        assert not self._input_used, "the plan input feeds exactly one (stem) convolution"
        self._input_used = True
        H, W = self._input_hw
        kern = self._get(name + "/k", (3, 3, 3, cout), "he_normal" if he else "glorot_uniform")
        bias = self._get(name + "/b", (cout,), "zeros")
        self._keras("Conv2D", [name + "/k", name + "/b"])
        w = kern.reshape(27, cout).T.copy()[None]  # (1, Cout, 27): column (kh*3+kw)*3+c
        if bn:
            sc, sh = self._bn(bn if isinstance(bn, str) else name + "_bn", cout)
            w = w * sc[None, :, None]
            bias = bias * sc + sh
        wp = np.zeros((1, cout, INPUT_C), np.float32)
        wp[:, :, :27] = w * np.float32(x.wscale)
        Ho, Wo = math.ceil(H / s), math.ceil(W / s)
        x.buf.H, x.buf.W = Ho, Wo  # the gather writes the input at the stem's output resolution
        self.plan.input_stride = s
        if out is None:
            out = self.new(Ho, Wo, cout)
        a = ACT_RELU if act == "relu" else ACT_NONE
        self._conv_op(x, wp, bias, [(0, 0)], 1, Ho, Wo, a, None, ACT_NONE, out, name=name, macs_per_pixel=cout * 27)
        return out

    def _emit(self, **op):
        self.plan.ops.append(op)

    # ------------------------------------------------------------------ CONV
    def _conv_op(self, x, w_tco, bias, taps, stride, Ho, Wo, act_pre, res, act_post, out,
                 out_scale=1, out_oy=0, out_ox=0, name="", macs_per_pixel=None, cout_true=None):
        """w_tco: (ntaps, Cout, Cin) fp32 (BN already folded)."""
        nt, cout, cin = w_tco.shape
        assert cin == x.C and len(taps) == nt
        if macs_per_pixel is None:
            macs_per_pixel = cout * cin * nt
        if self.split_weights and 2 * nt <= MAX_TAPS:
            # hi/lo split: W = fp16(W) + fp16(W - fp16(W)) as a second set of taps at the same offsets -- the weights
            # reach the tensor core to ~2^-22 instead of 2^-11 (HRNet's stem and layer1, DESIGN.md "Numerics")
            hi = h16_to_f32(to_h16(w_tco))
            w_tco = np.concatenate([hi, (w_tco - hi).astype(np.float32)], axis=0)
            taps = list(taps) + list(taps)
            nt *= 2
        if x.c0 == 0 and x.C > 64 and x.C % 64 and x.buf.C == -(-x.C // 64) * 64 and x.buf.dtype == "f16":
            # read the whole padded buffer (its tail channels are zero) with zero weights for the tail: the weight
            # rows become 128-byte aligned as well and every k-block is a full 64-channel chunk
            wp = np.zeros((nt, cout, x.buf.C), np.float32)
            wp[:, :, :cin] = w_tco
            w_tco, cin = wp, x.buf.C
            x = T(x.buf, 0, x.buf.C)
        assert out.C == cout and out.H == Ho * out_scale and out.W == Wo * out_scale, (name, out, Ho, Wo)
        if res is not None:
            assert res.C == cout and res.H == out.H and res.W == out.W and out_scale == 1
        path = "direct"
        if (self.umma and stride in (1, 2) and x.buf.dtype == "f16" and cin % 8 == 0 and x.c0 % 8 == 0
                and x.buf.C % 8 == 0 and Wo >= 8):
            if out.buf.dtype == "f16":
                if (cout % 8 == 0 and out.c0 % 8 == 0 and out.buf.C % 8 == 0
                        and (res is None or (res.c0 % 8 == 0 and res.buf.C % 8 == 0))):
                    path = "umma"
            elif cout <= 16 and res is None and out_scale == 1:
                path = "umma"  # fp32 logits / gate maps: one 16-column tile, direct stores
        # <= 16 true output channels and at most SMALL_MACS multiply-adds per pixel (BAM's 4- / 8-channel gate
        # convolutions, 1-channel gate logits, the 1x1 two-channel heads): HBM-bound CUDA-core kernel instead of
        # 16-column tensor-core tiles that spend their time in per-k-block bookkeeping
        if (self.umma and cout_true is not None and cout_true <= 16 and macs_per_pixel <= SMALL_MACS and stride == 1
                and out_scale == 1 and res is None and nt <= 9 and x.buf.dtype == "f16" and cin % 8 == 0 and x.c0 % 8 == 0
                and x.buf.C % 8 == 0 and (out.buf.dtype == "f32" or (cout % 8 == 0 and out.c0 % 8 == 0 and out.buf.C % 8 == 0))
                and not self.split_weights):
            path = "small"
        # algorithmic work: true (unpadded) channel counts
        self.plan.flops += 2 * self.plan.batch * Ho * Wo * macs_per_pixel
        self._emit(op=OP_CONV, name=name, path=path, x=x.ref(), y=out.ref(),
                   flops=2 * self.plan.batch * Ho * Wo * macs_per_pixel,  # algorithmic (unpadded channel counts)
                   res=None if res is None else res.ref(),
                   taps=[(int(dy), int(dx)) for dy, dx in taps], stride=stride, Ho=Ho, Wo=Wo,
                   act_pre=act_pre, act_post=act_post,
                   out_scale=out_scale, out_oy=out_oy, out_ox=out_ox,
                   w=to_h16(w_tco), b=np.ascontiguousarray(bias, np.float32),
                   w32=np.ascontiguousarray(w_tco, np.float32) if self.keep_f32 else None)

    def conv(self, x, name, cout, k=1, s=1, d=1, bn=False, act=None, res=None, res_after_act=False,
             out=None, f32_out=False, he=False, cout_pad=None):
        """Conv2D(padding='same') [+BN] [+ReLU] [+res] [+ReLU].

        act applies to conv(+BN).  With ``res`` the result is ``act(conv + res)`` or, when
        ``res_after_act`` (res34's res_block1, res34.py:40-45), ``relu(act(conv) + res)``."""
        if x.buf.id == self.plan.input:
            assert k == 3 and d == 1 and res is None and not f32_out and not cout_pad
            return self._stem(x, name, cout, s, bn, act, out, he)
        kern = self._get(name + "/k", (k, k, x.cin, cout), "he_normal" if he else "glorot_uniform")
        bias = self._get(name + "/b", (cout,), "zeros")
        self._keras("Conv2D", [name + "/k", name + "/b"])
        w = kern.reshape(k * k, x.cin, cout).transpose(0, 2, 1).copy()  # (taps, Cout, Cin)
        if bn:
            sc, sh = self._bn(bn if isinstance(bn, str) else name + "_bn", cout)
            w = w * sc[None, :, None]
            bias = bias * sc + sh
        w = w * np.float32(x.wscale)
        cout_true = cout
        if x.cin != x.C or cout_pad:  # zero weights for padded input channels / padded output channels
            cout = max(cout, cout_pad or 0)
            wp = np.zeros((k * k, cout, x.C), np.float32)
            wp[:, :cout_true, :x.cin] = w
            w = wp
            bias = np.concatenate([bias, np.zeros(cout - cout_true, np.float32)])
        Ho, Wo = math.ceil(x.H / s), math.ceil(x.W / s)
        pt, _ = same_pad(x.H, k, s, d)
        pl, _ = same_pad(x.W, k, s, d)
        taps = [(kh * d - pt, kw * d - pl) for kh in range(k) for kw in range(k)]
        if out is None:
            out = self.new(Ho, Wo, cout, "f32" if f32_out else "f16")
        a = ACT_RELU if act == "relu" else ACT_NONE
        if res is None:
            act_pre, act_post = a, ACT_NONE
        elif res_after_act:
            act_pre, act_post = a, ACT_RELU
        else:
            act_pre, act_post = ACT_NONE, a
        self._conv_op(x, w, bias, taps, s, Ho, Wo, act_pre, res, act_post, out, name=name,
                      macs_per_pixel=cout_true * x.cin * k * k, cout_true=cout_true)
        if cout != cout_true:
            out = T(out.buf, out.c0, out.C, ctrue=cout_true)
        return out

    def conv_up2(self, x, name, cout, bn=False, act=None, out=None):
        """Conv2D(3x3, 'same')(UpSampling2D(2, nearest)(x)) [+BN] [+ReLU] without materialising the up-sampled map:
        every output pixel of sub-pixel phase (py, px) sees only a 2x2 neighbourhood of x, so the 3x3 kernel collapses
        into four 2x2 kernels of summed taps (source row of up-sampled row 2i+py+kh-1 is i + floor((py+kh-1)/2):
        py=0 -> {i-1: W0, i: W1+W2}, py=1 -> {i: W0+W1, i+1: W2}; same for columns).  The zero padding of the
        up-sampled map coincides with the zero padding of x.  4/9 of the multiply-adds, and neither the write nor
        the nine-fold re-read of the 4x larger map (hrnet.py:198-199, v3plus.py:341-342)."""
        kern = self._get(name + "/k", (3, 3, x.cin, cout), "glorot_uniform")
        bias = self._get(name + "/b", (cout,), "zeros")
        self._keras("Conv2D", [name + "/k", name + "/b"])
        w = kern.reshape(9, x.cin, cout).transpose(0, 2, 1).copy()  # (taps, Cout, Cin)
        if bn:
            sc, sh = self._bn(bn if isinstance(bn, str) else name + "_bn", cout)
            w = w * sc[None, :, None]
            bias = bias * sc + sh
        w = w * np.float32(x.wscale)
        if x.cin != x.C:
            wp = np.zeros((9, cout, x.C), np.float32)
            wp[:, :, :x.cin] = w
            w = wp
        w = w.reshape(3, 3, cout, x.C)
        if out is None:
            out = self.new(2 * x.H, 2 * x.W, cout)
        a = ACT_RELU if act == "relu" else ACT_NONE
        groups = {0: [(-1, (0,)), (0, (1, 2))], 1: [(0, (0, 1)), (1, (2,))]}  # phase -> [(source offset, summed taps)]
        for py in (0, 1):
            for px in (0, 1):
                taps, ws = [], []
                for dy, khs in groups[py]:
                    for dx, kws in groups[px]:
                        taps.append((dy, dx))
                        ws.append(sum(w[kh, kw] for kh in khs for kw in kws))
                self._conv_op(x, np.stack(ws), bias, taps, 1, x.H, x.W, a, None, ACT_NONE, out,
                              out_scale=2, out_oy=py, out_ox=px, name=f"{name}[{py}{px}]",
                              macs_per_pixel=cout * x.cin * 9)
        return out

    def conv_transpose(self, x, name, cout, k, act=None, out=None):
        """Conv2DTranspose(k in {2,3}, strides=2, padding='same') as 4 sub-pixel convs:
        y[2m+a] = sum over kh with kh = a (mod 2) of x[m - (kh-a)/2] W[kh]."""
        kern = self._get(name + "/k", (k, k, cout, x.C), "glorot_uniform")  # (kh,kw,Cout,Cin)
        bias = self._get(name + "/b", (cout,), "zeros")
        self._keras("Conv2DTranspose", [name + "/k", name + "/b"])
        if out is None:
            out = self.new(2 * x.H, 2 * x.W, cout)
        a = ACT_RELU if act == "relu" else ACT_NONE
        for py in (0, 1):
            for px in (0, 1):
                taps, ws = [], []
                for kh in range(py, k, 2):
                    for kw in range(px, k, 2):
                        taps.append((-(kh - py) // 2, -(kw - px) // 2))
                        ws.append(kern[kh, kw])  # (Cout, Cin)
                self._conv_op(x, np.stack(ws), bias, taps, 1, x.H, x.W, a, None, ACT_NONE, out,
                              out_scale=2, out_oy=py, out_ox=px, name=f"{name}[{py}{px}]")
        return out

    def sepconv(self, x, name, cout, s=1, relu_in=False, bn=True, act=None, res=None, out=None):
        """[ReLU ->] SeparableConv2D(3x3, 'same', strides=s) [+BN] [+ReLU] [+res]."""
        dw = self._get(name + "/dw", (3, 3, x.C, 1), "glorot_uniform")
        pw = self._get(name + "/pw", (1, 1, x.C, cout), "glorot_uniform")
        bias = self._get(name + "/b", (cout,), "zeros")
        self._keras("SeparableConv2D", [name + "/dw", name + "/pw", name + "/b"])
        Ho, Wo = math.ceil(x.H / s), math.ceil(x.W / s)
        mid = self.new(Ho, Wo, x.C)
        pt, _ = same_pad(x.H, 3, s)
        pl, _ = same_pad(x.W, 3, s)
        self.plan.flops += 2 * self.plan.batch * Ho * Wo * x.C * 9
        self._emit(op=OP_DWCONV, name=name + "/dw", x=x.ref(), y=mid.ref(), stride=s, pad_t=pt, pad_l=pl,
                   relu_in=int(relu_in), w=h16_to_f32(to_h16(dw.reshape(9, x.C))),  # fp16 weights like every conv
                   w32=np.ascontiguousarray(dw.reshape(9, x.C), np.float32) if self.keep_f32 else None)
        w = pw.reshape(1, x.C, cout).transpose(0, 2, 1).copy()
        if bn:
            sc, sh = self._bn(name + "_bn", cout)
            w = w * sc[None, :, None]
            bias = bias * sc + sh
        if out is None:
            out = self.new(Ho, Wo, cout)
        a = ACT_RELU if act == "relu" else ACT_NONE
        i_dw = len(self.plan.ops) - 1
        self._conv_op(mid, w, bias, [(0, 0)], 1, Ho, Wo, ACT_NONE if res is not None else a, res,
                      a if res is not None else ACT_NONE, out, name=name + "/pw")
        pw_op = self.plan.ops[-1]
        # One kernel for both stages (runtime.NativePlan, bd_conv_desc::dw_w_host) when the pointwise stage runs on
        # the tensor cores with full 64-channel k-blocks: stride 1, map at least 8 x 16, and the depthwise input
        # readable at the width the pointwise stage reads the intermediate (padded buffers have zero tails).
        cin_k = pw_op["x"][2]
        # Only with a single N tile (Cout <= 256): every N tile of a pixel tile would recompute the depthwise stage,
        # and the shared-memory pipe (TMA writes + tensor-core operand reads + depthwise window reads) is what bounds
        # the fused kernel -- measured on B200: 728->728 @32^2 55 us fused vs 22 + 29 us as two kernels, but
        # 128->128 @256^2 156 vs 280 us, 256->256 @128^2 84 vs 150 us.
        if (s == 1 and cout <= 256 and pw_op["path"] == "umma" and cin_k % 64 == 0 and x.W >= 8 and x.H >= 16
                and out.buf.dtype == "f16" and x.buf.dtype == "f16"
                and (cin_k == x.C or (x.c0 == 0 and x.buf.C == cin_k))):
            self.plan.ops[i_dw]["fuse"] = True
            pw_op["fused_dw"] = i_dw
        return out

    # ------------------------------------------------------------------ memory-bound ops
    def maxpool(self, x, k, s, same=False, out=None):
        if same:
            Ho, Wo = math.ceil(x.H / s), math.ceil(x.W / s)
            pt, _ = same_pad(x.H, k, s)
            pl, _ = same_pad(x.W, k, s)
        else:
            Ho, Wo = (x.H - k) // s + 1, (x.W - k) // s + 1
            pt = pl = 0
        if out is None:
            out = self.new(Ho, Wo, x.C)
        assert (out.H, out.W, out.C) == (Ho, Wo, x.C)
        self._emit(op=OP_MAXPOOL, x=x.ref(), y=out.ref(), k=k, stride=s, pad_t=pt, pad_l=pl)
        return out

    def addn(self, terms, act=None, out=None):
        """out = act(sum_i nearest_upsample(x_i, f_i)); terms = [(T, f), ...], at most 4."""
        x0, f0 = terms[0]
        H, W, C = x0.H * f0, x0.W * f0, x0.C
        for x, f in terms:
            assert (x.H * f, x.W * f, x.C) == (H, W, C)
        if out is None:
            out = self.new(H, W, C)
        assert (out.H, out.W, out.C) == (H, W, C) and len(terms) <= 4
        self._emit(op=OP_ADDN, xs=[x.ref() for x, _ in terms], fs=[int(f) for _, f in terms], y=out.ref(),
                   act=ACT_RELU if act == "relu" else ACT_NONE)
        return out

    def upsample(self, x, f, out=None):
        return self.addn([(x, f)], out=out)

    def copy(self, x, out):
        return self.addn([(x, 1)], out=out)

    def gap(self, x):
        v = self.vec(x.C)
        self._emit(op=OP_GAP, x=x.ref(), y=v.buf.id)
        return v

    def dense(self, vs, name, cout, bn=None, act=None, conv_kernel=False):
        """y = act(BN(W . sum(vs) + b)).  conv_kernel: the reference layer is a 1x1 Conv2D on a
        (1,1,C) map (kernel shape (1,1,in,out)) rather than a Dense layer."""
        cin = vs[0].C
        shape = (1, 1, cin, cout) if conv_kernel else (cin, cout)
        kern = self._get(name + "/k", shape, "glorot_uniform").reshape(cin, cout)
        bias = self._get(name + "/b", (cout,), "zeros")
        self._keras("Conv2D" if conv_kernel else "Dense", [name + "/k", name + "/b"])
        w = kern.T.copy()  # (Cout, Cin)
        if bn:
            sc, sh = self._bn(bn, cout)
            w = w * sc[:, None]
            bias = bias * sc + sh
        y = self.vec(cout)
        self.plan.flops += 2 * self.plan.batch * cin * cout
        self._emit(op=OP_DENSE, xs=[v.buf.id for v in vs], y=y.buf.id,
                   act={None: ACT_NONE, "relu": ACT_RELU, "sigmoid": ACT_SIGMOID}[act],
                   w=np.ascontiguousarray(w, np.float32), b=np.ascontiguousarray(bias, np.float32))
        return y

    def gate_se(self, x, v, out=None):
        """x * v[n,c] (res34 attention_demo, res34.py:102-104; v already sigmoid-ed)."""
        if out is None:
            out = self.new(x.H, x.W, x.C)
        self._emit(op=OP_GATE, mode=GATE_SE, x=x.ref(), y=out.ref(), v=v.buf.id, s=None, w=None, b=0.0)
        return out

    def scse(self, x, name, out=None):
        """scSE block (scse.py:20-46, v3plus.py:141-167): x*sigmoid(conv1x1->1(x)) + x*sigmoid(W2 W1 GAP(x))."""
        c = x.C
        ks = self._get(name + "_s/k", (1, 1, c, 1), "glorot_uniform").reshape(c)
        bs = self._get(name + "_s/b", (1,), "zeros")
        self._keras("Conv2D", [name + "_s/k", name + "_s/b"])
        g = self.gap(x)
        h = self.dense([g], name + "_c1", c // 16, conv_kernel=True)
        cs = self.dense([h], name + "_c2", c, act="sigmoid", conv_kernel=True)
        if out is None:
            out = self.new(x.H, x.W, c)
        self.plan.flops += 2 * self.plan.batch * x.H * x.W * c
        self._emit(op=OP_GATE, mode=GATE_SCSE, x=x.ref(), y=out.ref(), v=cs.buf.id, s=None,
                   w=np.ascontiguousarray(ks, np.float32), b=float(bs[0]))
        return out

    def gate_bam(self, x, cg, sg, out=None):
        """x * (1 + sigmoid(cg[n,c] + sg[n,h,w]))  (bam.py:57-71)."""
        if out is None:
            out = self.new(x.H, x.W, x.C)
        assert sg.C == 1 and (sg.H, sg.W) == (x.H, x.W)
        self._emit(op=OP_GATE, mode=GATE_BAM, x=x.ref(), y=out.ref(), v=cg.buf.id, s=sg.ref(), w=None, b=0.0)
        return out

    def skfuse(self, ds, g, logits, bn_name, out=None):
        """Selective-kernel fusion (v3plus.py:114-136): softmax over the five 256-vectors, weighted
        sum of the four maps and the broadcast pooled branch, then BN + ReLU."""
        c = ds[0].C
        sc, sh = self._bn(bn_name, c)
        if out is None:
            out = self.new(ds[0].H, ds[0].W, c)
        self._emit(op=OP_SKFUSE, xs=[d.ref() for d in ds], g=g.buf.id, logits=[l.buf.id for l in logits],
                   y=out.ref(), scale=sc, shift=sh)
        return out

    def bcast(self, v, out):
        """UpSampling2D of a 1x1 map: broadcast v[n,c] over the slice ``out``."""
        assert out.C == v.C
        self._emit(op=OP_BCAST, v=v.buf.id, y=out.ref())
        return out

    def softmax_head(self, logits, up=1):
        """2-class softmax (+ nearest upsample of the logits by ``up``, which commutes with the
        1x1 conv + softmax head of bam.py:332-333)."""
        assert logits.buf.dtype == "f32" and logits.C == 2 and logits.c0 == 0
        self.plan.logits = logits.buf.id
        self.plan.logits_up = up
        self._emit(op=OP_SOFTMAX2, x=logits.ref(), up=up)


# ---------------------------------------------------------------------- weight init
def init_weights(spec, seed=0, randomize_bn=False):
    """Keras-default initialisation of a weight spec: glorot_uniform / he_normal kernels, zero
    biases, identity BatchNorm -- or, for parity tests, randomised BN statistics
    (gamma~U(.5,1.5), beta~N(0,.1), mean~N(0,.1), var~U(.5,1.5))."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, (shape, init) in spec.items():
        if init in ("glorot_uniform", "he_normal"):
            if len(shape) == 4:
                rf = shape[0] * shape[1]
                # Conv2D kernels are (kh,kw,in,out); Conv2DTranspose (kh,kw,out,in): Keras computes
                # fans as shape[-2]*rf and shape[-1]*rf either way.
                fan_in, fan_out = shape[2] * rf, shape[3] * rf
            else:
                fan_in, fan_out = shape
            if init == "glorot_uniform":
                lim = math.sqrt(6.0 / (fan_in + fan_out))
                a = rng.uniform(-lim, lim, shape)
            else:  # he_normal: truncated normal, stddev sqrt(2/fan_in)/.8796
                std = math.sqrt(2.0 / fan_in) / 0.87962566103423978
                a = np.clip(rng.standard_normal(shape), -2, 2) * std
        elif init == "zeros":
            a = np.zeros(shape) if not randomize_bn else rng.normal(0, 0.05, shape)
        elif init == "bn_gamma":
            a = rng.uniform(0.5, 1.5, shape) if randomize_bn else np.ones(shape)
        elif init == "bn_beta":
            a = rng.normal(0, 0.1, shape) if randomize_bn else np.zeros(shape)
        elif init == "bn_mean":
            a = rng.normal(0, 0.1, shape) if randomize_bn else np.zeros(shape)
        elif init == "bn_var":
            a = rng.uniform(0.5, 1.5, shape) if randomize_bn else np.ones(shape)
        else:
            raise ValueError(init)
        out[name] = a.astype(np.float32)
    return out


def count_params(spec, prefix=None):
    return sum(int(np.prod(s)) for n, (s, _) in spec.items() if prefix is None or prefix(n))
