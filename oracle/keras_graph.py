"""ORACLE (test infrastructure only): executes the graph the reference's own model-building code produced
(tests/golden/keras_graph_<model>.json, recorded by tools/keras_trace.py under a stand-in for tensorflow) with plain
per-layer semantics in torch fp32.  Unlike oracle/nets.py nothing about the STRUCTURE is transcribed by hand here --
which layer feeds which, in what order, with which arguments comes from the reference's code; only the meaning of each
Keras layer is restated.  tests/test_keras_graph.py holds oracle/nets.py and the product's
lowering to this."""
from __future__ import annotations

import json
import math
import os

import numpy as np
import torch
import torch.nn.functional as F

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_graph(model):
    with open(os.path.join(GOLDEN, f"keras_graph_{model}.json")) as f:
        return json.load(f)


def _same_pad(size, k, s, d=1):
    k_eff = (k - 1) * d + 1
    out = math.ceil(size / s)
    total = max((out - 1) * s + k_eff - size, 0)
    return total // 2, total - total // 2


def _act(x, a):
    if a in (None, "linear"):
        return x
    if a == "relu":
        return F.relu(x)
    if a == "sigmoid":
        return torch.sigmoid(x)
    if a == "softmax":
        return torch.softmax(x, dim=-1)
    raise NotImplementedError(a)


def _conv(x, cfg, kernel, bias, groups=1):
    """x NHWC; kernel HWIO"""
    kh, kw = cfg["kernel"]
    sh, sw = cfg["strides"]
    dh, dw = cfg["dilation"]
    xc = x.permute(0, 3, 1, 2)
    if cfg["padding"] == "same":
        pt, pb = _same_pad(xc.shape[2], kh, sh, dh)
        pl, pr = _same_pad(xc.shape[3], kw, sw, dw)
        xc = F.pad(xc, (pl, pr, pt, pb))
    w = kernel.permute(3, 2, 0, 1).contiguous()
    return F.conv2d(xc, w, bias, stride=(sh, sw), dilation=(dh, dw), groups=groups).permute(0, 2, 3, 1)


def run(graph, weights, x_nhwc):
    """weights: {'<keras layer name>/<weight name>': ndarray}; returns the model output as numpy."""
    t = {graph["input"]: torch.as_tensor(np.asarray(x_nhwc), dtype=torch.float32)}
    layers = graph["layers"]

    def W(layer, name):
        return torch.from_numpy(np.ascontiguousarray(weights[f"{layer['name']}/{name}"], dtype=np.float32))
    for call in graph["calls"]:
        L = layers[call["layer"]]
        cls, cfg = L["class"], L["config"]
        ins = [t[i] for i in call["in"]]
        x = ins[0]
        if cls == "Conv2D":
            y = _act(_conv(x, cfg, W(L, "kernel:0"), W(L, "bias:0") if cfg["use_bias"] else None), cfg["activation"])
        elif cls == "SeparableConv2D":
            c = x.shape[-1]
            dwk = W(L, "depthwise_kernel:0")  # (kh,kw,C,1) -> per-channel filters
            y = _conv(x, cfg, dwk.permute(0, 1, 3, 2), None, groups=c)  # HWIO with I = 1, O = C
            y = _conv(y, {"kernel": [1, 1], "strides": [1, 1], "dilation": [1, 1], "padding": "valid"}, W(L, "pointwise_kernel:0"),
                      W(L, "bias:0") if cfg["use_bias"] else None)
            y = _act(y, cfg["activation"])
        elif cls == "Conv2DTranspose":
            k, s = cfg["kernel"][0], cfg["strides"][0]
            assert cfg["kernel"][0] == cfg["kernel"][1] and cfg["strides"] == [s, s] and cfg["padding"] == "same"
            kern = W(L, "kernel:0")  # (kh,kw,Cout,Cin); full transposed conv, no flip: y[s i + a] += x[i] W[a]
            full = F.conv_transpose2d(x.permute(0, 3, 1, 2), kern.permute(3, 2, 0, 1).contiguous(), None, stride=s)
            n_out = x.shape[1] * s
            # TF 'same': total padding max(k - s, 0), the smaller half in front
            p0 = max(k - s, 0) // 2
            y = full[:, :, p0:p0 + n_out, p0:p0 + x.shape[2] * s].permute(0, 2, 3, 1)
            if cfg["use_bias"]:
                y = y + W(L, "bias:0")
            y = _act(y, cfg["activation"])
        elif cls == "BatchNormalization":
            eps = cfg.get("epsilon", 1e-3)
            y = (x - W(L, "moving_mean:0")) / torch.sqrt(W(L, "moving_variance:0") + eps) * W(L, "gamma:0") + W(L, "beta:0")
        elif cls == "Dense":
            y = x @ W(L, "kernel:0")
            if cfg["use_bias"]:
                y = y + W(L, "bias:0")
            y = _act(y, cfg["activation"])
        elif cls == "Activation":
            y = _act(x, cfg["activation"])
        elif cls == "ReLU":
            y = F.relu(x)
        elif cls == "Softmax":
            y = torch.softmax(x, dim=cfg["axis"])
        elif cls in ("MaxPooling2D", "MaxPool2D", "AveragePooling2D"):
            xc = x.permute(0, 3, 1, 2)
            k, s = cfg["pool"], cfg["strides"]
            if cfg["padding"] == "same":
                pt, pb = _same_pad(xc.shape[2], k[0], s[0])
                pl, pr = _same_pad(xc.shape[3], k[1], s[1])
                assert cls != "AveragePooling2D"
                xc = F.pad(xc, (pl, pr, pt, pb), value=float("-inf"))
            y = (F.avg_pool2d if cls == "AveragePooling2D" else F.max_pool2d)(xc, tuple(k), tuple(s)).permute(0, 2, 3, 1)
        elif cls == "UpSampling2D":
            assert cfg["interpolation"] == "nearest"
            y = x.repeat_interleave(cfg["size"][0], dim=1).repeat_interleave(cfg["size"][1], dim=2)
        elif cls in ("GlobalAveragePooling2D", "GlobalAvgPool2D"):
            y = x.mean(dim=(1, 2))
        elif cls == "Reshape":
            y = x.reshape((x.shape[0],) + tuple(cfg["target"]))
        elif cls == "RepeatVector":
            y = x[:, None, :].expand(-1, cfg["n"], -1)
        elif cls == "Cropping2D":
            (ct, cb), (cl, cr) = cfg["cropping"]
            y = x[:, ct:x.shape[1] - cb, cl:x.shape[2] - cr]
        elif cls == "Concatenate":
            y = torch.cat(ins, dim=cfg["axis"])
        elif cls == "Add":
            y = ins[0]
            for o in ins[1:]:
                y = y + o
        elif cls == "Multiply":
            y = ins[0]
            for o in ins[1:]:
                y = y * o
        elif cls == "TFOp":
            op = cfg["op"]
            if op == "add":
                y = ins[0] + ins[1]
            elif op == "multiply":
                y = ins[0] * ins[1]
            elif op == "concat":
                y = torch.cat(ins, dim=cfg["axis"])
            elif op == "reshape":
                y = x.reshape([x.shape[0]] + [int(v) for v in cfg["shape"][1:]])
            else:
                raise NotImplementedError(op)
        else:
            raise NotImplementedError(cls)
        want = call["shape"]
        assert list(y.shape[1:]) == want[1:], (L["name"], cls, tuple(y.shape), want)
        t[call["out"]] = y
    return t[graph["output"]].numpy()
