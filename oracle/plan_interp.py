"""ORACLE (test infrastructure only): CPU interpreter for the product's *plan* ops.

It executes a ``graph.Plan`` with torch on the CPU, reproducing what each sm_100a kernel is
specified to compute -- fp16 storage of feature maps, fp16 conv weights, fp32 accumulation,
fp32 vectors -- so that
  (1) the host-side lowering (fusions, concat slices, sub-pixel transposed convs, TF padding)
      can be checked against oracle/nets.py without a GPU, and
  (2) GPU kernels can be checked against a fp16-faithful expectation with a tight tolerance.
It mirrors the op semantics documented in include/bd_b200.h; it is never used by the product.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from building_detection_b200 import graph as G


def _q(t):  # fp16 storage round trip (saturating, like the device stores)
    return t.clamp(-G.H16_MAX, G.H16_MAX).to(torch.float16).to(torch.float32)


def _act(t, a):
    if a == G.ACT_RELU:
        return F.relu(t)
    if a == G.ACT_SIGMOID:
        return torch.sigmoid(t)
    return t


class Interp:
    def __init__(self, plan, emulate_h16=True):
        self.p = plan
        self.emu = emulate_h16
        n = plan.batch
        self.b = []
        for b in plan.bufs:
            if b.kind == "vec":
                self.b.append(torch.zeros(n, b.C))
            else:
                self.b.append(torch.zeros(n, b.H, b.W, b.C))

    def _store(self, ref, val):
        bid, c0, c = ref
        if self.emu and self.p.bufs[bid].dtype == "f16":
            val = _q(val)
        self.b[bid][..., c0:c0 + c] = val

    def _load(self, ref):
        bid, c0, c = ref
        return self.b[bid][..., c0:c0 + c]

    # -------------------------------------------------------------- ops
    def _conv(self, op):
        x = self._load(op["x"]).permute(0, 3, 1, 2)  # NCHW
        n, cin, H, W = x.shape
        if (not self.emu or op.get("w_exact")) and op.get("w32") is not None:
            w = torch.from_numpy(op["w32"])  # (study switch w_exact: fp32 weights with fp16 activations)
        else:
            w = torch.from_numpy(G.h16_to_f32(op["w"]).copy())  # (taps, Cout, Cin)
        s, Ho, Wo = op["stride"], op["Ho"], op["Wo"]
        acc = torch.zeros(n, w.shape[1], Ho, Wo)
        for t, (dy, dx) in enumerate(op["taps"]):
            # gather x[h*s+dy, w*s+dx] with zero fill out of bounds
            pt, pl = max(0, -dy), max(0, -dx)
            pb = max(0, (Ho - 1) * s + dy - (H - 1))
            pr = max(0, (Wo - 1) * s + dx - (W - 1))
            xp = F.pad(x, (pl, pr, pt, pb))
            y0, x0 = dy + pt, dx + pl
            patch = xp[:, :, y0:y0 + (Ho - 1) * s + 1:s, x0:x0 + (Wo - 1) * s + 1:s]
            acc += torch.einsum("nchw,oc->nohw", patch, w[t])
        acc += torch.from_numpy(op["b"]).view(1, -1, 1, 1)
        acc = _act(acc, op["act_pre"])
        if op["res"] is not None:
            acc = acc + self._load(op["res"]).permute(0, 3, 1, 2)
        acc = _act(acc, op["act_post"]).permute(0, 2, 3, 1)
        sc = op["out_scale"]
        if sc == 1:
            self._store(op["y"], acc)
        else:
            bid, c0, c = op["y"]
            if self.emu and self.p.bufs[bid].dtype == "f16":
                acc = _q(acc)
            self.b[bid][:, op["out_oy"]::sc, op["out_ox"]::sc, c0:c0 + c] = acc

    def _dwconv(self, op):
        x = self._load(op["x"]).permute(0, 3, 1, 2)
        if op["relu_in"]:
            x = F.relu(x)
        c = x.shape[1]
        s = op["stride"]
        H, W = x.shape[2], x.shape[3]
        Ho, Wo = -(-H // s), -(-W // s)
        pt, pl = op["pad_t"], op["pad_l"]
        pb = max(0, (Ho - 1) * s + 3 - pt - H)
        pr = max(0, (Wo - 1) * s + 3 - pl - W)
        xp = F.pad(x, (pl, pr, pt, pb))
        if not self.emu:
            wsrc = op["w32"] if op.get("w32") is not None else op["w"]
            w = torch.from_numpy(np.ascontiguousarray(wsrc)).view(3, 3, c).permute(2, 0, 1).unsqueeze(1).contiguous()
            y = F.conv2d(xp, w, None, stride=s, groups=c)
        else:
            # the device kernel accumulates the nine taps with packed half2 FMAs: fp16 accumulator, one rounding
            # per tap, kh-major order.  x*w + acc is exact in float64, so one rounding to fp16 reproduces the FMA.
            w = torch.from_numpy(np.ascontiguousarray(op["w"])).view(3, 3, c).double()
            acc = torch.zeros(x.shape[0], c, Ho, Wo, dtype=torch.float64)
            for kh in range(3):
                for kw in range(3):
                    tap = xp[:, :, kh:kh + (Ho - 1) * s + 1:s, kw:kw + (Wo - 1) * s + 1:s].double()
                    acc = (acc + tap * w[kh, kw].view(1, -1, 1, 1)).to(torch.float16).double()
            y = acc.float()
        self._store(op["y"], y.permute(0, 2, 3, 1))

    def _maxpool(self, op):
        x = self._load(op["x"]).permute(0, 3, 1, 2)
        k, s, pt, pl = op["k"], op["stride"], op["pad_t"], op["pad_l"]
        bid, c0, c = op["y"]
        Ho, Wo = self.p.bufs[bid].H, self.p.bufs[bid].W
        pb = max(0, (Ho - 1) * s + k - pt - x.shape[2])
        pr = max(0, (Wo - 1) * s + k - pl - x.shape[3])
        y = F.max_pool2d(F.pad(x, (pl, pr, pt, pb), value=float("-inf")), k, s)
        self._store(op["y"], y[:, :, :Ho, :Wo].permute(0, 2, 3, 1))

    def _addn(self, op):
        acc = None
        for ref, f in zip(op["xs"], op["fs"]):
            t = self._load(ref)
            if f > 1:
                t = t.repeat_interleave(f, dim=1).repeat_interleave(f, dim=2)
            acc = t if acc is None else acc + t
        self._store(op["y"], _act(acc, op["act"]))

    def _gap(self, op):
        self.b[op["y"]] = self._load(op["x"]).mean(dim=(1, 2))

    def _dense(self, op):
        x = sum(self.b[i] for i in op["xs"])
        y = x @ torch.from_numpy(op["w"]).T + torch.from_numpy(op["b"])
        self.b[op["y"]] = _act(y, op["act"])

    def _gate(self, op):
        x = self._load(op["x"])
        v = self.b[op["v"]][:, None, None, :]
        if op["mode"] == G.GATE_SE:
            y = x * v
        elif op["mode"] == G.GATE_SCSE:
            s = torch.sigmoid((x * torch.from_numpy(op["w"])).sum(-1, keepdim=True) + op["b"])
            y = x * s + x * v
        else:
            sg = self._load(op["s"])
            y = x * (1.0 + torch.sigmoid(v + sg))
        self._store(op["y"], y)

    def _skfuse(self, op):
        lg = torch.stack([self.b[i] for i in op["logits"]], dim=1)  # (N,5,C)
        sm = torch.softmax(lg, dim=1)
        acc = self.b[op["g"]][:, None, None, :] * sm[:, 4][:, None, None, :]
        for i, ref in enumerate(op["xs"]):
            acc = acc + self._load(ref) * sm[:, i][:, None, None, :]
        y = F.relu(acc * torch.from_numpy(op["scale"]) + torch.from_numpy(op["shift"]))
        self._store(op["y"], y)

    def _bcast(self, op):
        bid, c0, c = op["y"]
        self._store(op["y"], self.b[op["v"]][:, None, None, :].expand(-1, self.p.bufs[bid].H, self.p.bufs[bid].W, -1))

    def set(self, buf, arr):
        """Write a whole buffer (fp16 maps are rounded like a device store would)."""
        t = torch.as_tensor(np.asarray(arr), dtype=torch.float32)
        if self.emu and self.p.bufs[buf].kind == "map" and self.p.bufs[buf].dtype == "f16":
            t = _q(t)
        self.b[buf] = t.clone()

    def get(self, buf):
        return self.b[buf].numpy()

    def run(self, x_nhwc=None):
        if x_nhwc is not None:
            self.b[self.p.input] = im2col_input(x_nhwc, self.p.input_stride, self.emu)
        disp = {G.OP_CONV: self._conv, G.OP_DWCONV: self._dwconv, G.OP_MAXPOOL: self._maxpool,
                G.OP_ADDN: self._addn, G.OP_GAP: self._gap, G.OP_DENSE: self._dense, G.OP_GATE: self._gate,
                G.OP_SKFUSE: self._skfuse, G.OP_BCAST: self._bcast}
        probs = None
        for op in self.p.ops:
            if op["op"] == G.OP_SOFTMAX2:
                lg = self._load(op["x"])
                f = op["up"]
                if f > 1:
                    lg = lg.repeat_interleave(f, dim=1).repeat_interleave(f, dim=2)
                probs = torch.softmax(lg, dim=-1)
            else:
                disp[op["op"]](op)
        return None if probs is None else probs.numpy()


def im2col_input(x_nhwc, stride, emulate_h16=True):
    """The product's network-input layout (graph.Net.input): 255 * x, im2col'ed for the 3x3 stem conv with TF
    'same' padding at the stem's output resolution, channel (kh*3+kw)*3+c, padded to graph.INPUT_C channels."""
    x = torch.as_tensor(np.asarray(x_nhwc), dtype=torch.float32) * G.INPUT_SCALE
    n, H, W, _ = x.shape
    pt, pb = G.same_pad(H, 3, stride)
    pl, pr = G.same_pad(W, 3, stride)
    xp = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb)).permute(0, 2, 3, 1)
    Ho, Wo = -(-H // stride), -(-W // stride)
    cols = [xp[:, kh:kh + (Ho - 1) * stride + 1:stride, kw:kw + (Wo - 1) * stride + 1:stride, :]
            for kh in range(3) for kw in range(3)]
    out = torch.zeros(n, Ho, Wo, G.INPUT_C)
    out[..., :27] = torch.cat(cols, dim=-1)
    return _q(out) if emulate_h16 else out


def run_plan(plan, x_nhwc, emulate_h16=True):
    with torch.no_grad():
        return Interp(plan, emulate_h16).run(x_nhwc)
