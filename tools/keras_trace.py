"""Executes the reference's own model-building code (/root/reference/predict_model/*.py) under a RECORDING stand-in for
``tensorflow`` and writes what it did to tests/golden/keras_graph_<model>.json: every Keras layer in construction
order (class, Keras auto name, constructor arguments, weight shapes) and every call (inputs -> output, shapes).
TensorFlow is not installable here, but the graph structure is pure Python control flow around ``tf.keras.layers``
constructors, so this captures it exactly as the reference wrote it -- nothing is transcribed by hand.

The fixtures pin two things (tests/test_keras_graph.py):
  * oracle/nets.py (the hand restatement) computes the same function as the recorded graph executed by
    oracle/keras_graph.py with plain layer semantics, and
  * the Keras layer names / weight order a real ``.h5`` checkpoint carries map onto this package's weight names
    (building_detection_b200/keras_h5.py).
Run in the build container only (the reference tree is absent on the GPU box):  python tools/keras_trace.py"""
import importlib.util
import json
import math
import os
import re
import sys
import types

REF = "/root/reference/predict_model"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class Rec:
    def __init__(self):
        self.layers, self.calls, self.counts, self.ntensors = [], [], {}, 0

    def tensor(self, shape):
        t = Sym(self, self.ntensors, tuple(shape))
        self.ntensors += 1
        return t

    def auto_name(self, cls):
        base = re.sub(r"(?<!^)(?=[A-Z][a-z])|(?<=[a-z0-9])(?=[A-Z])", "_", cls).lower()
        base = {"conv2_d": "conv2d", "conv2_d_transpose": "conv2d_transpose", "separable_conv2_d": "separable_conv2d",
                "max_pooling2_d": "max_pooling2d", "average_pooling2_d": "average_pooling2d", "up_sampling2_d": "up_sampling2d",
                "global_average_pooling2_d": "global_average_pooling2d", "cropping2_d": "cropping2d", "re_lu": "re_lu",
                "zero_padding2_d": "zero_padding2d"}.get(base, base)
        k = self.counts.get(base, 0)
        self.counts[base] = k + 1
        return base if k == 0 else f"{base}_{k}"


REC = None


class Sym:
    """symbolic NHWC tensor"""

    def __init__(self, rec, tid, shape):
        self.rec, self.id, self.shape = rec, tid, shape

    def __getitem__(self, item):  # only used as w[i] on lists, never on tensors
        raise TypeError("tensor indexing is not used by the reference graphs")

    def __add__(self, o):
        return tf_op("add", [self, o])

    def __mul__(self, o):
        return tf_op("multiply", [self, o])


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def _same(n, k, s, d=1):
    return math.ceil(n / s)


class Layer:
    weighted = False

    def __init__(self, *args, name=None, **kw):
        self.cls = type(self).__name__
        self.cfg = self.config(*args, **kw)
        self.kname = name or REC.auto_name(self.ALIAS if hasattr(self, "ALIAS") else self.cls)
        self.index = len(REC.layers)
        self.weights = None
        REC.layers.append(self)

    def config(self, *a, **k):
        return {}

    def __call__(self, x):
        ins = x if isinstance(x, (list, tuple)) else [x]
        shape = self.out_shape([t.shape for t in ins])
        if self.weights is None:
            self.weights = self.weight_shapes(ins[0].shape)
        out = REC.tensor(shape)
        REC.calls.append({"layer": self.index, "in": [t.id for t in ins], "out": out.id, "shape": list(shape)})
        return out

    def out_shape(self, shapes):
        return shapes[0]

    def weight_shapes(self, shape):
        return []


class Conv2D(Layer):
    def config(self, filters, kernel_size, strides=(1, 1), padding="valid", dilation_rate=(1, 1), activation=None,
               use_bias=True, kernel_initializer="glorot_uniform", **kw):
        assert not kw or set(kw) <= {"bias_initializer", "kernel_regularizer"}, kw
        return {"filters": filters, "kernel": _pair(kernel_size), "strides": _pair(strides), "padding": padding,
                "dilation": _pair(dilation_rate), "activation": activation, "use_bias": use_bias, "init": str(kernel_initializer)}

    def out_shape(self, shapes):
        n, h, w, _ = shapes[0]
        c = self.cfg
        if c["padding"] == "same":
            return (n, _same(h, c["kernel"][0], c["strides"][0]), _same(w, c["kernel"][1], c["strides"][1]), c["filters"])
        ke = [(c["kernel"][i] - 1) * c["dilation"][i] + 1 for i in (0, 1)]
        return (n, (h - ke[0]) // c["strides"][0] + 1, (w - ke[1]) // c["strides"][1] + 1, c["filters"])

    def weight_shapes(self, shape):
        c = self.cfg
        w = [("kernel:0", [c["kernel"][0], c["kernel"][1], shape[-1], c["filters"]])]
        return w + ([("bias:0", [c["filters"]])] if c["use_bias"] else [])


class Conv2DTranspose(Conv2D):
    def out_shape(self, shapes):
        n, h, w, _ = shapes[0]
        c = self.cfg
        assert c["padding"] == "same"
        return (n, h * c["strides"][0], w * c["strides"][1], c["filters"])

    def weight_shapes(self, shape):
        c = self.cfg
        w = [("kernel:0", [c["kernel"][0], c["kernel"][1], c["filters"], shape[-1]])]
        return w + ([("bias:0", [c["filters"]])] if c["use_bias"] else [])


class SeparableConv2D(Conv2D):
    def weight_shapes(self, shape):
        c = self.cfg
        w = [("depthwise_kernel:0", [c["kernel"][0], c["kernel"][1], shape[-1], 1]),
             ("pointwise_kernel:0", [1, 1, shape[-1], c["filters"]])]
        return w + ([("bias:0", [c["filters"]])] if c["use_bias"] else [])


class BatchNormalization(Layer):
    def config(self, axis=-1, **kw):
        return {"axis": axis, **{k: v for k, v in kw.items() if k in ("epsilon", "momentum")}}

    def weight_shapes(self, shape):
        c = shape[-1]
        return [("gamma:0", [c]), ("beta:0", [c]), ("moving_mean:0", [c]), ("moving_variance:0", [c])]


class Dense(Layer):
    def config(self, units, activation=None, use_bias=True, **kw):
        return {"units": units, "activation": activation, "use_bias": use_bias}

    def out_shape(self, shapes):
        return tuple(shapes[0][:-1]) + (self.cfg["units"],)

    def weight_shapes(self, shape):
        w = [("kernel:0", [shape[-1], self.cfg["units"]])]
        return w + ([("bias:0", [self.cfg["units"]])] if self.cfg["use_bias"] else [])


class Activation(Layer):
    def config(self, activation):
        return {"activation": activation}


class ReLU(Layer):
    def config(self, **kw):
        assert not kw, kw
        return {}


class Softmax(Layer):
    def config(self, axis=-1):
        return {"axis": axis}


class MaxPooling2D(Layer):
    ALIAS = "MaxPooling2D"

    def config(self, pool_size=(2, 2), strides=None, padding="valid"):
        ps = _pair(pool_size)
        return {"pool": ps, "strides": _pair(strides) if strides is not None else ps, "padding": padding}

    def out_shape(self, shapes):
        n, h, w, c = shapes[0]
        k, s = self.cfg["pool"], self.cfg["strides"]
        if self.cfg["padding"] == "same":
            return (n, math.ceil(h / s[0]), math.ceil(w / s[1]), c)
        return (n, (h - k[0]) // s[0] + 1, (w - k[1]) // s[1] + 1, c)


class MaxPool2D(MaxPooling2D):  # Keras alias: the class (and its auto name) is MaxPooling2D
    pass


class AveragePooling2D(MaxPooling2D):
    ALIAS = "AveragePooling2D"


class UpSampling2D(Layer):
    def config(self, size=(2, 2), interpolation="nearest"):
        return {"size": _pair(size), "interpolation": interpolation}

    def out_shape(self, shapes):
        n, h, w, c = shapes[0]
        return (n, h * self.cfg["size"][0], w * self.cfg["size"][1], c)


class GlobalAveragePooling2D(Layer):
    ALIAS = "GlobalAveragePooling2D"

    def out_shape(self, shapes):
        return (shapes[0][0], shapes[0][-1])


class GlobalAvgPool2D(GlobalAveragePooling2D):
    pass


class Reshape(Layer):
    def config(self, target_shape):
        return {"target": list(target_shape)}

    def out_shape(self, shapes):
        return (shapes[0][0],) + tuple(self.cfg["target"])


class RepeatVector(Layer):
    def config(self, n):
        return {"n": n}

    def out_shape(self, shapes):
        return (shapes[0][0], self.cfg["n"], shapes[0][-1])


class Cropping2D(Layer):
    def config(self, cropping):
        return {"cropping": [list(c) for c in cropping]}

    def out_shape(self, shapes):
        n, h, w, c = shapes[0]
        (t, b), (l, r) = self.cfg["cropping"]
        return (n, h - t - b, w - l - r, c)


class Concatenate(Layer):
    def config(self, axis=-1):
        return {"axis": axis}

    def out_shape(self, shapes):
        ax = self.cfg["axis"] % len(shapes[0])
        s = list(shapes[0])
        s[ax] = sum(x[ax] for x in shapes)
        return tuple(s)


class Add(Layer):
    def out_shape(self, shapes):
        return _broadcast(shapes)


class Multiply(Layer):
    def out_shape(self, shapes):
        return _broadcast(shapes)


def _broadcast(shapes):
    rank = max(len(s) for s in shapes)
    out = [1] * rank
    for s in shapes:
        s = (1,) * (rank - len(s)) + tuple(s)
        for i, v in enumerate(s):
            if v is None or out[i] is None:
                out[i] = None
            elif v != 1:
                assert out[i] in (1, v), shapes
                out[i] = v
    return tuple(out)


class TFOp(Layer):
    """tf.add / tf.multiply / tf.concat / tf.reshape on Keras tensors (TFOpLambda layers in TF 2.4+: no weights)"""

    def __init__(self, op, **cfg):
        self._op, self._cfg = op, cfg
        super().__init__(name=REC.auto_name("tf_op_" + op))

    def config(self):
        return {"op": self._op, **self._cfg}

    def out_shape(self, shapes):
        op = self._op
        if op in ("add", "multiply"):
            return _broadcast(shapes)
        if op == "concat":
            ax = self._cfg["axis"] % len(shapes[0])
            s = list(shapes[0])
            s[ax] = sum(x[ax] for x in shapes)
            return tuple(s)
        if op == "reshape":
            tgt = list(self._cfg["shape"])
            known = 1
            for v in tgt[1:]:
                known *= v
            return (shapes[0][0],) + tuple(tgt[1:])
        raise NotImplementedError(op)


def tf_op(op, ins, **cfg):
    return TFOp(op, **cfg)(list(ins))


def make_stub():
    tf = types.ModuleType("tensorflow")
    keras = types.ModuleType("tensorflow.keras")
    layers = types.ModuleType("tensorflow.keras.layers")
    backend = types.ModuleType("tensorflow.keras.backend")
    models = types.ModuleType("tensorflow.keras.models")
    for cls in (Conv2D, Conv2DTranspose, SeparableConv2D, BatchNormalization, Dense, Activation, ReLU, Softmax, MaxPooling2D,
                MaxPool2D, AveragePooling2D, UpSampling2D, GlobalAveragePooling2D, GlobalAvgPool2D, Reshape, RepeatVector,
                Cropping2D, Concatenate, Add, Multiply):
        setattr(layers, cls.__name__, cls)

    def Input(shape=None, **kw):
        t = REC.tensor((None,) + tuple(shape))
        REC.input_id = t.id
        return t
    layers.Input = Input
    layers.add = lambda xs, **k: Add()(list(xs))
    layers.multiply = lambda xs, **k: Multiply()(list(xs))
    layers.concatenate = lambda xs, axis=-1, **k: Concatenate(axis=axis)(list(xs))

    class Model:
        def __init__(self, inputs=None, outputs=None, **kw):
            REC.output_id = outputs.id
            self.inputs, self.outputs = inputs, outputs
    keras.Model = models.Model = Model
    keras.Input = Input
    keras.layers, keras.backend, keras.models = layers, backend, models
    tf.keras = keras
    tf.add = lambda a, b, **k: tf_op("add", [a, b])
    tf.multiply = lambda a, b, **k: tf_op("multiply", [a, b])
    tf.concat = lambda xs, axis=-1, **k: tf_op("concat", list(xs), axis=axis)
    tf.reshape = lambda x, shape, **k: tf_op("reshape", [x], shape=[int(v) for v in shape])
    cfg = types.SimpleNamespace(experimental=types.SimpleNamespace(list_physical_devices=lambda *_: [],
                                                                   set_memory_growth=lambda *a: None))
    tf.config = cfg
    mods = {"tensorflow": tf, "tensorflow.keras": keras, "tensorflow.keras.layers": layers, "tensorflow.keras.backend": backend,
            "tensorflow.keras.models": models}
    return mods


BUILD = {
    "res34": lambda m: m.ResNetFamily(input_shape=(512, 512, 3)).run_model("res34"),
    "hrnet": lambda m: m.HRNet(shape=(512, 512, 3), num_classes=2),
    "v3plus": lambda m: m.Xception_DeepLabV3_Plus(shape=(512, 512, 3), num_classes=2),
    "scse": lambda m: m.UNet(2, input_shape=(512, 512, 3)),
    "bam": lambda m: m.Xception_DeepLabV3_Plus_bam(shape=(512, 512, 3), num_classes=2),
}


def trace(name):
    global REC
    REC = Rec()
    saved = {k: sys.modules.get(k) for k in ("tensorflow", "tensorflow.keras", "tensorflow.keras.layers",
                                             "tensorflow.keras.backend", "tensorflow.keras.models")}
    sys.modules.update(make_stub())
    try:
        spec = importlib.util.spec_from_file_location(f"_ref_{name}", os.path.join(REF, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        BUILD[name](mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    # keep only what the output depends on (the reference files build nothing else, but dead layers would carry no
    # weights into model.layers either)
    by_out = {c["out"]: c for c in REC.calls}
    live, stack = set(), [REC.output_id]
    while stack:
        t = stack.pop()
        if t in live or t not in by_out:
            continue
        live.add(t)
        stack.extend(by_out[t]["in"])
    calls = [c for c in REC.calls if c["out"] in live]
    used = {c["layer"] for c in calls}
    layers = [{"index": l.index, "class": l.cls, "name": l.kname, "config": l.cfg, "weights": l.weights or [],
               "live": l.index in used} for l in REC.layers]
    return {"model": name, "source": f"predict_model/{name}.py", "input": REC.input_id, "output": REC.output_id,
            "layers": layers, "calls": calls}


def main():
    os.makedirs(OUT, exist_ok=True)
    for name in BUILD:
        g = trace(name)
        nw = sum(1 for l in g["layers"] if l["weights"] and l["live"])
        params = sum(math.prod(s) for l in g["layers"] if l["live"] for _, s in l["weights"])
        dead = [l["name"] for l in g["layers"] if not l["live"]]
        path = os.path.join(OUT, f"keras_graph_{name}.json")
        with open(path, "w") as f:
            json.dump(g, f, separators=(",", ":"))
        print(f"{name}: {len(g['layers'])} layers ({nw} with weights, {len(dead)} dead), {len(g['calls'])} calls, "
              f"{params:,} parameters -> {os.path.relpath(path)} ({os.path.getsize(path) // 1024} KB)")


if __name__ == "__main__":
    main()
