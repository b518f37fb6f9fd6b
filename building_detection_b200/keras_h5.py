"""Keras ``.h5`` weight files <-> this package's weight dicts (the reference loads its checkpoints with
``model.load_weights('....h5')``, predict.py:21-49).

File layout (tf.keras ``save_weights`` in HDF5 format): root attribute ``layer_names``; one group per layer with
attribute ``weight_names`` (e.g. ``conv2d_7/kernel:0``) and the arrays below it.  ``load_weights`` without
``by_name`` matches layers by TOPOLOGY, not by name, and so does this loader: the file's weighted layers are ranked
per Keras class by the index in their auto-generated names (``conv2d``, ``conv2d_1``, ... -- creation order, whatever
the offset, since a process that built another model first only shifts the numbers) and matched to the layers this
package's builders record in the same creation order (``Model.keras_layers``).  That order is pinned to the
reference's own code: tools/keras_trace.py executes predict_model/*.py under a recording stand-in for tensorflow, and
tests/test_keras_graph.py checks class sequence, shapes and -- numerically -- the wiring.  Every array's shape is
checked; a mismatch raises ValueError naming the layer.
Arrays are stored in Keras layouts already (HWIO kernels, (kh,kw,Cout,Cin) transposed-conv kernels, (in,out) dense
kernels, BN gamma/beta/moving_mean/moving_variance), so no transposition happens here."""
from __future__ import annotations

import re

import numpy as np

from . import hdf5_min

CLASS_OF_PREFIX = {"conv2d": "Conv2D", "conv2d_transpose": "Conv2DTranspose", "separable_conv2d": "SeparableConv2D",
                   "batch_normalization": "BatchNormalization", "dense": "Dense"}
PREFIX_OF_CLASS = {v: k for k, v in CLASS_OF_PREFIX.items()}
KERAS_WEIGHT_NAMES = {"Conv2D": ["kernel:0", "bias:0"], "Conv2DTranspose": ["kernel:0", "bias:0"], "Dense": ["kernel:0", "bias:0"],
                      "SeparableConv2D": ["depthwise_kernel:0", "pointwise_kernel:0", "bias:0"],
                      "BatchNormalization": ["gamma:0", "beta:0", "moving_mean:0", "moving_variance:0"]}


def is_hdf5(path):
    with open(path, "rb") as f:
        return f.read(8) == hdf5_min.SIG


def _split(name):
    m = re.fullmatch(r"(.*?)(?:_(\d+))?", name)
    base, idx = m.group(1), m.group(2)
    if base not in CLASS_OF_PREFIX and idx is None:
        return name, 0
    return base, int(idx) if idx is not None else 0


def _decode(a):
    return [x.decode("utf-8") if isinstance(x, bytes) else str(x) for x in np.asarray(a).ravel().tolist()]


def read_layers(path):
    """[(keras layer name, [(weight name, ndarray), ...]), ...] in file order, weighted layers only."""
    f = hdf5_min.File(path)
    root = f["model_weights"] if "layer_names" not in f.attrs and "model_weights" in f else f  # full-model files nest them
    if root.attrs.get("layer_names") is None:
        names = root.keys()  # attribute too large for the object header (Keras splits it): fall back to the group listing
    else:
        names = _decode(root.attrs["layer_names"])
    out = []
    for ln in names:
        g = root[ln]
        wn = g.attrs.get("weight_names")
        wnames = _decode(wn) if wn is not None and np.asarray(wn).size else []
        if wnames:
            out.append((ln, [(w.split("/")[-1], np.asarray(g[w].read(), np.float32)) for w in wnames]))
    return out


def load_into(model, path):
    """Fill ``model`` (engine.Model) from a Keras .h5 weight file."""
    by_class = {}
    for ln, ws in read_layers(path):
        base, idx = _split(ln)
        cls = CLASS_OF_PREFIX.get(base)
        if cls is None:
            raise ValueError(f"{path}: layer {ln!r} of an unknown class carries weights")
        by_class.setdefault(cls, []).append((idx, ln, ws))
    ours = {}
    for cls, keys in model.keras_layers:
        ours.setdefault(cls, []).append(keys)
    weights = {}
    for cls, mine in ours.items():
        theirs = sorted(by_class.get(cls, []), key=lambda t: t[0])
        if len(theirs) != len(mine):
            raise ValueError(f"{path}: {len(theirs)} {cls} layers with weights, the {model.name} graph has {len(mine)}")
        for keys, (_idx, ln, ws) in zip(mine, theirs):
            if len(ws) != len(keys):
                raise ValueError(f"{path}: layer {ln} has {len(ws)} arrays, expected {len(keys)} ({keys[0]})")
            for key, (wn, arr) in zip(keys, ws):
                want = model.spec[key][0]
                if tuple(arr.shape) != tuple(want):
                    raise ValueError(f"{path}: {ln}/{wn} has shape {tuple(arr.shape)}, {model.name} expects {tuple(want)} for {key}")
                weights[key] = arr
    extra = set(by_class) - set(ours)
    if extra:
        raise ValueError(f"{path}: the file has {sorted(extra)} layers, the {model.name} graph has none")
    model.set_weights(weights)


def save(model, path):
    """Write ``model``'s weights as a Keras-format .h5 weight file (layer names as Keras would generate them in a
    fresh process), readable by ``load_into`` and by tf.keras ``load_weights``."""
    counts, tree, layer_names = {}, {}, []
    w = model.get_weights()
    for cls, keys in model.keras_layers:
        k = counts.get(cls, 0)
        counts[cls] = k + 1
        ln = PREFIX_OF_CLASS[cls] + (f"_{k}" if k else "")
        layer_names.append(ln)
        wn = [f"{ln}/{n}" for n in KERAS_WEIGHT_NAMES[cls][:len(keys)]]
        if cls in ("Conv2D", "Conv2DTranspose", "Dense") and len(keys) == 2:
            wn = [f"{ln}/kernel:0", f"{ln}/bias:0"]
        sub = {n.split("/")[-1]: np.asarray(w[key], np.float32) for n, key in zip(wn, keys)}
        tree[ln] = {"@weight_names": np.array([n.encode() for n in wn]), ln: sub}
    tree["@layer_names"] = np.array([n.encode() for n in layer_names])
    tree["@backend"] = np.array(b"tensorflow")
    tree["@keras_version"] = np.array(b"2.4.0")
    hdf5_min.write_file(path, tree)
