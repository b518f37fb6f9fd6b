"""Model object with the two Keras methods the reference's hot path uses (predict.py:21-49,109):
``load_weights(path)`` and ``predict(x)``.  ``predict`` runs the model's native plan -- a launch
list of sm_100a kernels built once per batch size through the C ABI -- on the current CUDA device.
There is no CPU execution path: without the compiled library or a GPU it raises."""
from __future__ import annotations

import os
import zlib

import numpy as np

from . import graph as G

MAX_LARGE_PLANS = 1  # per model and device
MAX_PLAN_BATCH = 32  # largest plan batch (BASELINE configs 2-4 run 16; the scene loop runs 32: +3.6 % per tile on B200)


class Model:
    def __init__(self, name, build_fn, input_shape=(512, 512, 3), seed=None):
        if tuple(input_shape) != (512, 512, 3):
            raise ValueError("the reference tiles every scene into 512x512x3 inputs (predict.py:107)")
        self.name = name
        self._build = build_fn
        net = G.Net(name, 1, None)
        build_fn(net)
        self.spec = net.spec
        self.keras_layers = net.keras_layers  # (Keras class, weight keys) in this builder's construction order
        self.flops_per_tile = net.plan.flops
        # a fresh Keras model carries its random initialisation (predict.py:23-24 keeps it when the
        # checkpoint is missing); seeded per model so runs are reproducible
        self.weights = G.init_weights(self.spec, seed=zlib.crc32(name.encode()) if seed is None else seed)
        self._native = {}  # batch -> engine.NativePlan
        self._version = 0

    # ------------------------------------------------------------------ weights
    def get_weights(self):
        return dict(self.weights)

    def set_weights(self, weights):
        for k, (shape, _) in self.spec.items():
            if k not in weights:
                raise KeyError(f"{self.name}: missing weight {k}")
            if tuple(np.shape(weights[k])) != shape:
                raise ValueError(f"{self.name}: weight {k} has shape {np.shape(weights[k])}, expected {shape}")
        self.weights = {k: np.asarray(weights[k], np.float32) for k in self.spec}
        self._drop_native()

    def load_weights(self, path):
        """Keras raises OSError for a missing or unreadable file and the reference catches exactly that
        (predict.py:23).  Containers: a Keras ``.h5`` weight file (what the reference's checkpoints are; matched by
        topology like ``load_weights`` does, keras_h5.py) or an ``.npz`` keyed by this package's weight names."""
        if not os.path.exists(path):
            raise OSError(f"Unable to open file (name = '{path}')")
        from . import keras_h5
        if keras_h5.is_hdf5(path):
            keras_h5.load_into(self, path)
            return
        try:
            z = np.load(path)
        except ValueError as e:
            raise OSError(f"Unable to open file (neither HDF5 nor npz): {path}: {e}") from e
        with z:
            self.set_weights({k: z[k] for k in z.files})

    def save_weights(self, path):
        if str(path).endswith((".h5", ".hdf5")):
            from . import keras_h5
            keras_h5.save(self, path)
        else:
            np.savez(path, **self.weights)

    def count_params(self):
        return G.count_params(self.spec)

    # ------------------------------------------------------------------ plans
    def build_plan(self, batch, umma=True, keep_f32=False):
        net = G.Net(self.name, batch, self.weights, umma=umma, keep_f32=keep_f32)
        self._build(net)
        return net.plan

    def _drop_native(self):
        for p in self._native.values():
            p.close()
        self._native = {}

    def native_plan(self, batch, umma=True, device=None, keep_buffers=False):
        """The native plan of this model for ``batch`` tiles on ``device`` (None: the device the C ABI context
        resolves from LOCAL_RANK / BD_DEVICE).  One arena per (batch, device): callers that see ragged batch
        sizes go through ``plan_batch_for`` so that the cache stays bounded."""
        from .runtime import NativePlan, resolve_device
        device = resolve_device(device)
        key = (batch, umma, device, bool(keep_buffers))
        if key in self._native:
            self._native[key] = self._native.pop(key)  # most recently used last
            return self._native[key]
        # bound the arenas a long-lived process holds: at most MAX_LARGE_PLANS plans of 16 or more tiles per device
        # (0.45 GB of activations per tile and model: 14 GB for one batch-32 plan), least recently used first out;
        # the small ones (1, 2, 4, 8 tiles) stay
        large = [k for k in self._native if k[0] >= 16 and k[2] == device]
        while batch >= 16 and len(large) >= MAX_LARGE_PLANS:
            self._native.pop(large.pop(0)).close()
        # keep_buffers: every buffer in a range of its own so that intermediates can be read back (tests, tools);
        # the product path shares ranges between buffers with disjoint lifetimes (5-10x smaller arenas)
        self._native[key] = NativePlan(self.build_plan(batch, umma=umma), device, reuse=not keep_buffers)
        return self._native[key]

    @staticmethod
    def plan_batch_for(n):
        """Plan batch that serves ``n`` tiles: the next power of two up to MAX_PLAN_BATCH, so that at most
        five plans per model and device ever exist (a tile's result does not depend on its batch neighbours)."""
        b = 1
        while b < n and b < MAX_PLAN_BATCH:
            b *= 2
        return b

    # ------------------------------------------------------------------ inference
    def predict(self, x, batch_size=None, verbose=0):
        """x: (N,512,512,3) float in [-1,1] (predict.py:93,108).  Returns (N,512,512,2) float32 softmax
        probabilities, like ``tf.keras.Model.predict``."""
        x = np.asarray(x)
        if x.ndim != 4 or x.shape[1:] != (512, 512, 3):
            raise ValueError(f"expected input of shape (N,512,512,3), got {x.shape}")
        x = np.ascontiguousarray(x, dtype=np.float32)  # Keras casts float64 inputs to float32
        out = np.empty((x.shape[0], 512, 512, 2), np.float32)
        i = 0
        while i < x.shape[0]:
            n = min(MAX_PLAN_BATCH, x.shape[0] - i)
            b = self.plan_batch_for(n)
            xb = x[i:i + n]
            if b != n:  # ragged tail: pad with zero tiles, drop their outputs
                xb = np.concatenate([xb, np.zeros((b - n,) + x.shape[1:], np.float32)], axis=0)
            out[i:i + n] = self.native_plan(b).run_host(xb)[:n]
            i += n
        return out

    __call__ = predict
