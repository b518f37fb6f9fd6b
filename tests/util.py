"""Shared helpers for the parity tests: single-op plans run natively and through the CPU interpreter."""
import numpy as np
import torch

from building_detection_b200 import graph as G
from oracle import plan_interp


def build_two_pass(builder, batch, seed=0, umma=True, randomize_bn=True):
    """Run ``builder(net)`` once for the weight spec, initialise, run again with weights."""
    n0 = G.Net("case", batch, None, umma=umma)
    builder(n0)
    w = G.init_weights(n0.spec, seed=seed, randomize_bn=randomize_bn)
    n1 = G.Net("case", batch, w, umma=umma, keep_f32=True)
    out = builder(n1)
    return n1.plan, out, w


def rand_map(rng, plan, buf, scale=1.0):
    b = plan.bufs[buf]
    if b.kind == "vec":
        return (rng.standard_normal((plan.batch, b.C)) * scale).astype(np.float32)
    return (rng.standard_normal((plan.batch, b.H, b.W, b.C)) * scale).astype(np.float32)


def run_interp(plan, inputs, emulate_h16=True):
    it = plan_interp.Interp(plan, emulate_h16)
    for k, v in inputs.items():
        it.set(k, v)
    with torch.no_grad():
        it.run(None)
    return it


def run_native(plan, inputs):
    from building_detection_b200.runtime import NativePlan
    npn = NativePlan(plan)
    for k, v in inputs.items():
        npn.write_buffer(k, v)
    npn.run_device(0, 0, 0)
    return npn


def h16_ulp(x):
    """Size of one fp16 ulp at |x| (11 significant bits; subnormal spacing 2^-24 below 2^-14)."""
    x = np.maximum(np.abs(x), 2.0 ** -14)
    return 2.0 ** (np.floor(np.log2(x)) - 10)
