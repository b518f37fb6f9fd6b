"""End-to-end parity of the five network forwards on the GPU (through the C ABI) against

 (a) oracle/nets.py -- the fp32 CPU restatement of predict_model/*.py -- with the north star's
     tolerance: per-model softmax probabilities within max-abs 2e-2, argmax masks equal on >= 99.9 %
     of the pixels whose oracle probability is not within 2e-2 of the 0.5 decision boundary;
 (b) oracle/plan_interp.py -- the fp16-faithful interpreter of the same plan -- which isolates kernel
     bugs from fp16 rounding.

Weights (oracle.nets.parity_weights, the same dict on both sides): seeded Keras-default initialisers with
randomised BN gamma/beta and BN moving variance calibrated to each layer's input for the four networks that
have BatchNormalization; He-scaled kernels for the SCSE U-Net, which has none (with glorot_uniform its output
collapses to p1 in [0.42, 0.51] and a parity test says nothing).  Uncalibrated random-init networks are
numerically chaotic (DESIGN.md "Numerics") and cannot meet any tolerance in reduced precision; the reference
ships no trained checkpoints.
Inputs: SURVEY.md section 8d -- rng.integers(0,256) RGB tiles, x/127.5-1.
"""
import numpy as np
import pytest
import torch

from building_detection_b200.predict_model import CTORS, MODEL_NAMES
from building_detection_b200 import runtime as R
from oracle import nets, plan_interp

pytestmark = pytest.mark.gpu

PROB_TOL = 2e-2      # north star: per-model probabilities within max-abs 2e-2
MASK_AGREE = 0.999   # north star: masks agree on >= 99.9 % of pixels
SEEDS = {"res34": 0, "v3plus": 1, "scse": 2, "bam": 2, "hrnet": 3}


def tiles(seed, n):
    rng = np.random.default_rng(seed)
    return (rng.integers(0, 256, (n, 512, 512, 3), dtype=np.uint8) / 127.5 - 1).astype(np.float32)


MAX_EXCLUDED = 0.12  # pixels within 2e-2 of the decision boundary that the mask comparison may skip


@pytest.fixture(scope="module")
def prepared(parity_models):
    cache = {}

    def get(name):
        if name not in cache:
            m = parity_models(name)
            x = tiles(SEEDS[name], 2)
            with torch.no_grad():
                ref = nets.FORWARD[name](m.get_weights(), x)
            cache[name] = (m, x, ref)
        return cache[name]
    return get


@pytest.mark.parametrize("name", MODEL_NAMES)
def test_forward_matches_fp32_oracle(gpu, prepared, name):
    m, x, ref = prepared(name)
    got = m.predict(x)  # engine.Model.predict -> NativePlan.run_host -> bd_plan_run_host
    assert got.shape == ref.shape == (2, 512, 512, 2) and got.dtype == np.float32
    err = np.abs(got - ref).max()
    sure = np.abs(ref[..., 1] - 0.5) > PROB_TOL
    agree = (got.argmax(-1) == ref.argmax(-1))
    print(f"{name}: max|dp|={err:.3e} mean|dp|={np.abs(got - ref).mean():.3e} mask agree all={agree.mean():.5f} "
          f"sure={agree[sure].mean():.5f} (excluded {1 - sure.mean():.4f})")
    assert agree[sure].mean() >= MASK_AGREE
    assert 1 - sure.mean() <= MAX_EXCLUDED, f"{1 - sure.mean():.3f} of the pixels sit within 2e-2 of the boundary"
    # the comparison must bite: both classes present and the probabilities spread out
    assert ref[..., 1].std() >= 0.1 and 0.02 < (ref[..., 1] > 0.5).mean() < 0.98, (ref[..., 1].std(), (ref[..., 1] > 0.5).mean())
    assert err <= PROB_TOL, err


@pytest.mark.parametrize("name", MODEL_NAMES)
def test_forward_matches_fp16_interpreter(gpu, prepared, name):
    m, x, _ = prepared(name)
    plan = m.build_plan(1)
    want = plan_interp.run_plan(plan, x[:1], emulate_h16=True)
    got, mask = m.native_plan(1).run_host(x[:1], want_probs=True, want_mask=True)
    # Summation order is the only difference: 5e-3.  HRNet: 1e-2 -- its 40 stacked residual blocks amplify even on the
    # damped parity recipe (flipping 1e-4 of the interpreter's own stored activations by one ulp moves its output by
    # 3.8e-3, 1e-2 of them by more: tests/test_oracle_nets.py, profiles/r2_hrnet_chaos.txt), and the per-layer trace
    # (profiles/r2a_layer_trace_hrnet_*.txt) shows the GPU differing from the interpreter by at most one or two ulp on a
    # fraction of the elements that grows smoothly from 2.5e-5 (stem) to 0.4 (head): no kernel introduces a jump.
    tol = 1e-2 if name == "hrnet" else 5e-3
    assert np.abs(got - want).max() < tol, np.abs(got - want).max()
    np.testing.assert_array_equal(mask, (got[..., 1] > got[..., 0]).astype(np.uint8))


@pytest.mark.parametrize("name", MODEL_NAMES)
def test_batch16_equals_batch1(gpu, prepared, name):
    """BASELINE configs 2-4 and the scene loop run batch 16; a tile's result must not depend on its batch
    neighbours, nor on the slot it occupies (the ragged last batch of a scene leaves stale tiles in the others)."""
    m, x, _ = prepared(name)
    xb = np.concatenate([x, tiles(7, 14)], axis=0)
    full = m.predict(xb)
    np.testing.assert_array_equal(full[5:6], m.predict(xb[5:6]))
    np.testing.assert_array_equal(full[15:16], m.predict(xb[15:16]))
    three = m.predict(xb[13:16])  # 3 tiles -> batch-4 plan, zero-padded
    np.testing.assert_array_equal(full[13:16], three)


def test_hrnet_undamped_recipe_stays_within_its_conditioning(gpu):
    """HRNet with gamma ~ U(0.5,1.5) on every BatchNormalization (the recipe round 1 tested, 2.6e-2): the map is
    ill conditioned -- tests/test_oracle_nets.py shows on the CPU that one-ulp flips of 1e-4 of the stored
    activations move the probabilities by > 1e-2 -- so the kernels are held to the spread the interpreter itself
    shows under such flips (hard limit 4e-2) and to full mask agreement on the sure pixels; the 2e-2 bar is
    asserted on the well-conditioned parity recipe above."""
    m = CTORS["hrnet"]()
    w = nets.parity_weights("hrnet", m.spec, hrnet_damped=False)
    m.set_weights(w)
    x = tiles(SEEDS["hrnet"], 1)
    with torch.no_grad():
        ref = nets.FORWARD["hrnet"](w, x)
    got = m.predict(x)
    err = np.abs(got - ref).max()
    sure = np.abs(ref[..., 1] - 0.5) > PROB_TOL
    agree = got.argmax(-1) == ref.argmax(-1)
    print(f"hrnet undamped: max|dp|={err:.3e} p99.99={np.quantile(np.abs(got - ref), 0.9999):.3e} sure-agree={agree[sure].mean():.5f}")
    assert err < 4e-2 and agree[sure].mean() >= MASK_AGREE
    m._drop_native()


@pytest.mark.parametrize("name", MODEL_NAMES)
def test_arena_reuse_is_invisible(gpu, prepared, name):
    """The product plans share arena ranges between buffers with disjoint lifetimes (bd_plan_set_arena_reuse).  The
    allocator's own invariant -- two buffers whose lifetimes intersect never overlap in memory -- is checked from the
    lifetimes and addresses it reports, and the probabilities of a forward are bit-identical to the plan that keeps
    every buffer apart, also on a second forward over the dirty arena (nothing relies on the arena's zero fill)."""
    m, x, _ = prepared(name)
    xb = np.concatenate([x[:1], tiles(11, 1)], axis=0)
    shared = m.native_plan(2)
    flat = m.native_plan(2, keep_buffers=True)
    assert shared.reuse and not flat.reuse
    assert flat.arena_bytes == flat.arena_bytes_flat == shared.arena_bytes_flat
    ratio = shared.arena_bytes / shared.arena_bytes_flat
    print(f"{name}: arena {shared.arena_bytes / 2**20:.0f} MiB with reuse, {shared.arena_bytes_flat / 2**20:.0f} MiB without ({ratio:.2f})")
    assert ratio < 0.5
    spans = []
    for b in shared.plan.bufs:
        first, last, is_shared = shared.buffer_lifetime(b.id)
        if b.kind != "map" or last < 0:
            assert not is_shared
            continue
        p0 = shared.buffer_ptr(b.id)
        spans.append((first, last, p0, p0 + R.lib().bd_plan_buffer_bytes(shared.h, b.id), b.id))
        if b.id in (shared.plan.input, shared.plan.logits):
            assert not is_shared
    assert len(spans) > 10
    for i, (f0, l0, a0, e0, id0) in enumerate(spans):
        for f1, l1, a1, e1, id1 in spans[i + 1:]:
            if f0 <= l1 and f1 <= l0:
                assert e0 <= a1 or e1 <= a0, f"buffers {id0} and {id1} are alive together and overlap"
    want = flat.run_host(xb)
    np.testing.assert_array_equal(shared.run_host(xb), want)
    np.testing.assert_array_equal(shared.run_host(xb[::-1].copy()), flat.run_host(xb[::-1].copy()))
    np.testing.assert_array_equal(shared.run_host(xb), want)
    with pytest.raises(R.NativeError):
        some = next(b.id for b in shared.plan.bufs if shared.buffer_lifetime(b.id)[2])
        shared.read_buffer(some)
    m._drop_native()
