"""CPU tests of the oracle and of the host-side lowering (no GPU, no native code).

* known answers the reference holds: the res34 encoder parameter count (predict_model/res34.py:164)
  and the per-model totals of SURVEY.md Appendix A;
* oracle/nets.py (fp32 restatement of the Keras graphs) against oracle/plan_interp.py running the
  product's *plan* in fp32: proves the fusions (BN folding, concat elision, sub-pixel transposed convs,
  TF padding, pooled-sum identities) are exact re-writes.
"""
import numpy as np
import pytest
import torch

from building_detection_b200 import graph as G
from building_detection_b200.predict_model import CTORS, MODEL_NAMES
from oracle import nets, plan_interp

TOTAL_PARAMS = {"res34": 38_545_314, "hrnet": 9_607_810, "v3plus": 64_615_674, "scse": 34_558_914, "bam": 62_969_170}
GFLOP = {"res34": 499.06, "hrnet": 187.48, "v3plus": 202.12, "scse": 406.91, "bam": 151.7}


def test_res34_encoder_param_count_kat():
    """'# Trainable params: 22,910,272' -- predict_model/res34.py:164 (encoder = res34() only; BN moving
    statistics are not trainable)."""
    m = CTORS["res34"]()
    enc = ("conv1_", "conv2_", "conv3_", "conv4_", "conv5_", "pool")
    n = G.count_params(m.spec, lambda k: k.startswith(enc) and not k.endswith(("/mean", "/var")))
    assert n == 22_910_272


@pytest.mark.parametrize("name", MODEL_NAMES)
def test_param_and_flop_totals(name):
    m = CTORS[name]()
    assert m.count_params() == TOTAL_PARAMS[name]
    assert abs(m.flops_per_tile / 1e9 - GFLOP[name]) < 0.1


@pytest.mark.parametrize("name", MODEL_NAMES)
def test_plan_lowering_matches_oracle_fp32(name):
    rng = np.random.default_rng(0)
    x = (rng.integers(0, 256, (1, 512, 512, 3), dtype=np.uint8) / 127.5 - 1).astype(np.float32)
    m = CTORS[name]()
    w = nets.calibrated_weights(name, m.spec, 1, x)
    m.set_weights(w)
    with torch.no_grad():
        ref = nets.FORWARD[name](w, x)
    got = plan_interp.run_plan(m.build_plan(1, keep_f32=True), x, emulate_h16=False)
    assert ref.shape == got.shape == (1, 512, 512, 2)
    assert np.abs(ref - got).max() < 2e-4, np.abs(ref - got).max()
    np.testing.assert_allclose(ref.sum(-1), 1.0, atol=1e-5)


def test_keras_surface():
    """Names and error behaviour predict.py relies on (predict.py:5-9, 19-52; res34.py:167)."""
    from building_detection_b200.predict_model.res34 import ResNetFamily
    with pytest.raises(ValueError):
        ResNetFamily().run_model("res50")
    m = CTORS["hrnet"]()
    with pytest.raises(OSError):
        m.load_weights("/nonexistent/hrnet.h5")
    with pytest.raises(ValueError):
        m.predict(np.zeros((1, 256, 256, 3), np.float32))


def _tiny(builder, batch=1, seed=0):
    n0 = G.Net("case", batch, None)
    builder(n0)
    w = G.init_weights(n0.spec, seed=seed, randomize_bn=True)
    n1 = G.Net("case", batch, w, keep_f32=True)
    out = builder(n1)
    return n1.plan, out


def test_conv_up2_is_conv_of_the_upsampled_map():
    """graph.conv_up2 (four sub-pixel 2x2 convolutions, hrnet.py:198-199 / v3plus.py:341-342) against the literal
    lowering: nearest up-sampling followed by the 3x3 convolution with the same weights, in fp32."""
    def fused(g):
        x = g.new(16, 16, 24)
        return x, g.conv_up2(x, "c", 8, bn=True, act="relu")

    def literal(g):
        x = g.new(16, 16, 24)
        return x, g.conv(g.upsample(x, 2), "c", 8, k=3, bn=True, act="relu")

    rng = np.random.default_rng(1)
    xin = rng.standard_normal((1, 16, 16, 24)).astype(np.float32)
    outs = []
    for b in (fused, literal):
        plan, (x, y) = _tiny(b)
        it = plan_interp.Interp(plan, emulate_h16=False)
        it.set(x.buf.id, xin)
        with torch.no_grad():
            it.run(None)
        outs.append(it.get(y.buf.id))
    assert outs[0].shape == outs[1].shape == (1, 32, 32, 8)
    np.testing.assert_allclose(outs[0], outs[1], atol=2e-5, rtol=1e-5)
    # 4 taps per phase instead of 9, algorithmic FLOPs of the reference layer
    plan, _ = _tiny(fused)
    assert [len(op["taps"]) for op in plan.ops] == [4, 4, 4, 4]
    assert sum(op["flops"] for op in plan.ops) == 2 * 32 * 32 * 8 * 24 * 9


def test_split_weights_reproduce_fp32_weights():
    """graph.Net.split_weights: every weight as two fp16 taps at the same offset whose sum is the fp32 weight to
    ~2^-22 relative; in fp32 the lowering is unchanged."""
    def b(split):
        def f(g):
            g.split_weights = split
            x = g.new(16, 16, 32)
            return x, g.conv(x, "c", 16, k=3, bn=True, act="relu")
        return f
    plan_s, _ = _tiny(b(True))
    plan_n, _ = _tiny(b(False))
    ops, opn = plan_s.ops[0], plan_n.ops[0]
    assert len(ops["taps"]) == 18 and ops["taps"][:9] == ops["taps"][9:] == opn["taps"]
    w16 = G.h16_to_f32(ops["w"])
    rec = w16[:9] + w16[9:]
    ref = opn["w32"]
    assert np.abs(rec - ref).max() <= 2.0 ** -21 * np.abs(ref).max()
    assert ops["flops"] == opn["flops"]


def test_separable_convs_are_marked_for_fusion_only_with_one_n_tile():
    def b(cout):
        def f(g):
            x = g.new(32, 32, 128)
            return x, g.sepconv(x, "s", cout, act="relu")
        return f
    small, _ = _tiny(b(256))
    big, _ = _tiny(b(728))
    assert small.ops[0].get("fuse") and small.ops[1].get("fused_dw") == 0
    assert not big.ops[0].get("fuse") and big.ops[1].get("fused_dw") is None


def test_hrnet_parity_recipe_is_well_conditioned_and_the_undamped_one_is_not():
    """Why the HRNet parity weights damp the closing BN of every residual block (oracle.nets.parity_weights): flip
    1e-4 of the stored fp16 activations by ONE ulp in every layer of the fp16-faithful interpreter and look at the
    output.  Undamped (gamma ~ U(0.5,1.5) everywhere) the probabilities move by more than 1e-2 -- as much as the
    whole fp16-vs-fp32 gap, so a 2e-2 max-abs bar would measure summation order -- damped they stay within 5e-3 and
    fp16 meets the north star's 2e-2 against the fp32 oracle with a 4x margin (tools/hrnet_chaos_study.py)."""
    from oracle.plan_interp import _q

    class Perturb(plan_interp.Interp):
        def __init__(self, plan, frac):
            super().__init__(plan, True)
            self.frac, self.rng = frac, np.random.default_rng(0)

        def _store(self, ref, val):
            bid, c0, c = ref
            if self.p.bufs[bid].dtype == "f16":
                h = _q(val).to(torch.float16).contiguous().numpy().view(np.int16).copy()
                h[self.rng.random(h.shape) < self.frac] += 1
                val = torch.from_numpy(h.view(np.float16).astype(np.float32))
            self.b[bid][..., c0:c0 + c] = val

    rng = np.random.default_rng(3)
    x = (rng.integers(0, 256, (1, 512, 512, 3), dtype=np.uint8) / 127.5 - 1).astype(np.float32)
    m = CTORS["hrnet"]()
    spread = {}
    for damped in (False, True):
        w = nets.parity_weights("hrnet", m.spec, hrnet_damped=damped)
        m.set_weights(w)
        plan = m.build_plan(1)
        with torch.no_grad():
            base = plan_interp.Interp(plan, True).run(x)
            spread[damped] = float(np.abs(Perturb(plan, 1e-4).run(x) - base).max())
            if damped:
                ref = nets.FORWARD["hrnet"](w, x)
                assert np.abs(base - ref).max() < 1e-2  # fp16 vs fp32 on the parity recipe: 4.8e-3
                assert ref[..., 1].std() > 0.1 and 0.2 < (ref[..., 1] > 0.5).mean() < 0.8
    assert spread[False] > 1e-2 and spread[True] < 5e-3, spread
