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
