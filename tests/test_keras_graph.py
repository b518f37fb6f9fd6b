"""The network graphs pinned to the reference's own code.  tools/keras_trace.py executed predict_model/*.py under a
recording stand-in for tensorflow (build container) and committed what the code did -- every layer in construction
order with its arguments and weight shapes, every call with its inputs -- as tests/golden/keras_graph_<model>.json.
Here (CPU, no reference tree needed):
  * parameter totals of the recorded graphs == the package's specs (and SURVEY App. A);
  * the package's builders create their weighted layers in the reference's construction order, class by class, with the
    same shapes -- the order a Keras checkpoint is matched by;
  * oracle/nets.py (hand restatement) == the recorded graph executed layer by layer (oracle/keras_graph.py) on the same
    weights: the wiring of the oracle is the reference's;
  * a Keras-format .h5 written from those weights loads back through Model.load_weights bit for bit (hdf5_min.py)."""
import math
import os

import numpy as np
import pytest
import torch

from building_detection_b200 import graph as G, hdf5_min, keras_h5
from building_detection_b200.predict_model import CTORS, MODEL_NAMES
from oracle import keras_graph, nets


def keras_weights(model, graph, w):
    """{'<keras layer>/<weight>': array} from a weight dict keyed by the package's names, matched by creation order"""
    ours, theirs, out = {}, {}, {}
    for cls, keys in model.keras_layers:
        ours.setdefault(cls, []).append(keys)
    for L in graph["layers"]:
        if L["weights"]:
            theirs.setdefault(L["class"], []).append(L)
    assert {k: len(v) for k, v in ours.items()} == {k: len(v) for k, v in theirs.items()}
    for cls in theirs:
        for keys, L in zip(ours[cls], theirs[cls]):
            assert len(keys) == len(L["weights"]), (cls, keys, L["name"])
            for key, (wn, shape) in zip(keys, L["weights"]):
                assert tuple(model.spec[key][0]) == tuple(shape), (key, model.spec[key][0], L["name"], wn, shape)
                out[f"{L['name']}/{wn}"] = w[key]
    return out


@pytest.mark.parametrize("name", MODEL_NAMES)
def test_recorded_graph_matches_builder_and_oracle(name):
    m = CTORS[name]()
    g = keras_graph.load_graph(name)
    assert all(L["live"] for L in g["layers"])
    assert sum(math.prod(s) for L in g["layers"] for _, s in L["weights"]) == m.count_params()
    w = G.init_weights(m.spec, seed=3, randomize_bn=True)
    kw = keras_weights(m, g, w)
    rng = np.random.default_rng(5)
    x = (rng.integers(0, 256, (1, 512, 512, 3), dtype=np.uint8) / 127.5 - 1).astype(np.float32)
    with torch.no_grad():
        got = nets.FORWARD[name](w, x)
        want = keras_graph.run(g, kw, x)
    assert got.shape == want.shape == (1, 512, 512, 2)
    assert np.abs(got - want).max() < 5e-6, np.abs(got - want).max()


def test_hdf5_subset_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    tree = {"@layer_names": np.array([b"conv2d", b"dense_3"]), "@backend": np.array(b"tensorflow"),
            "conv2d": {"@weight_names": np.array([b"conv2d/kernel:0", b"conv2d/bias:0"]),
                       "conv2d": {"kernel:0": rng.standard_normal((3, 3, 4, 8)).astype(np.float32), "bias:0": np.zeros(8, np.float32)}},
            "dense_3": {"@weight_names": np.array([b"dense_3/kernel:0"]),
                        "dense_3": {"kernel:0": rng.standard_normal((5, 7)).astype(np.float64)}},
            "many": {f"k{i:04d}": np.full((2,), i, np.int32) for i in range(300)}}  # a group larger than the default leaf K
    p = str(tmp_path / "t.h5")
    hdf5_min.write_file(p, tree)
    f = hdf5_min.File(p)
    assert sorted(f.keys()) == ["conv2d", "dense_3", "many"]
    assert [b.decode() for b in f.attrs["layer_names"]] == ["conv2d", "dense_3"] and f.attrs["backend"] == b"tensorflow"
    np.testing.assert_array_equal(f["conv2d/conv2d/kernel:0"].read(), tree["conv2d"]["conv2d"]["kernel:0"])
    np.testing.assert_array_equal(f["dense_3"]["dense_3"]["kernel:0"].read(), tree["dense_3"]["dense_3"]["kernel:0"])
    assert len(f["many"].keys()) == 300 and int(f["many/k0123"].read()[0]) == 123
    with pytest.raises(KeyError):
        f["nope"]
    bad = tmp_path / "bad.h5"
    bad.write_bytes(b"definitely not hdf5" * 10)
    with pytest.raises(OSError):
        hdf5_min.File(str(bad))


@pytest.mark.parametrize("name", ["hrnet", "bam"])
def test_keras_h5_checkpoint_round_trip(name, tmp_path):
    """save in Keras' HDF5 weight layout -> load_weights: every array back in place; shifted layer numbering (a model
    built second in a process), a wrong architecture and a missing file behave like Keras says"""
    m = CTORS[name]()
    w = G.init_weights(m.spec, seed=11, randomize_bn=True)
    m.set_weights(w)
    p = str(tmp_path / f"{name}.h5")
    m.save_weights(p)
    assert keras_h5.is_hdf5(p)
    layers = keras_h5.read_layers(p)
    assert len(layers) == len(m.keras_layers) and layers[0][0] == "conv2d" and layers[0][1][0][0] == "kernel:0"
    m2 = CTORS[name]()
    m2.load_weights(p)
    for k in m.spec:
        np.testing.assert_array_equal(m2.weights[k], w[k], err_msg=k)
    # the same checkpoint with every auto-name index shifted, as when another model was built first (predict.py:17-54)
    f = hdf5_min.File(p)
    shifted = {}
    names = []
    for ln, ws in layers:
        base, idx = keras_h5._split(ln)
        new = f"{base}_{idx + 57}"
        names.append(new)
        shifted[new] = {"@weight_names": np.array([f"{new}/{wn}".encode() for wn, _ in ws]), new: {wn: a for wn, a in ws}}
    shifted["@layer_names"] = np.array([n.encode() for n in names])
    p2 = str(tmp_path / "shifted.h5")
    hdf5_min.write_file(p2, shifted)
    m3 = CTORS[name]()
    m3.load_weights(p2)
    for k in m.spec:
        np.testing.assert_array_equal(m3.weights[k], w[k], err_msg=k)
    other = CTORS["scse"]()
    with pytest.raises(ValueError):
        other.load_weights(p)
    with pytest.raises(OSError):
        m2.load_weights(str(tmp_path / "missing.h5"))
    junk = tmp_path / "junk.h5"
    junk.write_bytes(b"\x00" * 64)
    with pytest.raises(OSError):
        m2.load_weights(str(junk))


def test_load_model_with_keras_checkpoints(tmp_path, capsys):
    """predict.load_model (predict.py:17-54): five constructors, `load_weights` per model, a missing / unreadable file
    is reported and the model keeps its random initialisation (the reference catches OSError only)."""
    from building_detection_b200 import predict as P
    src = CTORS["hrnet"]()
    w = G.init_weights(src.spec, seed=21, randomize_bn=True)
    src.set_weights(w)
    good = str(tmp_path / "hrnet.h5")
    src.save_weights(good)
    junk = tmp_path / "bad.h5"
    junk.write_bytes(b"not a checkpoint")
    models = P.load_model({"hrnet": good, "res34": str(tmp_path / "missing.h5"), "scse": str(junk)})
    out = capsys.readouterr().out
    assert len(models) == 5 and P.hr_model is models[1] and P.res_model is models[0]
    for k in src.spec:
        np.testing.assert_array_equal(P.hr_model.weights[k], w[k])
    assert "load weights hrnet 2/5" in out and out.count("error while loading weights") == 2
    fresh = CTORS["res34"]()
    for k in fresh.spec:  # the failed loads left the seeded random initialisation in place
        np.testing.assert_array_equal(P.res_model.weights[k], fresh.weights[k])
    P.res_model = P.hr_model = P.v3_model = P.unet_model = P.bam_model = None
    P._runner = None
