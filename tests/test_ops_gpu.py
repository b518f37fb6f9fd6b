"""GPU parity of every kernel class, called through the C ABI (NativePlan -> libbd_b200.so), against the
fp16-faithful CPU interpreter of the same plan (oracle/plan_interp.py): same fp16 inputs and weights,
fp32 accumulation on both sides, so a single fused op must agree to a couple of fp16 ulps (summation
order is the only difference)."""
import numpy as np
import pytest

from building_detection_b200 import graph as G
from util import build_two_pass, h16_ulp, rand_map, run_interp, run_native

pytestmark = pytest.mark.gpu


def assert_close(got, ref, ulps=2.0, rel_rms=2e-4):
    err = np.abs(got - ref)
    tol = ulps * h16_ulp(ref) + rel_rms * np.sqrt((ref.astype(np.float64) ** 2).mean())
    bad = err > tol
    assert not bad.any(), f"{int(bad.sum())}/{err.size} elements off; max|d|={err.max():.3e}"


def conv_case(N, H, W, Cin, Cout, k=3, s=1, d=1, res=False, res_after_act=False, act="relu", umma=True, bn=True,
              in_slice=None, out_slice=None, seed=0, expect=None, split=False):
    def builder(g):
        g.split_weights = split  # hi/lo fp16 weight taps (18-tap halo path / 2-tap 1x1)
        x = G.T(g.buf(H, W, in_slice[1]), in_slice[0], Cin) if in_slice else g.new(H, W, Cin)
        Ho, Wo = -(-H // s), -(-W // s)
        r = g.new(Ho, Wo, Cout) if res else None
        out = G.T(g.buf(Ho, Wo, out_slice[1]), out_slice[0], Cout) if out_slice else None
        y = g.conv(x, "c", Cout, k=k, s=s, d=d, bn=bn, act=act, res=r, res_after_act=res_after_act, out=out)
        return x, r, y

    plan, (x, r, y), _ = build_two_pass(builder, N, seed=seed, umma=umma)
    if expect:
        assert plan.ops[0]["path"] == expect
    rng = np.random.default_rng(seed + 1)
    inputs = {x.buf.id: rand_map(rng, plan, x.buf.id)}
    if r is not None:
        inputs[r.buf.id] = rand_map(rng, plan, r.buf.id)
    ref = run_interp(plan, inputs).get(y.buf.id)
    nat = run_native(plan, inputs)
    got = nat.read_buffer(y.buf.id)
    nat.close()
    assert_close(got, ref)  # whole buffer: channels outside the output slice must stay untouched (zero)


UMMA_CASES = {
    "1x1_64_64_plain": dict(N=1, H=32, W=32, Cin=64, Cout=64, k=1, act=None, bn=False),
    "1x1_bn_relu": dict(N=1, H=32, W=32, Cin=64, Cout=64, k=1),
    "3x3_64_64": dict(N=1, H=32, W=32, Cin=64, Cout=64),
    "1x1_two_kchunks": dict(N=1, H=32, W=32, Cin=128, Cout=128, k=1),
    "3x3_256_256": dict(N=1, H=32, W=32, Cin=256, Cout=256),
    "3x3_32_32_k_oob_fill": dict(N=1, H=32, W=32, Cin=32, Cout=32),
    "1x1_64_16": dict(N=1, H=32, W=32, Cin=64, Cout=16, k=1),
    "3x3_batch3": dict(N=3, H=64, W=64, Cin=64, Cout=64),
    "1x1_728_728_tails": dict(N=2, H=32, W=32, Cin=728, Cout=728, k=1),
    "1x1_1536_2048": dict(N=1, H=32, W=32, Cin=1536, Cout=2048, k=1),
    "3x3_d6": dict(N=1, H=32, W=32, Cin=256, Cout=256, d=6),
    "3x3_d12": dict(N=1, H=32, W=32, Cin=256, Cout=256, d=12),
    "3x3_d18_2048_256": dict(N=1, H=32, W=32, Cin=2048, Cout=256, d=18),
    "3x3_s2_tf_same_pad": dict(N=2, H=64, W=64, Cin=64, Cout=128, s=2),
    "1x1_s2": dict(N=2, H=64, W=64, Cin=64, Cout=128, k=1, s=2, act=None, bn=False),
    "3x3_res_act_after_add": dict(N=1, H=32, W=32, Cin=64, Cout=64, res=True),
    "3x3_res_after_act": dict(N=1, H=32, W=32, Cin=64, Cout=64, res=True, res_after_act=True),
    "3x3_res_no_act": dict(N=1, H=32, W=32, Cin=64, Cout=64, res=True, act=None),
    "3x3_channel_slices": dict(N=1, H=32, W=32, Cin=64, Cout=64, in_slice=(64, 192), out_slice=(32, 128)),
    "3x3_16x16_batch2": dict(N=2, H=16, W=16, Cin=64, Cout=64),
    "3x3_8x8_batch4": dict(N=4, H=8, W=8, Cin=64, Cout=64),
    "3x3_ragged_40x24": dict(N=1, H=40, W=24, Cin=64, Cout=64),
    "3x3_384_384": dict(N=1, H=64, W=64, Cin=384, Cout=384),
    "3x3_640_640": dict(N=1, H=64, W=64, Cin=640, Cout=640),
    "3x3_64_64_256x256": dict(N=1, H=256, W=256, Cin=64, Cout=64),
    "3x3_1024_1024": dict(N=1, H=32, W=32, Cin=1024, Cout=1024),
    "3x3_64_64_split_weights_halo18": dict(N=2, H=32, W=32, Cin=64, Cout=64, split=True),
    "3x3_256_256_split_weights": dict(N=1, H=32, W=32, Cin=256, Cout=256, split=True),
    "1x1_64_256_split_weights": dict(N=1, H=32, W=32, Cin=64, Cout=256, k=1, split=True),
    # halo path with streamed weights (conv_umma_kernel<5,1>): three taps per ring stage for N <= 128, one for wider
    # tiles; ragged maps, a partial last channel chunk, residual, channel slices, several N tiles
    "3x3_128_128_stream_ragged": dict(N=2, H=40, W=24, Cin=128, Cout=128),
    "3x3_128_128_stream_res": dict(N=3, H=32, W=32, Cin=128, Cout=128, res=True, res_after_act=True),
    "3x3_96_128_stream_partial_chunk": dict(N=2, H=32, W=16, Cin=96, Cout=128, in_slice=(64, 256)),
    "3x3_256_128_stream_slices": dict(N=1, H=32, W=32, Cin=256, Cout=128, in_slice=(64, 384), out_slice=(128, 256)),
    "3x3_128_320_stream_two_n_tiles": dict(N=2, H=16, W=24, Cin=128, Cout=320),
    "3x3_512_256_stream_batch5": dict(N=5, H=16, W=8, Cin=512, Cout=256, res=True),
}


@pytest.mark.parametrize("case", sorted(UMMA_CASES))
def test_conv_tcgen05(gpu, case):
    conv_case(expect="umma", **UMMA_CASES[case])


DIRECT_CASES = {
    "1x1_16": dict(N=1, H=16, W=16, Cin=16, Cout=16, k=1),
    "3x3_cin8_64": dict(N=1, H=32, W=32, Cin=8, Cout=64),
    "3x3_s2": dict(N=2, H=32, W=32, Cin=16, Cout=24, s=2),
    "3x3_d4_c4": dict(N=1, H=32, W=32, Cin=4, Cout=4, d=4),
    "3x3_d4_c45": dict(N=1, H=32, W=32, Cin=45, Cout=45, d=4),
    "1x1_to_1": dict(N=2, H=32, W=32, Cin=45, Cout=1, k=1, act=None, bn=False),
    "3x3_tiny_map_4x4": dict(N=2, H=4, W=4, Cin=32, Cout=32),
    "3x3_res_after_act": dict(N=1, H=16, W=16, Cin=16, Cout=16, res=True, res_after_act=True),
}


@pytest.mark.parametrize("case", sorted(DIRECT_CASES))
def test_conv_direct(gpu, case):
    conv_case(umma=False, expect="direct", **DIRECT_CASES[case])


SMALL_CASES = {  # <= 16 output channels, <= 600 multiply-adds per pixel: the CUDA-core small-channel kernel
    "1x1_64_8": dict(N=2, H=64, W=64, Cin=64, Cout=8, k=1),
    "1x1_32_16_plain": dict(N=1, H=32, W=32, Cin=32, Cout=16, k=1, act=None, bn=False),
    "3x3_d4_8_8": dict(N=2, H=40, W=24, Cin=8, Cout=8, d=4),
    "1x1_16_8_slices": dict(N=1, H=32, W=32, Cin=16, Cout=8, k=1, in_slice=(16, 48), out_slice=(8, 24)),
}


@pytest.mark.parametrize("case", sorted(SMALL_CASES))
def test_conv_small_channels(gpu, case):
    conv_case(expect="small", **SMALL_CASES[case])


def test_stem_conv(gpu):
    """3x3 stem on the RGB tile, lowered to a K=32 tensor-core 1x1 conv over the im2col'ed input buffer that
    bd_plan_run builds from the float tile (res34.py:50 stride 1, hrnet.py:168 stride 2)."""
    from building_detection_b200.runtime import NativePlan
    from oracle import plan_interp
    import torch
    for s in (1, 2):
        def builder(g):
            x = g.input()
            return x, g.conv(x, "stem", 64, k=3, s=s, bn=True, act="relu")
        plan, (x, y), _ = build_two_pass(builder, 2)
        assert plan.ops[0]["path"] == "umma" and plan.input_stride == s
        rng = np.random.default_rng(3)
        xin = (rng.integers(0, 256, (2, 512, 512, 3)) / 127.5 - 1).astype(np.float32)
        it = plan_interp.Interp(plan, True)
        with torch.no_grad():
            it.run(xin)
        nat = NativePlan(plan)
        xd = torch.from_numpy(xin).cuda()
        nat.run_device(xd.data_ptr(), 0, 0)
        assert_close(nat.read_buffer(x.buf.id), it.get(x.buf.id), ulps=0.0, rel_rms=0.0)  # the im2col'ed input, exact
        assert_close(nat.read_buffer(y.buf.id), it.get(y.buf.id))
        nat.close()


@pytest.mark.parametrize("cout,k", [(2, 3), (2, 1), (1, 1)])
def test_fp32_head_conv(gpu, cout, k):
    """2-channel logits / 1-channel gate maps: one 16-column tensor-core tile, fp32 direct stores (res34.py:87)."""
    def builder(g):
        x = g.new(64, 64, 64)
        return x, g.conv(x, "head", cout, k=k, f32_out=True)
    plan, (x, y), _ = build_two_pass(builder, 2)
    assert plan.ops[0]["path"] == ("small" if k == 1 else "umma")  # 64 x cout multiply-adds per pixel: CUDA cores
    rng = np.random.default_rng(4)
    inputs = {x.buf.id: rand_map(rng, plan, x.buf.id)}
    ref = run_interp(plan, inputs).get(y.buf.id)
    nat = run_native(plan, inputs)
    got = nat.read_buffer(y.buf.id)
    nat.close()
    assert np.abs(got - ref).max() < 2e-4 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("k,umma", [(2, True), (3, True), (2, False), (3, False)])
def test_conv_transpose(gpu, k, umma):
    """Conv2DTranspose k2/k3 stride 2 'same' as four sub-pixel convolutions (res34.py:144, scse.py:71)."""
    C = (128, 64) if umma else (16, 8)

    def builder(g):
        x = g.new(32, 32, C[0])
        return x, g.conv_transpose(x, "t", C[1], k, act="relu")
    plan, (x, y), _ = build_two_pass(builder, 2, umma=umma)
    assert all(op["path"] == ("umma" if umma else "direct") for op in plan.ops)
    rng = np.random.default_rng(5)
    inputs = {x.buf.id: rand_map(rng, plan, x.buf.id)}
    ref = run_interp(plan, inputs).get(y.buf.id)
    nat = run_native(plan, inputs)
    assert_close(nat.read_buffer(y.buf.id), ref)
    nat.close()


def run_case(builder, n=2, seed=0, **tol):
    plan, (ins, outs), _ = build_two_pass(builder, n, seed=seed)
    rng = np.random.default_rng(seed + 7)
    inputs = {}
    for t in ins:
        inputs.setdefault(t.buf.id, rand_map(rng, plan, t.buf.id))
    it = run_interp(plan, inputs)
    nat = run_native(plan, inputs)
    try:
        for t in outs:
            assert_close(nat.read_buffer(t.buf.id), it.get(t.buf.id), **tol)
    finally:
        nat.close()


@pytest.mark.parametrize("C,H,s,relu_in", [(728, 32, 1, True), (64, 64, 2, False), (128, 256, 1, False), (256, 64, 2, True)])
def test_separable_conv(gpu, C, H, s, relu_in):
    """depthwise 3x3 (CUDA cores) + pointwise 1x1 (tcgen05), v3plus.py:187-278; two chained roundings."""
    def b(g):
        x = g.new(H, H, C)
        return [x], [g.sepconv(x, "s", C, s=s, relu_in=relu_in, act="relu")]
    run_case(b, ulps=4.0, rel_rms=1e-3)


@pytest.mark.parametrize("C,Cout,H,N,relu_in,res", [(728, 256, 32, 2, True, True), (64, 128, 64, 1, False, False),
                                                     (128, 128, 32, 3, True, False), (1024, 192, 16, 1, True, False),
                                                     (256, 256, 40, 1, False, False)])
def test_fused_separable_conv_equals_two_kernels(gpu, monkeypatch, C, Cout, H, N, relu_in, res):
    """The one-kernel SeparableConv2D (depthwise warps feed the A tile of the pointwise tcgen05 GEMM,
    bd_conv_desc::dw_w_host) against the two-kernel sequence: same fp16 rounding points, so bit-identical."""
    def b(g):
        x = g.new(H, H, C)
        r = g.new(H, H, Cout) if res else None
        return x, r, g.sepconv(x, "s", Cout, relu_in=relu_in, act="relu", res=r)
    plan, (x, r, y), _ = build_two_pass(b, N)
    assert plan.ops[0].get("fuse") and plan.ops[1].get("fused_dw") == 0
    rng = np.random.default_rng(3)
    inputs = {x.buf.id: rand_map(rng, plan, x.buf.id)}
    if r is not None:
        inputs[r.buf.id] = rand_map(rng, plan, r.buf.id)
    outs = []
    for fuse in ("1", "0"):
        monkeypatch.setenv("BD_FUSE_SEPCONV", fuse)
        nat = run_native(plan, inputs)
        assert (len(nat.native_to_plan) == 1) == (fuse == "1")
        outs.append(nat.read_buffer(y.buf.id))
        nat.close()
    np.testing.assert_array_equal(outs[0], outs[1])
    assert_close(outs[0], run_interp(plan, inputs).get(y.buf.id), ulps=4.0, rel_rms=1e-3)


def test_maxpools(gpu):
    """MaxPool 2x2/s2 (scse.py:54), 2x2/s4 valid (res34.py:153), 3x3/s2 same (v3plus.py:192): exact."""
    def b(g):
        x = g.new(64, 64, 64)
        return [x], [g.maxpool(x, 2, 2), g.maxpool(x, 2, 4), g.maxpool(x, 3, 2, same=True)]
    run_case(b, ulps=0.0, rel_rms=0.0)


def test_addn_upsample_into_slices(gpu):
    """HRNet fuse adds with nearest upsampling, written into concat slices (hrnet.py:99-162)."""
    def b(g):
        a, bb, c = g.new(64, 64, 32), g.new(32, 32, 32), g.new(16, 16, 32)
        cat = g.buf(64, 64, 64)
        y = g.addn([(a, 1), (bb, 2), (c, 4)], out=G.T(cat, 32, 32))
        g.upsample(c, 4, out=G.T(cat, 0, 32))
        return [a, bb, c], [y]
    run_case(b, ulps=1.0)


def test_se_attention(gpu):
    """res34 attention_demo: GAP -> Dense+BN+ReLU -> Dense+BN+sigmoid -> scale (res34.py:90-105)."""
    def b(g):
        x = g.new(32, 32, 64)
        v = g.gap(x)
        v1 = g.dense([v], "fc1", 32, bn="bn1", act="relu")
        v2 = g.dense([v1], "fc2", 64, bn="bn2", act="sigmoid")
        return [x], [g.gate_se(x, v2)]
    run_case(b)


@pytest.mark.parametrize("C,H", [(64, 64), (128, 32), (256, 32), (512, 16)])
def test_scse_gate(gpu, C, H):
    """scSE: x*sigmoid(conv1x1->1(x)) + x*sigmoid(W2 W1 GAP(x)) (scse.py:20-46)."""
    def b(g):
        x = g.new(H, H, C)
        return [x], [g.scse(x, "q")]
    run_case(b)


@pytest.mark.parametrize("C", [64, 128, 720])
def test_bam_block(gpu, C):
    """BAM: x*(1+sigmoid(channel_gate + spatial_gate)), C/16-channel dilated convs on CUDA cores (bam.py:20-71)."""
    from building_detection_b200.predict_model._xception import bam_attention

    def b(g):
        x = g.new(32, 32, C)
        return [x], [bam_attention(g, x, "bam")]
    run_case(b, ulps=4.0, rel_rms=1e-3)


def test_sk_aspp_neck(gpu):
    """SK block || ASPP -> projection -> 2 convs -> scSE (v3plus.py:74-138, 295-316): a chain of ~10
    roundings with K up to 18432, so the bound is a relative one on the chain output."""
    from building_detection_b200.predict_model._xception import neck

    def b(g):
        x = g.new(32, 32, 2048)
        return [x], [neck(g, x)]
    plan, (ins, outs), _ = build_two_pass(b, 1)
    rng = np.random.default_rng(7)
    inputs = {ins[0].buf.id: rand_map(rng, plan, ins[0].buf.id)}
    it = run_interp(plan, inputs)
    nat = run_native(plan, inputs)
    ref, got = it.get(outs[0].buf.id), nat.read_buffer(outs[0].buf.id)
    nat.close()
    rms = np.sqrt((ref ** 2).mean())
    assert np.abs(got - ref).max() < 1e-2 * rms, (np.abs(got - ref).max(), rms)


def test_softmax_head_and_mask(gpu):
    """2-class softmax + argmax mask (ties -> class 0, predict.py:110), with the bam head's x4 nearest upsample."""
    from building_detection_b200.runtime import NativePlan
    for up in (1, 4):
        def builder(g):
            x = g.new(512 // up, 512 // up, 16)
            lg = g.conv(x, "head", 2, k=1, f32_out=True)
            g.softmax_head(lg, up=up)
            return x, lg
        plan, (x, lg), _ = build_two_pass(builder, 2)
        plan.input = -1
        nat = NativePlan(plan)
        rng = np.random.default_rng(11)
        logits = rng.standard_normal((2, 512 // up, 512 // up, 2)).astype(np.float32)
        logits[0, :8, :8, 1] = logits[0, :8, :8, 0]  # exact ties
        import torch
        probs = torch.empty((2, 512, 512, 2), dtype=torch.float32, device="cuda")
        mask = torch.empty((2, 512, 512), dtype=torch.uint8, device="cuda")
        # run the conv first, then overwrite the logits so that the head sees the crafted values
        nat.run_device(0, 0, 0)
        nat.write_buffer(lg.buf.id, logits)
        from building_detection_b200 import runtime as R
        R.check(R.lib().bd_plan_run_head(nat.h, probs.data_ptr(), mask.data_ptr(), None))
        torch.cuda.synchronize()
        lu = logits.repeat(up, axis=1).repeat(up, axis=2)
        e = np.exp(lu - lu.max(-1, keepdims=True))
        np.testing.assert_allclose(probs.cpu().numpy(), e / e.sum(-1, keepdims=True), atol=1e-6)
        np.testing.assert_array_equal(mask.cpu().numpy(), (lu[..., 1] > lu[..., 0]).astype(np.uint8))
        nat.close()
