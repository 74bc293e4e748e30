"""GPU parity: convolution engines, encoder, decoder and style_transfer vs the oracle and the
reference's golden vectors.

Tolerances (BASELINE.json north_star): fp32 mode 1e-4 max-abs; tensor-core path 1e-2 max-abs on
images in [0,1].  The tcgen05 kernels run with f16 or bf16 operands (fp32 accumulation in TMEM):
  * "fp16" (default) meets the 1e-2 bar (measured ~2e-3);
  * "bf16" does NOT on the synthetic He-init weights: its 8-bit significand gives ~1.3 % rms relative
    error after 19 layers (max-abs 1.4e-2 on these cases, reproduced by a CPU emulation that rounds
    activations/weights to bf16, see DESIGN.md "Numerics").  It is checked against TOL_BF16 = 2.5e-2
    and reported as a deviation, not hidden.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import ccst_b200
from ccst_b200 import synth
from ccst_b200.overall import OverallStyleAccumulator
from oracle import ccst_oracle as O

pytestmark = pytest.mark.gpu
TOL_FP32 = 1e-4
TOL_TC = 1e-2      # north_star bar for the tensor-core path, met by f16 operands
TOL_BF16 = 2.5e-2  # documented deviation of bf16 operands (see module docstring)
T = torch.from_numpy
DEV = "cuda:0"


@pytest.fixture(scope="module")
def engine(models):
    vgg, dec = models
    return ccst_b200.engine_for(vgg, dec, torch.device(DEV))


def ref_conv(x_nhwc, w, b, relu, mode):
    x = x_nhwc.permute(0, 3, 1, 2).double()
    if mode == 4:  # Upsample -> ReflectionPad2d -> Conv2d (net.py:10-11)
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    y = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), w.double(), b.double())
    if relu:
        y = F.relu(y)
    if mode == 1:
        y = F.interpolate(y, scale_factor=2, mode="nearest")
    if mode == 2:
        y = F.max_pool2d(y, 2, 2, 0, ceil_mode=True)
    if mode == 3:
        return y
    return y.permute(0, 2, 3, 1).contiguous()


CONV_CASES = [
    # N, H, W, Cin, Cout, relu, mode
    (1, 8, 16, 64, 64, True, 0),      # exactly one tile
    (2, 24, 40, 64, 128, True, 0),    # several tiles, partial in x
    (1, 13, 19, 128, 256, True, 0),   # ragged both ways
    (1, 16, 16, 256, 512, True, 0),   # two N tiles, K = 2304
    (1, 12, 12, 512, 256, True, 1),   # Camelyon-sized map, fused nearest x2
    (2, 10, 14, 128, 64, False, 1),
    (1, 16, 32, 64, 64, True, 2),     # fused ceil-mode max-pool
    (1, 15, 21, 128, 128, True, 2),   # odd sizes: partial pooling windows
    (1, 9, 11, 256, 256, True, 2),
    (2, 16, 24, 64, 3, False, 3),     # last decoder conv: 3 channels, NCHW fp32 store
    (1, 5, 7, 64, 3, False, 3),
    # mode 4: nearest x2 BEFORE the conv, fused as four 2x2 phase convolutions (tcgen05 path only)
    (1, 8, 16, 64, 64, True, 4),      # dec8 shape class: resident phase weights, one tile
    (2, 13, 19, 64, 64, True, 4),     # ragged, several tiles per phase
    (1, 2, 2, 64, 64, False, 4),      # smallest map: every pixel is a corner
    (2, 10, 14, 128, 128, True, 4),   # dec6 class (CTA pairs, odd tile count)
    (1, 12, 12, 256, 256, True, 4),   # dec2 class @96^2
    (1, 9, 20, 128, 64, True, 4),
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("precision", ["fp32", "bf16", "fp16"])
def test_single_conv_engines(engine, case, precision):
    n, h, w, cin, cout, relu, mode = case
    if mode == 4 and precision == "fp32":
        pytest.skip("the upsample-fused conv exists on the tcgen05 path only")
    g = torch.Generator().manual_seed(h * 131 + w)
    x = torch.randn((n, h, w, cin), generator=g)
    wt = torch.randn((cout, cin, 3, 3), generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = torch.randn((cout,), generator=g) * 0.1
    if precision != "fp32":
        # operands pre-rounded to the operand type: the tcgen05 path then differs from the fp64
        # reference only by fp32 accumulation order and the rounding of the stored activation
        q = torch.bfloat16 if precision == "bf16" else torch.float16
        x = x.to(q).float()
        wt = wt.to(q).float()
    ref = ref_conv(x, wt, b, relu, mode)
    out = engine.debug_conv3x3(x.to(DEV), wt, b, relu=relu, mode=mode, precision=precision).cpu().double()
    assert out.shape == ref.shape
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item()
    if precision == "fp32" or mode == 3:
        assert err < 2e-5 * max(1.0, scale), (err, scale)
    elif mode == 4:  # + one more rounding of the pre-summed phase weights
        assert err < (2 ** -6 if precision == "bf16" else 2 ** -9) * max(1.0, scale), (err, scale)
    else:  # one rounding of the stored output
        assert err < (2 ** -8 if precision == "bf16" else 2 ** -11) * max(1.0, scale), (err, scale)


@pytest.mark.parametrize("tag", ["sq40", "odd37x45", "r96"])
def test_encoder_decoder_golden_fp32(engine, golden, tag):
    g = golden["net"]
    feat = engine.encode(T(g[tag + "/x"]).to(DEV), "fp32")
    np.testing.assert_allclose(feat.cpu().numpy(), g[tag + "/feat"], rtol=0, atol=1e-4)
    img = engine.decode(T(g[tag + "/feat"]).to(DEV), "fp32")
    np.testing.assert_allclose(img.cpu().numpy(), g[tag + "/dec_of_feat"], rtol=0, atol=1e-4)


@pytest.mark.parametrize("tag", ["sq40", "odd37x45", "r96"])
@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_encoder_decoder_golden_tensor_core(engine, golden, tag, precision):
    """tcgen05 path per stage (catches halo / border errors that a loose image tolerance hides)."""
    g = golden["net"]
    rel = 4e-3 if precision == "fp16" else 3e-2
    feat = engine.encode(T(g[tag + "/x"]).to(DEV), precision).cpu().numpy()
    assert np.abs(feat - g[tag + "/feat"]).max() < rel * np.abs(g[tag + "/feat"]).max()
    img = engine.decode(T(g[tag + "/feat"]).to(DEV), precision).cpu().numpy()
    assert np.abs(img - g[tag + "/dec_of_feat"]).max() < rel * np.abs(g[tag + "/dec_of_feat"]).max()


@pytest.mark.parametrize("tag", ["sq40", "odd37x45", "r96"])
@pytest.mark.parametrize("alpha", [1.0, 0.6])
def test_style_transfer_golden(models, golden, tag, alpha):
    g = golden["net"]
    vgg, dec = models
    x = T(g[tag + "/x"]).to(DEV)
    stat = [T(g[tag + "/style_mean"]).to(DEV), T(g[tag + "/style_std"]).to(DEV)]
    ref = g[f"{tag}/out_a{alpha}"]
    out32 = ccst_b200.style_transfer(vgg, dec, x, stat, alpha, precision="fp32")
    assert out32.shape == ref.shape and out32.dtype == torch.float32 and out32.is_cuda
    assert np.abs(out32.cpu().numpy() - ref).max() < TOL_FP32
    out16 = ccst_b200.style_transfer(vgg, dec, x, stat, alpha)  # default: tcgen05 path, f16 operands
    assert np.abs(out16.cpu().numpy() - ref).max() < TOL_TC
    outb = ccst_b200.style_transfer(vgg, dec, x, stat, alpha, precision="bf16")
    assert np.abs(outb.cpu().numpy() - ref).max() < TOL_BF16


def test_style_transfer_interpolation_and_image_style(models, golden):
    g = golden["net"]
    vgg, dec = models
    x = T(g["sq40/x"]).to(DEV)
    stat = [T(g["sq40/style_mean"]).to(DEV), T(g["sq40/style_std"]).to(DEV)]
    out = ccst_b200.style_transfer(vgg, dec, x, stat, 1.0, [0.25, 0.75], precision="fp32")
    assert np.abs(out.cpu().numpy() - g["sq40/out_interp"]).max() < 1e-4
    # upstream form: style given as an image batch (per-sample statistics)
    s = synth.images(2, 48, 40, 3)
    with torch.no_grad():
        ref = O.style_transfer_image_style(vgg, dec, x.cpu(), s, 0.8)
    out = ccst_b200.style_transfer(vgg, dec, x, s.to(DEV), 0.8, precision="fp32")
    assert (out.cpu() - ref).abs().max().item() < TOL_FP32
    out = ccst_b200.style_transfer(vgg, dec, x, s.to(DEV), 0.8, precision="fp16")
    assert (out.cpu() - ref).abs().max().item() < TOL_TC
    out = ccst_b200.style_transfer(vgg, dec, x, s.to(DEV), 0.8, precision="bf16")
    assert (out.cpu() - ref).abs().max().item() < TOL_BF16


def test_style_transfer_reference_default_size_222(models):
    """--image_size default 222 (mean_std_computation_effcientMem.py:51): 222 -> 28 -> 224."""
    vgg, dec = models
    x = synth.images(1, 222, 222, 8)
    with torch.no_grad():
        f = O.encode_relu4_1(vgg, x)
        stat = O.single_style_stats(O.encode_relu4_1(vgg, synth.images(1, 222, 300, 9)))
        ref = O.style_transfer(vgg, dec, x, stat, 1.0)
    assert f.shape[-2:] == (28, 28) and ref.shape[-2:] == (224, 224)
    sd = [t.to(DEV) for t in stat]
    out32 = ccst_b200.style_transfer(vgg, dec, x.to(DEV), sd, 1.0, precision="fp32")
    out16 = ccst_b200.style_transfer(vgg, dec, x.to(DEV), sd, 1.0, precision="fp16")
    outb = ccst_b200.style_transfer(vgg, dec, x.to(DEV), sd, 1.0, precision="bf16")
    assert (out32.cpu() - ref).abs().max().item() < TOL_FP32
    assert (out16.cpu() - ref).abs().max().item() < TOL_TC
    assert (outb.cpu() - ref).abs().max().item() < TOL_BF16


def test_fused_pool_equals_separate_pool(models, monkeypatch):
    vgg, dec = models
    x = synth.images(2, 72, 56, 4).to(DEV)
    for prec in ("bf16", "fp16"):
        monkeypatch.setenv("CCST_FUSE_POOL", "1")
        a = ccst_b200.Engine(vgg, dec, DEV).encode(x, prec)
        monkeypatch.setenv("CCST_FUSE_POOL", "0")
        b = ccst_b200.Engine(vgg, dec, DEV).encode(x, prec)
        assert torch.equal(a, b)


def test_fused_upsample_matches_unfused_decoder(models, monkeypatch):
    """Upsample folded into the next conv (phase-decomposed 2x2 kernels) vs the 4x-replicating store:
    same function, weights pre-summed and rounded once more, so equal within the 16-bit rounding."""
    vgg, dec = models
    feat = synth.features((2, 512, 9, 13), 11).to(DEV)
    for prec, tol in (("fp16", 2e-3), ("bf16", 1.6e-2)):
        monkeypatch.setenv("CCST_FUSE_UP", "1")
        a = ccst_b200.Engine(vgg, dec, DEV).decode(feat, prec)
        monkeypatch.setenv("CCST_FUSE_UP", "0")
        b = ccst_b200.Engine(vgg, dec, DEV).decode(feat, prec)
        assert a.shape == b.shape == (2, 3, 72, 104)
        assert (a - b).abs().max().item() < tol


def test_overall_statistics_loop(models):
    """mean_std_computation_effcientMem.py:117-137 on 3 batches of images, fp32 and bf16 engines."""
    vgg, dec = models
    batches = [synth.images(n, 64, 64, 50 + i) for i, n in enumerate((3, 3, 2))]
    with torch.no_grad():
        feats = [O.encode_relu4_1(vgg, b) for b in batches]
    mean64, std64, count, imgs = O.overall_style_stats(feats, dtype=torch.float64)
    eng = ccst_b200.engine_for(vgg, dec, torch.device(DEV))
    for prec, tol in (("fp32", 1e-4), ("fp16", 5e-3), ("bf16", 3e-2)):
        acc = OverallStyleAccumulator(eng, prec)
        for b in batches:
            acc.add_images(b.to(DEV))
        mean, std = acc.finalize()
        assert acc.state.count == count and acc.img_count == imgs
        assert (mean.cpu().double() - mean64).abs().max().item() < tol * max(1.0, mean64.abs().max().item())
        assert (std.cpu().double() - std64).abs().max().item() < tol * max(1.0, std64.abs().max().item())
    # given identical features, the accumulator meets the 1e-5 bar
    acc = OverallStyleAccumulator(eng)
    for f in feats:
        acc.add_features(f.to(DEV))
    mean, std = acc.finalize()
    rel = lambda a, b: ((a.cpu().double() - b).abs() / (b.abs() + 1e-4)).max().item()
    assert rel(mean, mean64) < 1e-5 and rel(std, std64) < 1e-5


def test_full_size_batch_is_deterministic_and_shards(models):
    """512x512 (BASELINE config 3 shape, batch 4): two runs are bit-identical, a batch split in two
    (the multi-GPU sharding unit) reproduces the unsplit result, and the tensor-core path agrees
    with the fp32 engine at full size."""
    vgg, dec = models
    eng = ccst_b200.engine_for(vgg, dec, torch.device(DEV))
    x = synth.images(4, 512, 512, 21).to(DEV)
    acc = OverallStyleAccumulator(eng, "fp32").add_images(synth.images(1, 512, 512, 22).to(DEV))
    stat = list(acc.finalize())  # overall style statistics of a one-image "client"
    a = ccst_b200.style_transfer(vgg, dec, x, stat, 1.0)
    b = ccst_b200.style_transfer(vgg, dec, x, stat, 1.0)
    assert torch.equal(a, b) and torch.isfinite(a).all()
    c = torch.cat([ccst_b200.style_transfer(vgg, dec, x[:2], stat, 1.0),
                   ccst_b200.style_transfer(vgg, dec, x[2:], stat, 1.0)])
    assert torch.equal(a, c)
    d = ccst_b200.style_transfer(vgg, dec, x[:1], stat, 1.0, precision="fp32")
    assert (a[:1] - d).abs().max().item() < TOL_TC
    e = ccst_b200.style_transfer(vgg, dec, x[:1], stat, 1.0, precision="bf16")
    assert (e - d).abs().max().item() < 2 * TOL_BF16  # 786k pixels: heavier tail than the 96^2 cases


def test_pipelines_are_bit_reproducible_over_many_runs(models):
    """Race detector: the warp-specialised kernels hand shared-memory buffers between TMA, the tensor
    core and ordinary loads/stores; a missing proxy fence once corrupted a 32-pixel quarter tile about
    once per 10^5 tiles.  40 repetitions of a 4 x 512^2 batch (8192 conv1_1 tiles each) must agree
    bit for bit, encoder and decoder separately."""
    vgg, dec = models
    eng = ccst_b200.engine_for(vgg, dec, torch.device(DEV))
    x = synth.images(4, 512, 512, 21).to(DEV)
    for precision in ("fp16", "bf16"):
        feats = [eng.encode(x, precision).clone() for _ in range(40)]
        assert all(torch.equal(f, feats[0]) for f in feats[1:]), precision
        imgs = [eng.decode(feats[0], precision).clone() for _ in range(20)]
        assert all(torch.equal(f, imgs[0]) for f in imgs[1:]), precision


@pytest.mark.parametrize("hw", [(24, 24), (17, 40), (16, 16)])
def test_style_transfer_smallest_feature_maps(models, hw):
    """relu4_1 maps of 3x3 / 3x5 / 2x2: every pixel is on a border, rows and columns alias both halos."""
    vgg, dec = models
    h, w = hw
    x = synth.images(2, h, w, 60 + h)
    with torch.no_grad():
        stat = O.single_style_stats(O.encode_relu4_1(vgg, synth.images(1, 32, 32, 61)))
        ref = O.style_transfer(vgg, dec, x, stat, 0.9)
    sd = [t.to(DEV) for t in stat]
    out32 = ccst_b200.style_transfer(vgg, dec, x.to(DEV), sd, 0.9, precision="fp32")
    out16 = ccst_b200.style_transfer(vgg, dec, x.to(DEV), sd, 0.9, precision="fp16")
    assert out32.shape == ref.shape
    assert (out32.cpu() - ref).abs().max().item() < TOL_FP32
    assert (out16.cpu() - ref).abs().max().item() < TOL_TC
