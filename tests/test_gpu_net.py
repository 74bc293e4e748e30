"""GPU parity: convolution engines, encoder, decoder and style_transfer vs the oracle and the
reference's golden vectors.

Tolerances (BASELINE.json north_star): fp32 mode 1e-4 max-abs; tensor-core path 1e-2 max-abs on
images in [0,1]; style statistics 1e-5 relative.  No assert in this file is looser than those bars.
The tcgen05 kernels run with f16 or bf16 operands (fp32 accumulation in TMEM):
  * "fp16" (default) meets the 1e-2 image bar (measured ~2e-3);
  * "bf16" does NOT on the synthetic He-init weights: its 8-bit significand gives ~1.3 % rms relative
    error after 19 layers (max-abs 1.3e-2 .. 1.6e-2, reproduced by a CPU emulation that rounds
    activations/weights to bf16, see DESIGN.md "Numerics").  Every bf16 image comparison is asserted
    at the SAME 1e-2 bar and marked xfail(strict=True): a documented miss, not a widened pass.
  * statistics taken through a 16-bit encoder (either operand type) cannot meet 1e-5; those cases are
    strict xfails too and the measured error is printed.  The fp32 engine meets both bars.
  * "fp16x3" / "bf16x3" (split hi + lo operands on the same tensor pipe, csrc/conv_x3.cuh) run every entry
    point; both are asserted at the fp32 mode's 1e-4 image bar (measured 2e-6 / 3e-5), fp16x3 also at the
    1e-5 statistics bar.  bf16x3 is the tensor-core mode that meets the bf16 path's 1e-2 bar.
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
TOL_TC = 1e-2      # north_star bar for the tensor-core path (f16 and bf16 operands alike)
TOL_STATS = 1e-5   # north_star bar for style statistics (relative)
T = torch.from_numpy
DEV = "cuda:0"
BF16_MISS = pytest.mark.xfail(strict=True, reason="bf16 operands: 1.3e-2..1.6e-2 max-abs on these weights, above the "
                                                  "1e-2 bar of BASELINE.json (DESIGN.md Numerics); f16 is the default")
STATS16_MISS = pytest.mark.xfail(strict=True, reason="style statistics through a 16-bit encoder are ~1e-3 relative, above "
                                                     "the 1e-5 bar; the statistics drivers default to the fp32 engine")
PREC_IMG = ["fp32", "fp16x3", "fp16", "bf16x3", pytest.param("bf16", marks=BF16_MISS)]
# x3 engines (split operands on the tensor pipe): f16 halves give 22 significand bits, bf16 halves 16 bits at
# the fp32 exponent range; both are held to the fp32 mode's 1e-4 bar (measured 2e-6 / 3e-5)
IMG_TOL = {"fp32": TOL_FP32, "fp16x3": TOL_FP32, "fp16": TOL_TC, "bf16x3": TOL_FP32, "bf16": TOL_TC}


def report(name, **kv):
    print("[measured] " + name + ": " + ", ".join(f"{k}={v:.3e}" if isinstance(v, float) else f"{k}={v}" for k, v in kv.items()))


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
    (2, 24, 40, 128, 64, True, 0),    # dec7 shape class: s-merged kernel with six resident weight tiles (Cin = 128)
    (3, 9, 33, 128, 64, True, 0),     # odd tile count, ragged in x and y
    (1, 16, 32, 64, 64, True, 2),     # fused ceil-mode max-pool
    (1, 15, 21, 128, 128, True, 2),   # odd sizes: partial pooling windows
    (1, 9, 11, 256, 256, True, 2),
    (2, 16, 24, 64, 3, False, 3),     # last decoder conv: 3 channels, NCHW fp32 store
    (1, 5, 7, 64, 3, False, 3),
    (1, 12, 20, 64, 3, False, 3),     # 3 tiles: the CTA-pair kernel's last pair has a dummy second tile
    (3, 20, 61, 64, 3, True, 3),      # 45 tiles (odd), three tile columns, ReLU variant
    # mode 4: nearest x2 BEFORE the conv, fused as four 2x2 phase convolutions (tcgen05 path only)
    (1, 8, 16, 64, 64, True, 4),      # dec8 shape class: resident phase weights, one tile
    (1, 12, 16, 64, 64, True, 4),     # 3 tiles: odd count for the CTA-pair form (dummy second tile)
    (1, 20, 70, 64, 64, False, 4),    # 15 tiles, three tile columns, no ReLU
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
def test_encoder_golden_f16x3(engine, golden, tag):
    """Split-operand tensor-core encoder (f16 hi + lo parts, three MMAs per product term): the fp32 mode's
    1e-4 bar on relu4_1 against the reference's golden features, odd sizes / ragged tiles included."""
    g = golden["net"]
    feat = engine.encode(T(g[tag + "/x"]).to(DEV), "fp16x3").cpu().numpy()
    err = float(np.abs(feat - g[tag + "/feat"]).max())
    report(f"encoder f16x3 {tag}", max_abs=err, feat_max=float(np.abs(g[tag + "/feat"]).max()))
    np.testing.assert_allclose(feat, g[tag + "/feat"], rtol=0, atol=1e-4)


@pytest.mark.parametrize("precision", ["fp16x3", "bf16x3"])
@pytest.mark.parametrize("tag", ["sq40", "odd37x45", "r96"])
def test_decoder_golden_x3(engine, golden, tag, precision):
    """The decoder on the split-operand engines (nearest x2 stored by the producing conv's epilogue, the last
    conv's 3 channels in a zero-padded 64-channel tile): the fp32 mode's 1e-4 bar for both operand types."""
    g = golden["net"]
    img = engine.decode(T(g[tag + "/feat"]).to(DEV), precision).cpu().numpy()
    err = float(np.abs(img - g[tag + "/dec_of_feat"]).max())
    report(f"decoder {precision} {tag}", max_abs=err, img_max=float(np.abs(g[tag + "/dec_of_feat"]).max()))
    assert err < IMG_TOL[precision]


def test_debug_conv_rejects_x3(engine):
    with pytest.raises(RuntimeError, match="x3"):
        engine.debug_conv3x3(torch.zeros((1, 8, 16, 64), device=DEV), torch.zeros((64, 64, 3, 3)), torch.zeros(64),
                             precision="fp16x3")


@pytest.mark.parametrize("tag", ["sq40", "odd37x45", "r96"])
def test_encoder_decoder_golden_tensor_core(engine, golden, tag):
    """tcgen05 path per stage (catches halo / border errors that a loose image tolerance hides): f16
    operands within 4e-3 of each stage's range; the bf16 figures are printed, not asserted (the image-
    level bf16 comparisons below carry the strict xfail)."""
    g = golden["net"]
    for precision in ("fp16", "bf16"):
        feat = engine.encode(T(g[tag + "/x"]).to(DEV), precision).cpu().numpy()
        e_enc = np.abs(feat - g[tag + "/feat"]).max() / np.abs(g[tag + "/feat"]).max()
        img = engine.decode(T(g[tag + "/feat"]).to(DEV), precision).cpu().numpy()
        e_dec = np.abs(img - g[tag + "/dec_of_feat"]).max() / np.abs(g[tag + "/dec_of_feat"]).max()
        report(f"per-stage {tag} {precision}", encoder_rel=float(e_enc), decoder_rel=float(e_dec))
        if precision == "fp16":
            assert e_enc < 4e-3 and e_dec < 4e-3


@pytest.mark.parametrize("tag", ["sq40", "odd37x45", "r96"])
@pytest.mark.parametrize("alpha", [1.0, 0.6])
@pytest.mark.parametrize("precision", PREC_IMG)
def test_style_transfer_golden(models, golden, tag, alpha, precision):
    g = golden["net"]
    vgg, dec = models
    x = T(g[tag + "/x"]).to(DEV)
    stat = [T(g[tag + "/style_mean"]).to(DEV), T(g[tag + "/style_std"]).to(DEV)]
    ref = g[f"{tag}/out_a{alpha}"]
    kw = {} if precision == "fp16" else {"precision": precision}  # fp16 = the default tcgen05 path
    out = ccst_b200.style_transfer(vgg, dec, x, stat, alpha, **kw)
    assert out.shape == ref.shape and out.dtype == torch.float32 and out.is_cuda
    err = float(np.abs(out.cpu().numpy() - ref).max())
    report(f"style_transfer {tag} a={alpha} {precision}", max_abs=err)
    assert err < IMG_TOL[precision]


def test_style_transfer_interpolation_branch(models, golden):
    g = golden["net"]
    vgg, dec = models
    x = T(g["sq40/x"]).to(DEV)
    stat = [T(g["sq40/style_mean"]).to(DEV), T(g["sq40/style_std"]).to(DEV)]
    out = ccst_b200.style_transfer(vgg, dec, x, stat, 1.0, [0.25, 0.75], precision="fp32")
    assert np.abs(out.cpu().numpy() - g["sq40/out_interp"]).max() < 1e-4


@pytest.mark.parametrize("precision", PREC_IMG)
def test_style_transfer_image_style(models, golden, precision):
    """upstream form: style given as an image batch (per-sample statistics)"""
    vgg, dec = models
    x = T(golden["net"]["sq40/x"]).to(DEV)
    s = synth.images(2, 48, 40, 3)
    with torch.no_grad():
        ref = O.style_transfer_image_style(vgg, dec, x.cpu(), s, 0.8)
    out = ccst_b200.style_transfer(vgg, dec, x, s.to(DEV), 0.8, precision=precision)
    err = (out.cpu() - ref).abs().max().item()
    report(f"image-style {precision}", max_abs=err)
    assert err < IMG_TOL[precision]


@pytest.mark.parametrize("precision", PREC_IMG)
def test_style_transfer_reference_default_size_222(models, precision):
    """--image_size default 222 (mean_std_computation_effcientMem.py:51): 222 -> 28 -> 224."""
    vgg, dec = models
    x = synth.images(1, 222, 222, 8)
    with torch.no_grad():
        f = O.encode_relu4_1(vgg, x)
        stat = O.single_style_stats(O.encode_relu4_1(vgg, synth.images(1, 222, 300, 9)))
        ref = O.style_transfer(vgg, dec, x, stat, 1.0)
    assert f.shape[-2:] == (28, 28) and ref.shape[-2:] == (224, 224)
    sd = [t.to(DEV) for t in stat]
    out = ccst_b200.style_transfer(vgg, dec, x.to(DEV), sd, 1.0, precision=precision)
    err = (out.cpu() - ref).abs().max().item()
    report(f"222^2 {precision}", max_abs=err)
    assert err < IMG_TOL[precision]


def test_fused_pool_equals_separate_pool(models):
    from ccst_b200 import _lib
    vgg, dec = models
    x = synth.images(2, 72, 56, 4).to(DEV)
    eng = ccst_b200.Engine(vgg, dec, DEV)
    for prec in ("bf16", "fp16"):
        eng.set_fusion(_lib.FUSE_ALL)
        a = eng.encode(x, prec)
        eng.set_fusion(_lib.FUSE_ALL & ~_lib.FUSE_POOL)
        b = eng.encode(x, prec)
        assert torch.equal(a, b)


def test_fused_upsample_matches_unfused_decoder(models):
    """Upsample folded into the next conv (phase-decomposed 2x2 kernels) vs the 4x-replicating store:
    same function, weights pre-summed and rounded once more, so equal within the 16-bit rounding
    (an implementation-equivalence bound, not a parity bar: both sides are this library)."""
    from ccst_b200 import _lib
    vgg, dec = models
    feat = synth.features((2, 512, 9, 13), 11).to(DEV)
    eng = ccst_b200.Engine(vgg, dec, DEV)
    # (the x3 engines take the same two paths with split operands: equal within their 22 / 16 significand bits)
    for prec, tol in (("fp16", 2e-3), ("bf16", 1.6e-2), ("fp16x3", 5e-6), ("bf16x3", 1e-4)):
        eng.set_fusion(_lib.FUSE_ALL)
        a = eng.decode(feat, prec)
        eng.set_fusion(_lib.FUSE_ALL & ~_lib.FUSE_UPSAMPLE)
        b = eng.decode(feat, prec)
        assert a.shape == b.shape == (2, 3, 72, 104)
        assert (a - b).abs().max().item() < tol


def test_folded_adain_matches_affine_pass(models):
    """AdaIN folded into dec1's per-image weights / bias (maps >= 2048 px) vs the affine pass over the
    feature map: the same function up to one 16-bit rounding (of W*A instead of x*A + B); both within
    the image bar of the oracle, the folded form usually closer (the feature map is not re-rounded)."""
    from ccst_b200 import _lib
    vgg, dec = models
    x = synth.images(3, 448, 320, 31)  # relu4_1 56 x 40 = 2240 px, 7 x 3 = 21 tiles per image: odd (pair padding)
    with torch.no_grad():
        stat = O.single_style_stats(O.encode_relu4_1(vgg, synth.images(1, 200, 180, 32)))
        ref = O.style_transfer(vgg, dec, x, stat, 0.7)
    sd = [t.to(DEV) for t in stat]
    per = [torch.cat([t] * 3).to(DEV) * torch.tensor([1.0, 0.9, 1.1], device=DEV).view(3, 1, 1, 1) for t in stat]
    eng = ccst_b200.Engine(vgg, dec, DEV)
    for prec in ("fp16", "bf16"):
        eng.set_fusion(_lib.FUSE_ALL)
        a = eng.transfer(x.to(DEV), sd, 0.7, prec)
        ap = eng.transfer(x.to(DEV), per, 0.7, prec)  # per-image style statistics ([N,512,1,1])
        eng.set_fusion(_lib.FUSE_ALL & ~_lib.FUSE_ADAIN)
        b = eng.transfer(x.to(DEV), sd, 0.7, prec)
        bp = eng.transfer(x.to(DEV), per, 0.7, prec)
        d, dp = (a - b).abs().max().item(), (ap - bp).abs().max().item()
        ea, eb = (a.cpu() - ref).abs().max().item(), (b.cpu() - ref).abs().max().item()
        report(f"folded AdaIN {prec}", fold_vs_pass=d, per_image=dp, fold_vs_oracle=ea, pass_vs_oracle=eb)
        assert d < (4e-3 if prec == "fp16" else 2.5e-2) and dp < (4e-3 if prec == "fp16" else 2.5e-2)
        if prec == "fp16":
            assert ea < TOL_TC and eb < TOL_TC


def _stat_rel(a, b):
    """max |a - b| relative to the statistic's scale (its largest magnitude over the channels)."""
    b = b.double().flatten()
    return ((a.cpu().double().flatten() - b).abs().max() / b.abs().max()).item()


@pytest.mark.parametrize("precision", ["fp32", "fp16x3", pytest.param("fp16", marks=STATS16_MISS),
                                       pytest.param("bf16", marks=STATS16_MISS)])
def test_overall_statistics_loop(models, precision):
    """mean_std_computation_effcientMem.py:117-137 on 3 batches of images against the reference formula
    evaluated in fp64: 1e-5 relative, per engine."""
    vgg, dec = models
    batches = [synth.images(n, 64, 64, 50 + i) for i, n in enumerate((3, 3, 2))]
    with torch.no_grad():
        feats = [O.encode_relu4_1(vgg, b) for b in batches]
    mean64, std64, count, imgs = O.overall_style_stats(feats, dtype=torch.float64)
    eng = ccst_b200.engine_for(vgg, dec, torch.device(DEV))
    acc = OverallStyleAccumulator(eng, precision)
    for b in batches:
        acc.add_images(b.to(DEV))
    mean, std = acc.finalize()
    assert acc.state.count == count and acc.img_count == imgs
    em, es = _stat_rel(mean, mean64), _stat_rel(std, std64)
    report(f"overall statistics {precision}", mean_rel=em, std_rel=es)
    assert em < TOL_STATS and es < TOL_STATS


def test_overall_statistics_default_engine_and_identical_features(models):
    """The statistics drivers default to an engine that meets 1e-5 (f16x3: split f16 operands on the tensor
    pipe); and given identical features the accumulator itself meets the bar per element."""
    from ccst_b200 import overall
    vgg, dec = models
    assert overall.STATS_PRECISION == "fp16x3"
    batches = [synth.images(n, 64, 64, 50 + i) for i, n in enumerate((3, 3, 2))]
    with torch.no_grad():
        feats = [O.encode_relu4_1(vgg, b) for b in batches]
    mean64, std64, count, imgs = O.overall_style_stats(feats, dtype=torch.float64)
    eng = ccst_b200.engine_for(vgg, dec, torch.device(DEV))
    acc = OverallStyleAccumulator(eng)
    assert acc.precision == "fp16x3"
    for f in feats:
        acc.add_features(f.to(DEV))
    mean, std = acc.finalize()
    rel = lambda a, b: ((a.cpu().double() - b).abs() / (b.abs() + 1e-4)).max().item()
    assert rel(mean, mean64) < TOL_STATS and rel(std, std64) < TOL_STATS


@pytest.mark.parametrize("precision", PREC_IMG)
def test_config1_batch6_512_vs_oracle(models, precision):
    """BASELINE config 1 at its full shape: style_transfer, batch 6 @512x512, against the ORACLE (the
    reference's own PyTorch path on the CPU), every precision at the north_star bar."""
    vgg, dec = models
    x = synth.images(6, 512, 512, 1)
    with torch.no_grad():
        stat = O.single_style_stats(O.encode_relu4_1(vgg, synth.images(1, 512, 512, 2)))
        ref = O.style_transfer(vgg, dec, x, stat, 1.0)
    sd = [t.to(DEV) for t in stat]
    out = ccst_b200.style_transfer(vgg, dec, x.to(DEV), sd, 1.0, precision=precision)
    assert tuple(out.shape) == (6, 3, 512, 512)
    err = (out.cpu() - ref).abs().max().item()
    report(f"config 1 (6 x 512^2) {precision}", max_abs=err, ref_min=ref.min().item(), ref_max=ref.max().item())
    assert err < IMG_TOL[precision]


@pytest.mark.parametrize("precision", PREC_IMG)
def test_config5_batch1024_96_vs_oracle(models, precision):
    """BASELINE config 5 (Camelyon17-like): 96x96 patches, ONE batch of 1024 through the library; the
    oracle is run on every 32nd image (images are independent; the whole batch costs the CPU minutes)
    and the full batch is checked against sub-batches bit for bit."""
    vgg, dec = models
    n = 1024 if precision != "fp32" else 128  # the FFMA validation engine is ~25x slower
    x = synth.images(n, 96, 96, 5)
    with torch.no_grad():
        stat = O.single_style_stats(O.encode_relu4_1(vgg, synth.images(4, 96, 96, 6)))
        idx = torch.arange(0, n, 32)
        ref = O.style_transfer(vgg, dec, x[idx], stat, 1.0)
    sd = [t.to(DEV) for t in stat]
    xd = x.to(DEV)
    out = ccst_b200.style_transfer(vgg, dec, xd, sd, 1.0, precision=precision)
    assert tuple(out.shape) == (n, 3, 96, 96) and torch.isfinite(out).all()
    err = (out[idx.to(DEV)].cpu() - ref).abs().max().item()
    report(f"config 5 ({n} x 96^2) {precision}", max_abs=err, oracle_images=len(idx))
    part = ccst_b200.style_transfer(vgg, dec, xd[96:160].contiguous(), sd, 1.0, precision=precision)
    assert torch.equal(part, out[96:160])
    assert err < IMG_TOL[precision]


def test_full_size_batch_is_deterministic_and_shards(models):
    """512x512 (BASELINE config 3 shape, batch 4): two runs are bit-identical and a batch split in two
    (the multi-GPU sharding unit) reproduces the unsplit result."""
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


def test_f16_saturation_is_reported(models):
    """f16 stores clamp at +-65504: weights that drive activations out of the f16 range must raise,
    not silently produce an image (the counter is kept by every epilogue, one atomic per warp)."""
    import copy
    from ccst_b200.transfer import F16SaturationError
    vgg, dec = models
    x = synth.images(2, 64, 64, 71).to(DEV)
    g = torch.Generator().manual_seed(3)
    stat = [torch.randn((1, 512, 1, 1), generator=g).abs().to(DEV), (torch.rand((1, 512, 1, 1), generator=g) + 0.1).to(DEV)]
    eng = ccst_b200.Engine(vgg, dec, DEV)
    eng.transfer(x, stat, 1.0, "fp16")
    assert eng.saturation_count() == 0
    eng.check_saturation()  # in-range weights: no error
    big = copy.deepcopy(vgg)
    convs = [m for m in big if isinstance(m, torch.nn.Conv2d)]
    with torch.no_grad():
        convs[2].weight.mul_(3.0e5)  # conv1_2: activations ~1e5 > 65504
    eng2 = ccst_b200.Engine(big, dec, DEV)
    eng2.transfer(x, stat, 1.0, "fp16")
    assert eng2.saturation_count() > 0
    with pytest.raises(F16SaturationError):
        eng2.check_saturation()
    assert eng2.saturation_count() == 0  # reset by the check
    eng2.transfer(x, stat, 1.0, "bf16")  # bf16 has the fp32 range: nothing to report
    assert eng2.saturation_count() == 0


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
    # the x3 engines (promoted partial sums handed between the MMA warp and two epilogue groups, conv1_1's builder /
    # epilogue-group rings) and the uint8 entry point (window ring fed from the uint8 rows, table look-ups)
    x_u8 = (x.permute(0, 2, 3, 1) * 255).round().to(torch.uint8).contiguous()
    g = torch.Generator().manual_seed(3)
    stat = [torch.randn((1, 512, 1, 1), generator=g).abs().to(DEV), (torch.rand((1, 512, 1, 1), generator=g) + 0.2).to(DEV)]
    for precision in ("fp16x3", "bf16x3"):
        feats = [eng.encode(x, precision).clone() for _ in range(12)]
        assert all(torch.equal(f, feats[0]) for f in feats[1:]), precision
        imgs = [eng.decode(feats[0], precision).clone() for _ in range(8)]
        assert all(torch.equal(f, imgs[0]) for f in imgs[1:]), precision
    for precision in ("fp16", "fp16x3"):
        outs = [eng.transfer_u8(x_u8, stat, 1.0, precision).clone() for _ in range(12)]
        assert all(torch.equal(o, outs[0]) for o in outs[1:]), precision


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


def test_net_forward_losses_golden(models, golden):
    """SURVEY 8f rank 4, forward only: `ccst_b200.net.Net.forward` (level statistics taken from the arena,
    deterministic MSE kernel) against the losses of the reference's real `Net` class (net.py:138-152)."""
    from ccst_b200 import net as B
    g = golden["f4"]
    vgg, dec = models
    content, style = T(g["content"]).to(DEV), T(g["style"]).to(DEV)
    for precision, tol in (("fp32", 1e-4), ("fp16", 2e-2)):
        model = B.Net(vgg, dec, precision=precision)
        for alpha in (1.0, 0.7):
            lc, ls = model(content, style, alpha)
            rc, rs = float(g[f"loss_c_a{alpha}"]), float(g[f"loss_s_a{alpha}"])
            report(f"Net.forward a={alpha} {precision}", loss_c=lc.item(), ref_c=rc, loss_s=ls.item(), ref_s=rs)
            assert abs(lc.item() - rc) < tol * rc and abs(ls.item() - rs) < tol * rs
    # the level statistics themselves (calc_mean_std on relu1_1 .. relu4_1), fp32 engine, 1e-5 of each level's scale
    eng = ccst_b200.engine_for(vgg, dec, torch.device(DEV))
    feat, stats = eng.encode_levels(style, "fp32")
    for i, (m, s) in enumerate(stats):
        assert _stat_rel(m, T(g[f"style_level{i}_mean"])) < TOL_STATS and _stat_rel(s, T(g[f"style_level{i}_std"])) < TOL_STATS
    with torch.no_grad():
        assert (feat.cpu() - O.encode_relu4_1(vgg, style.cpu())).abs().max().item() < TOL_FP32
    with pytest.raises(AssertionError):  # calc_style_loss's size assert (net.py:131)
        B.Net(vgg, dec)(content, synth.images(2, 48, 48, 1).to(DEV), 1.0)
