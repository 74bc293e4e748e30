"""GPU: the batch loops of the three CCST scripts (ccst_b200.drivers) reproduce per-batch calls and
the oracle's restatement of the reference loops."""
import random

import numpy as np
import pytest
import torch

import ccst_b200
from ccst_b200 import drivers, synth
from oracle import ccst_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def engine(models):
    vgg, dec = models
    return ccst_b200.engine_for(vgg, dec, torch.device(DEV))


def test_overall_transfer_pipeline_equals_per_batch_calls(models, engine):
    """CCST_OverallStyleTransfer.py:149-167 with overlapped copies: same bits as blocking calls, for
    more batches than pipeline slots and a ragged last batch."""
    vgg, dec = models
    batches = [synth.images(n, 64, 80, 300 + i).pin_memory() for i, n in enumerate((3, 3, 3, 3, 2))]
    g = torch.Generator().manual_seed(2)
    stat = [torch.randn((1, 512, 1, 1), generator=g).abs(), torch.rand((1, 512, 1, 1), generator=g) + 0.2]
    got = {}
    for i, out in drivers.overall_transfer(engine, iter(batches), stat, 0.7):
        got[i] = out.clone()
    assert sorted(got) == list(range(5))
    sd = [t.to(DEV) for t in stat]
    for i, b in enumerate(batches):
        ref = ccst_b200.style_transfer(vgg, dec, b.to(DEV), sd, 0.7)
        assert torch.equal(got[i], ref.cpu())


def test_single_transfer_matches_reference_loop(models, engine):
    """CCST_SingleStyleTransfer.py:176-223: one random style image per batch (python RNG seed 1)."""
    vgg, dec = models
    batches = [synth.images(2, 48, 48, 400 + i).pin_memory() for i in range(3)]
    styles = [synth.images(1, 56, 40 + 8 * k, 500 + k) for k in range(4)]
    rng = random.Random(1)
    outs = {}
    for i, o in drivers.single_transfer(engine, iter(batches), styles, 1.0, "fp32", seed=1):
        outs[i] = o.clone()  # pipeline buffers are reused: keep a copy
    with torch.no_grad():
        for i, b in enumerate(batches):
            img = rng.choice(styles)
            stat = O.single_style_stats(O.encode_relu4_1(vgg, img))
            ref = O.style_transfer(vgg, dec, b, stat, 1.0)
            assert (outs[i] - ref).abs().max().item() < 1e-4


def test_overall_statistics_loop_with_uploads(models, engine):
    vgg, dec = models
    batches = [synth.images(n, 64, 64, 600 + i).pin_memory() for i, n in enumerate((4, 4, 3))]
    with torch.no_grad():
        feats = [O.encode_relu4_1(vgg, b) for b in batches]
    mean64, std64, count, imgs = O.overall_style_stats(feats, dtype=torch.float64)
    mean, std, seen = drivers.overall_statistics(engine, iter(batches), "fp32")
    assert seen == imgs
    assert (mean.cpu().double() - mean64).abs().max().item() < 1e-4 * max(1.0, mean64.abs().max().item())
    assert (std.cpu().double() - std64).abs().max().item() < 1e-4 * max(1.0, std64.abs().max().item())


def test_to_tensor_and_quantize_operators_bit_exact(golden):
    """ToTensor (float(u)/255) and save_image's quantisation on the GPU: bit-exact vs torchvision."""
    g = golden["io_u8"]
    x = ccst_b200.to_tensor_u8(torch.from_numpy(g["x_u8"]).to(DEV))
    assert torch.equal(x.cpu(), torch.from_numpy(g["x_tensor"]))
    for a in ("1.0", "0.5"):
        q = ccst_b200.save_image_quantize(torch.from_numpy(g[f"out_f32_a{a}"]).to(DEV))
        assert torch.equal(q.cpu(), torch.from_numpy(g[f"out_u8_a{a}"]))
    # every uint8 value and the rounding boundaries k/255 - tiny .. k/255 + tiny
    u = torch.arange(256, dtype=torch.uint8).view(1, 16, 16, 1).to(DEV)
    t = ccst_b200.to_tensor_u8(u)
    assert torch.equal(t.cpu(), O.to_tensor_u8(u.cpu()))
    assert torch.equal(ccst_b200.save_image_quantize(t).cpu(), u.cpu())
    v = torch.cat([t - 0.5 / 255, t - 0.4999 / 255, t + 0.4999 / 255, t * 3 - 1])
    assert torch.equal(ccst_b200.save_image_quantize(v).cpu(), O.save_image_batch_u8(v.cpu()))


BF16_MISS = pytest.mark.xfail(strict=True, reason="bf16 operands miss the 1e-2 image bar (DESIGN.md Numerics)")


@pytest.mark.parametrize("precision", ["fp32", "fp16x3", "fp16", "bf16x3", pytest.param("bf16", marks=BF16_MISS)])
def test_style_transfer_u8_golden(models, golden, precision):
    """uint8 in -> uint8 out against the reference pipeline (ToTensor, style_transfer, save_image; the
    vector's float output lies inside [0,1]): fp32 engine within one grey level (rounding ties) and
    almost always exact; tensor-core engines within the 1e-2 image tolerance = 3 grey levels."""
    vgg, dec = models
    g = golden["io_u8"]
    assert 0.0 <= g["out_f32_a1.0"].min() and g["out_f32_a1.0"].max() <= 1.0
    x_u8 = torch.from_numpy(g["x_u8"]).to(DEV)
    stat = [torch.from_numpy(g["style_mean"]).to(DEV), torch.from_numpy(g["style_std"]).to(DEV)]
    ref = torch.from_numpy(g["out_u8_a1.0"]).int()
    max_lv, min_exact = {"fp32": (1, 0.995), "fp16x3": (1, 0.995), "fp16": (3, 0.6), "bf16x3": (1, 0.95),
                         "bf16": (3, 0.0)}[precision]
    out = ccst_b200.style_transfer_u8(vgg, dec, x_u8, stat, 1.0, precision=precision)
    assert out.dtype == torch.uint8 and tuple(out.shape) == tuple(ref.shape)
    d = (out.cpu().int() - ref).abs()
    print(f"[measured] u8 golden {precision}: max {d.max().item()} grey levels, exact {(d == 0).float().mean().item():.3f}")
    assert d.max().item() <= max_lv, (precision, d.max().item())
    assert (d == 0).float().mean().item() >= min_exact, (precision, (d == 0).float().mean().item())


def test_pipeline_result_survives_one_more_step(models, engine):
    """A yielded pinned buffer stays valid while the NEXT result is produced and handed out (slots + 1
    host buffers): materialising pairs of consecutive results needs no clone."""
    vgg, dec = models
    batches = [synth.images(2, 48, 64, 900 + i).pin_memory() for i in range(5)]
    g = torch.Generator().manual_seed(5)
    stat = [torch.randn((1, 512, 1, 1), generator=g).abs(), torch.rand((1, 512, 1, 1), generator=g) + 0.2]
    sd = [t.to(DEV) for t in stat]
    refs = [ccst_b200.style_transfer(vgg, dec, b.to(DEV), sd, 1.0).cpu() for b in batches]
    prev = None
    for i, out in drivers.overall_transfer(engine, iter(batches), stat, 1.0):
        if prev is not None:
            assert torch.equal(prev[1], refs[prev[0]])  # still intact after the next hand-out
        assert torch.equal(out, refs[i])
        prev = (i, out)


def test_pipeline_is_kept_per_engine_and_reusable(models, engine):
    """The drivers keep one pipeline per (engine, precision, u8): a second loop reuses its device slots and pinned
    buffers (also with another batch shape) and gives the same results; a loop that is still open gets its own."""
    vgg, dec = models
    g = torch.Generator().manual_seed(12)
    stat = [torch.randn((1, 512, 1, 1), generator=g).abs(), torch.rand((1, 512, 1, 1), generator=g) + 0.2]
    sd = [t.to(DEV) for t in stat]
    a = [synth.images(2, 48, 64, 950 + i).pin_memory() for i in range(3)]
    b = [synth.images(3, 32, 80, 960 + i).pin_memory() for i in range(2)]
    p0 = drivers.TransferPipeline.for_engine(engine, drivers.DEFAULT_PRECISION)
    for batches in (a, b, a):
        refs = [ccst_b200.style_transfer(vgg, dec, x.to(DEV), sd, 1.0).cpu() for x in batches]
        for i, out in drivers.overall_transfer(engine, iter(batches), stat, 1.0):
            assert torch.equal(out, refs[i])
        assert drivers.TransferPipeline.for_engine(engine, drivers.DEFAULT_PRECISION) is p0
    it = drivers.overall_transfer(engine, iter(a), stat, 1.0)
    next(it)  # open loop: the cached pipeline is busy
    assert drivers.TransferPipeline.for_engine(engine, drivers.DEFAULT_PRECISION) is not p0
    it.close()


def test_style_transfer_u8_equals_float_path_quantised(models):
    """Integer work is bit-exact: the fused uint8 path == ToTensor -> style_transfer -> quantise,
    all on the same engine, for every precision, on a ragged non-multiple-of-8 size."""
    vgg, dec = models
    g = torch.Generator().manual_seed(9)
    x_u8 = torch.randint(0, 256, (3, 52, 70, 3), generator=g, dtype=torch.uint8).to(DEV)
    stat = [torch.randn((1, 512, 1, 1), generator=g).abs().to(DEV), (torch.rand((1, 512, 1, 1), generator=g) + 0.2).to(DEV)]
    for prec in ("fp32", "fp16", "bf16", "fp16x3", "bf16x3"):
        a = ccst_b200.style_transfer_u8(vgg, dec, x_u8, stat, 0.8, precision=prec)
        f = ccst_b200.style_transfer(vgg, dec, ccst_b200.to_tensor_u8(x_u8), stat, 0.8, precision=prec)
        assert tuple(a.shape) == (3, 56, 72, 3)
        assert torch.equal(a, ccst_b200.save_image_quantize(f))
        assert torch.equal(a.cpu(), O.save_image_batch_u8(f.cpu()))


@pytest.mark.parametrize("hw", [(48, 64), (40, 144), (24, 256), (16, 400)])
def test_fused_totensor_equals_conversion_pass(models, engine, hw):
    """ToTensor fused into conv1_1's window loader (the uint8 HWC rows fetched by TMA, float(b)/255 from a table) ==
    the separate conversion pass, bit for bit, for every tensor-core engine; widths with one partial 128-pixel
    tile, several tiles, and a ragged last tile (W % 16 == 0 is what the fused loader needs)."""
    from ccst_b200 import _lib
    vgg, dec = models
    h, w = hw
    g = torch.Generator().manual_seed(h * 1000 + w)
    x_u8 = torch.randint(0, 256, (3, h, w, 3), generator=g, dtype=torch.uint8).to(DEV)
    x_u8[0, :, :, :] = 255  # the table's last entry
    x_u8[1, 0, :, 0] = 0
    stat = [torch.randn((1, 512, 1, 1), generator=g).abs().to(DEV), (torch.rand((1, 512, 1, 1), generator=g) + 0.2).to(DEV)]
    try:
        for prec in ("fp16", "bf16", "fp16x3", "bf16x3"):
            engine.set_fusion(_lib.FUSE_ALL)
            a = ccst_b200.style_transfer_u8(vgg, dec, x_u8, stat, 0.9, precision=prec)
            st_a = ccst_b200.function.WelfordState(512, torch.device(DEV))
            engine.accumulate_u8(x_u8, st_a, prec)
            engine.set_fusion(_lib.FUSE_ALL & ~_lib.FUSE_TOTENSOR)
            b = ccst_b200.style_transfer_u8(vgg, dec, x_u8, stat, 0.9, precision=prec)
            st_b = ccst_b200.function.WelfordState(512, torch.device(DEV))
            engine.accumulate_u8(x_u8, st_b, prec)
            assert torch.equal(a, b), (prec, hw, (a.int() - b.int()).abs().max().item())
            assert torch.equal(st_a.buf, st_b.buf), (prec, hw)
    finally:
        engine.set_fusion(_lib.FUSE_ALL)


def test_overall_transfer_u8_pipeline(models, engine):
    """The overlapped batch loop with uint8 batches equals per-batch calls."""
    vgg, dec = models
    g = torch.Generator().manual_seed(10)
    batches = [torch.randint(0, 256, (n, 64, 48, 3), generator=g, dtype=torch.uint8).pin_memory() for n in (3, 3, 2)]
    stat = [torch.randn((1, 512, 1, 1), generator=g).abs(), torch.rand((1, 512, 1, 1), generator=g) + 0.2]
    sd = [t.to(DEV) for t in stat]
    got = {i: out.clone() for i, out in drivers.overall_transfer(engine, iter(batches), stat, 1.0, u8=True)}
    assert sorted(got) == [0, 1, 2]
    for i, b in enumerate(batches):
        assert torch.equal(got[i], ccst_b200.style_transfer_u8(vgg, dec, b.to(DEV), sd, 1.0).cpu())


def test_resize_output_golden(golden):
    """Anti-aliased bilinear `transforms.Resize` of the output tensor on the GPU vs the real
    torchvision (down-size, up-size, non-square, the Camelyon 512 -> 96 case); fp32, 1e-5."""
    g = golden["io_u8"]
    o = torch.from_numpy(g["out_f32_a1.0"]).to(DEV)
    for tag, size in (("s24", 24), ("s17", 17), ("s64", 64), ("hw", (20, 31))):
        ref = torch.from_numpy(g[f"resize_{tag}"])
        r = ccst_b200.resize(o, size)
        assert tuple(r.shape) == tuple(ref.shape)
        assert (r.cpu() - ref).abs().max().item() < 1e-5
        q = ccst_b200.save_image_quantize(r).cpu().int()
        assert (q - torch.from_numpy(g[f"resize_{tag}_u8"]).int()).abs().max().item() <= 1
    big = torch.rand((1, 3, 512, 512), generator=torch.Generator().manual_seed(77)).to(DEV)
    r = ccst_b200.resize(big, 96)
    assert (r.cpu() - torch.from_numpy(g["resize_512_to_96"])).abs().max().item() < 1e-5
    assert ccst_b200.resize(big, 512) is not None and ccst_b200.resize(big, 512).shape == big.shape
    # properties at the full batch size: constants stay constant, the op is linear
    x = torch.rand((32, 3, 512, 512), device=DEV)
    c = torch.full((2, 3, 512, 512), 0.37, device=DEV)
    assert (ccst_b200.resize(c, 96) - 0.37).abs().max().item() < 1e-6
    a, b = ccst_b200.resize(x, 96), ccst_b200.resize(x * 0.5 + 0.25, 96)
    assert tuple(a.shape) == (32, 3, 96, 96)
    assert (b - (a * 0.5 + 0.25)).abs().max().item() < 1e-5
    assert (a.cpu() - O.resize_output(x.cpu(), 96)).abs().max().item() < 1e-5


def test_overall_statistics_u8_equals_float_path(models, engine):
    """mean_std_computation_effcientMem.py:117-137 fed with the loader's uint8 images: same bits as
    feeding ToTensor'd floats (the conversion is exact), through the double-buffered loop."""
    g = torch.Generator().manual_seed(12)
    b8 = [torch.randint(0, 256, (n, 64, 72, 3), generator=g, dtype=torch.uint8).pin_memory() for n in (3, 2, 3)]
    bf = [O.to_tensor_u8(b).pin_memory() for b in b8]
    m8, s8, c8 = drivers.overall_statistics(engine, iter(b8))
    mf, sf, cf = drivers.overall_statistics(engine, iter(bf))
    assert c8 == cf == 8
    assert torch.equal(m8, mf) and torch.equal(s8, sf)


def test_single_transfer_per_image_sampling(models, engine):
    """Config 4 as BASELINE.json words it: one style image drawn per IMAGE; equals running the
    reference's statistic (CCST_SingleStyleTransfer.py:199-203) and style_transfer image by image."""
    vgg, dec = models
    batches = [synth.images(3, 48, 48, 700 + i).pin_memory() for i in range(2)]
    for styles in ([synth.images(1, 56, 56, 800 + k) for k in range(4)],                 # same size: batched
                   [synth.images(1, 56, 40 + 8 * k, 810 + k) for k in range(4)]):       # ragged: one by one
        rng = random.Random(3)
        outs = {}
        for i, o in drivers.single_transfer_per_image(engine, iter(batches), styles, 1.0, "fp32", seed=3):
            outs[i] = o.clone()
        with torch.no_grad():
            for i, b in enumerate(batches):
                for j in range(b.shape[0]):
                    stat = O.single_style_stats(O.encode_relu4_1(vgg, rng.choice(styles)))
                    ref = O.style_transfer(vgg, dec, b[j:j + 1], stat, 1.0)
                    assert (outs[i][j:j + 1] - ref).abs().max().item() < 1e-4


def test_input_resize_is_bit_exact_with_pillow(golden, tmp_path):
    """`transforms.Resize((S, S))` of the loader (data_helper.py:45-49) on the GPU, bit-exact
    with the real Pillow (golden) and with the oracle at PACS size for a whole batch; then the list reader +
    host decode + GPU resize path end to end on PNG files."""
    import hashlib
    from PIL import Image
    from ccst_b200 import data
    g = golden["io_u8"]
    for tag, S in (("up", 64), ("down", 48), ("pacs", 512)):
        x = torch.from_numpy(g[f"resize_in_{tag}/x"])[None].to(DEV)
        y = ccst_b200.resize_input_u8(x, S)[0].cpu().numpy()
        if tag == "pacs":
            assert hashlib.sha256(y.tobytes()).hexdigest().encode() == g[f"resize_in_{tag}/sha256"].tobytes()
        else:
            assert np.array_equal(y, g[f"resize_in_{tag}/y"])
    gen = torch.Generator().manual_seed(3)
    xb = torch.randint(0, 256, (8, 227, 227, 3), generator=gen, dtype=torch.uint8)
    yb = ccst_b200.resize_input_u8(xb.to(DEV), 512)
    assert torch.equal(yb.cpu(), O.pil_resize_bilinear_u8(xb, 512, 512))
    assert ccst_b200.resize_input_u8(xb.to(DEV), 227).shape == xb.shape  # same size: the image itself
    non_sq = torch.randint(0, 256, (2, 40, 90, 3), generator=gen, dtype=torch.uint8)
    assert torch.equal(ccst_b200.resize_input_u8(non_sq.to(DEV), (30, 100)).cpu(), O.pil_resize_bilinear_u8(non_sq, 30, 100))
    # list file -> PIL decode on the host -> upload at the original size -> Resize on the GPU
    lines = []
    for i, (h, w) in enumerate(((50, 70), (64, 64), (33, 41))):
        im = torch.randint(0, 256, (h, w, 3), generator=gen, dtype=torch.uint8).numpy()
        Image.fromarray(im).save(tmp_path / f"img{i}.png")
        lines.append(f"img{i}.png {i}\n")
    (tmp_path / "list.txt").write_text("".join(lines))
    from torchvision import transforms
    tr = transforms.Resize((48, 48))
    got = list(data.list_batches(str(tmp_path / "list.txt"), str(tmp_path), 48, 2, DEV))
    assert [b.shape[0] for b, _ in got] == [2, 1]
    flat = torch.cat([b for b, _ in got]).cpu().numpy()
    for i in range(3):
        ref = np.asarray(tr(Image.open(tmp_path / f"img{i}.png").convert("RGB")))
        assert np.array_equal(flat[i], ref)
