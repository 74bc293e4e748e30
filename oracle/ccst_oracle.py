"""CPU oracle for the CCST AdaIN hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a plain PyTorch-CPU / numpy restatement of the reference's
algorithm for the hot path.  It exists so that `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
leg can check and time the CUDA path against it.  Nothing under `ccst_b200/`
may import it; the product has no CPU fallback.

Parity status: PINNED against the reference itself.  The reference ships no
tests or golden vectors, so `tests/golden/make_golden.py`
imports the *real* reference modules from /root/reference (function.py, net.py,
and the AST-extracted `style_transfer` / `calc_sum` of the CCST scripts), runs
them on seeded inputs, and commits the outputs under `tests/golden/*.npz`.
`tests/test_oracle_golden.py` checks every function below against those
vectors (bit-exact for the pure restatements, 1e-6 for fp64 variants).

Each function cites the reference lines it follows.  All arithmetic is fp32
unless the name ends in `_f64`.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-5


# --------------------------------------------------------------------------
# feature statistics
# --------------------------------------------------------------------------
def calc_mean_std(feat: torch.Tensor, eps: float = EPS):
    """function.py:4-13 -- per-(n,c) mean and sqrt(unbiased var + eps)."""
    assert feat.dim() == 4
    n, c = feat.shape[:2]
    flat = feat.reshape(n, c, -1)
    std = (flat.var(dim=2) + eps).sqrt().view(n, c, 1, 1)
    mean = flat.mean(dim=2).view(n, c, 1, 1)
    return mean, std


def calc_mean_std_f64(feat: torch.Tensor, eps: float = EPS):
    """Same formula evaluated in fp64 (ground truth for the 1e-5 bar)."""
    m, s = calc_mean_std(feat.double(), eps)
    return m, s


def calc_mean_std_vector(feat: torch.Tensor, eps: float = EPS):
    """reconstruct_img/test.py:36-46 -- cat(mean, std) as an [N, 2C] vector."""
    m, s = calc_mean_std(feat, eps)
    return torch.cat([m, s], dim=1).squeeze(-1).squeeze(-1)


def calc_sum(feat: torch.Tensor):
    """mean_std_computation_effcientMem.py:103-115 (== CCST_SingleStyleTransfer.py:55-67)
    per-channel sum and sum of squares over N*H*W, plus the element count."""
    feat = feat.detach()
    assert feat.dim() == 4
    n, c, h, w = feat.shape
    per_c = feat.transpose(0, 1).reshape(c, -1)
    s1 = per_c.sum(dim=1).view(1, c, 1, 1)
    s2 = (per_c ** 2).sum(dim=1).view(1, c, 1, 1)
    return s1, s2, n * h * w


def finalize_sums(s1, s2, count: int, eps: float = EPS):
    """mean_std_computation_effcientMem.py:135-137 (== CCST_SingleStyleTransfer.py:201-203)
    mean = S1/n, biased var = S2/n - mean^2, std = sqrt(var + eps)."""
    mean = s1 / float(count)
    var = s2 / float(count) - mean ** 2
    std = torch.sqrt(var + eps)
    return mean, std


def overall_style_stats(feature_batches, eps: float = EPS, dtype=torch.float32):
    """mean_std_computation_effcientMem.py:117-137 -- the running accumulation
    over the batches of one client followed by the finalisation.  `dtype`
    float64 gives the cancellation-free ground truth."""
    tot1, tot2, tot_n, imgs = 0, 0, 0, 0
    for feat in feature_batches:
        imgs += feat.shape[0]
        s1, s2, cnt = calc_sum(feat.to(dtype))
        tot1 = tot1 + s1
        tot2 = tot2 + s2
        tot_n += cnt
    mean, std = finalize_sums(tot1, tot2, tot_n, eps)
    return mean, std, tot_n, imgs


def single_style_stats(style_feat: torch.Tensor, eps: float = EPS):
    """CCST_SingleStyleTransfer.py:199-203 -- stats of one style image's relu4_1."""
    s1, s2, cnt = calc_sum(style_feat)
    return finalize_sums(s1, s2, cnt, eps)


def pack_style_npy(mean: torch.Tensor, std: torch.Tensor) -> np.ndarray:
    """mean_std_computation_effcientMem.py:146 -- array that np.save would write:
    float32 (2,1,C,1,1)."""
    return np.asarray([mean.cpu().numpy(), std.cpu().numpy()])


# --------------------------------------------------------------------------
# AdaIN
# --------------------------------------------------------------------------
def _renorm(content_feat, c_mean, c_std, s_mean, s_std):
    size = content_feat.size()
    normalised = (content_feat - c_mean.expand(size)) / c_std.expand(size)
    return normalised * s_std.expand(size) + s_mean.expand(size)


def adaptive_instance_normalization(content_feat, style_feat):
    """function.py:16-24."""
    assert content_feat.size()[:2] == style_feat.size()[:2]
    s_mean, s_std = calc_mean_std(style_feat)
    c_mean, c_std = calc_mean_std(content_feat)
    return _renorm(content_feat, c_mean, c_std, s_mean, s_std)


def adaIN_StyleStat_ContentFeat(content_feat, style_stat):
    """function.py:26-33 -- style given as precomputed (mean, std)."""
    s_mean, s_std = style_stat
    c_mean, c_std = calc_mean_std(content_feat)
    return _renorm(content_feat, c_mean, c_std, s_mean, s_std)


# --------------------------------------------------------------------------
# encoder / decoder (functional restatement of net.py)
# --------------------------------------------------------------------------
def _conv_list(seq):
    return [m for m in seq if isinstance(m, torch.nn.Conv2d)]


def _rconv(x, conv, relu=True):
    y = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), conv.weight, conv.bias)
    return F.relu(y) if relu else y


def encode_relu4_1(vgg, x):
    """net.py:38-69 -- 1x1 conv, then conv blocks [2,2,4,1] with ceil-mode
    2x2 max-pools between them, ending at relu4_1."""
    convs = _conv_list(vgg)
    assert len(convs) >= 10
    x = F.conv2d(x, convs[0].weight, convs[0].bias)
    idx = 1
    for bi, reps in enumerate((2, 2, 4, 1)):
        if bi:
            x = F.max_pool2d(x, 2, 2, 0, ceil_mode=True)
        for _ in range(reps):
            x = _rconv(x, convs[idx])
            idx += 1
    return x


def decode(decoder, x):
    """net.py:6-36 -- 9 reflect-pad 3x3 convs, nearest x2 after convs 1, 5, 7,
    no ReLU after the last."""
    convs = _conv_list(decoder)
    assert len(convs) == 9
    for i, conv in enumerate(convs):
        x = _rconv(x, conv, relu=(i != 8))
        if i in (0, 4, 6):
            x = F.interpolate(x, scale_factor=2, mode="nearest")
    return x


def style_transfer(vgg, decoder, content, style_stat, alpha=1.0, interpolation_weights=None):
    """CCST_OverallStyleTransfer.py:32-46 (== CCST_SingleStyleTransfer.py:39-53)."""
    assert 0.0 <= alpha <= 1.0
    content_f = encode_relu4_1(vgg, content)
    if interpolation_weights:
        base = adaIN_StyleStat_ContentFeat(content_f, style_stat)
        feat = torch.zeros_like(content_f[0:1])
        for i, w in enumerate(interpolation_weights):
            feat = feat + w * base[i:i + 1]
        content_f = content_f[0:1]
    else:
        feat = adaIN_StyleStat_ContentFeat(content_f, style_stat)
    feat = feat * alpha + content_f * (1 - alpha)
    return decode(decoder, feat)


def style_transfer_image_style(vgg, decoder, content, style, alpha=1.0):
    """Upstream AdaIN form named in BASELINE.json (style given as images);
    follows net.py:138-143 (encode both, adain, alpha blend, decode)."""
    assert 0.0 <= alpha <= 1.0
    content_f = encode_relu4_1(vgg, content)
    style_f = encode_relu4_1(vgg, style)
    feat = adaptive_instance_normalization(content_f, style_f)
    feat = feat * alpha + content_f * (1 - alpha)
    return decode(decoder, feat)


def save_image_quantize(img: torch.Tensor) -> torch.Tensor:
    """torchvision.utils.save_image's tensor->uint8 step
    (CCST_OverallStyleTransfer.py:167): mul(255).add(0.5).clamp(0,255).to(uint8)."""
    return img.mul(255).add_(0.5).clamp_(0, 255).to(torch.uint8)


def to_tensor_u8(images_u8: torch.Tensor) -> torch.Tensor:
    """`transforms.ToTensor()` on the loader's uint8 HWC image (cjm_util/data_helper.py:45;
    torchvision 0.26 functional.to_tensor: `img.permute(2,0,1).to(float32).div(255)`), batched:
    [N,H,W,C] uint8 -> [N,C,H,W] fp32."""
    return images_u8.permute(0, 3, 1, 2).contiguous().to(torch.float32).div(255)


def save_image_batch_u8(images: torch.Tensor) -> torch.Tensor:
    """What `for out_img in output: save_image(out_img, name)` hands to the image encoder
    (CCST_OverallStyleTransfer.py:158-167; torchvision 0.26 utils.save_image:
    `grid.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to("cpu", torch.uint8)`; make_grid of a
    single 3-channel image is the image itself): [N,C,H,W] fp32 -> [N,H,W,C] uint8."""
    return save_image_quantize(images.clone()).permute(0, 2, 3, 1).contiguous()


def resize_output(images: torch.Tensor, size) -> torch.Tensor:
    """`transforms.Resize(args.output_size)` applied to the stylised batch
    (CCST_OverallStyleTransfer.py:134-135,154-155).  torchvision 0.26 `F.resize` of a float tensor is
    `torch.nn.functional.interpolate(mode="bilinear", align_corners=False, antialias=True)` at the
    size `_compute_resized_output_size` gives (an int matches the smaller edge)."""
    n, c, h, w = images.shape
    if isinstance(size, (tuple, list)) and len(size) == 2:
        oh, ow = int(size[0]), int(size[1])
    else:
        s = int(size[0] if isinstance(size, (tuple, list)) else size)
        oh, ow = (s, int(s * w / h)) if h <= w else (int(s * h / w), s)
    if (oh, ow) == (h, w):
        return images
    return torch.nn.functional.interpolate(images, size=(oh, ow), mode="bilinear", align_corners=False,
                                           antialias=True)


def calc_mean_std_batch(feat: torch.Tensor, eps: float = EPS):
    """The batch-wide `calc_mean_std` of mean_std_computation_effcientMem.py:89-101 (defined, never
    called): per channel over N*H*W with torch's default unbiased `var`."""
    assert feat.dim() == 4
    c = feat.shape[1]
    flat = feat.swapaxes(1, 0).reshape(c, -1)
    var = flat.var(dim=1) + eps
    return flat.mean(dim=1).reshape(1, c, 1, 1), torch.sqrt(var.reshape(1, c, 1, 1))


# --------------------------------------------------------------------------
# SURVEY 8f rank 4 (forward only): the training wrapper's losses and MixStyle's statistics
# --------------------------------------------------------------------------
def encode_with_intermediate(vgg, x):
    """Net.encode_with_intermediate (style_transfer/AdaIN/net.py:112-117): relu1_1, relu2_1, relu3_1, relu4_1
    (enc_1 = children[:4], enc_2 = [4:11], enc_3 = [11:18], enc_4 = [18:31], net.py:98-102)."""
    convs = _conv_list(vgg)
    x = F.conv2d(x, convs[0].weight, convs[0].bias)
    feats = []
    idx = 1
    for bi, reps in enumerate((2, 2, 4, 1)):
        if bi:
            x = F.max_pool2d(x, 2, 2, 0, ceil_mode=True)
        for k in range(reps):
            x = _rconv(x, convs[idx])
            idx += 1
            if k == 0:
                feats.append(x)  # relu{bi+1}_1
    return feats


def net_forward_losses(vgg, decoder, content, style, alpha=1.0):
    """Net.forward (net.py:138-152): (loss_c, loss_s), forward values only."""
    assert 0 <= alpha <= 1
    mse = torch.nn.functional.mse_loss
    style_feats = encode_with_intermediate(vgg, style)
    content_feat = encode_relu4_1(vgg, content)
    t = adaptive_instance_normalization(content_feat, style_feats[-1])
    t = alpha * t + (1 - alpha) * content_feat
    g_t = decode(decoder, t)
    g_feats = encode_with_intermediate(vgg, g_t)
    loss_c = mse(g_feats[-1], t)
    loss_s = 0
    for gf, sf in zip(g_feats, style_feats):
        gm, gs = calc_mean_std(gf)
        sm, ss = calc_mean_std(sf)
        loss_s = loss_s + mse(gm, sm) + mse(gs, ss)
    return loss_c, loss_s


def mixstyle_forward(x, lmda, perm, eps: float = 1e-6):
    """MixStyle.forward after its random draws (nets/layers.py:46-74): lmda [B,1,1,1], perm [B]."""
    mu = x.mean(dim=[2, 3], keepdim=True)
    var = x.var(dim=[2, 3], keepdim=True)
    sig = (var + eps).sqrt()
    x_normed = (x - mu) / sig
    mu2, sig2 = mu[perm], sig[perm]
    mu_mix = mu * lmda + mu2 * (1 - lmda)
    sig_mix = sig * lmda + sig2 * (1 - lmda)
    return x_normed * sig_mix + mu_mix


# --------------------------------------------------------------------------
# the loader's Resize((S, S)) on the PIL image (cjm_util/data_helper.py:45-49)
# --------------------------------------------------------------------------
def _pil_bilinear_coeffs(in_size: int, out_size: int):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter (un-vendored dependency of
    the reference: torchvision's Resize calls Image.resize(BILINEAR); restated from Pillow 12.2.0
    src/libImaging/Resample.c and pinned against the real Pillow in tests/golden/io_u8.npz)."""
    import math

    prec = 32 - 8 - 2
    scale = in_size / out_size
    fscale = max(scale, 1.0)
    support = 1.0 * fscale
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize), np.int64)
    bounds = np.zeros((out_size, 2), np.int64)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / fscale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = np.zeros(ksize)
        ww = 0.0
        for x in range(xmax):
            v = abs((x + xmin - center + 0.5) * ss)
            w[x] = 1.0 - v if v < 1.0 else 0.0
            ww += w[x]
        for x in range(xmax):
            if ww != 0.0:
                w[x] /= ww
        for x in range(ksize):
            kk[xx, x] = int(-0.5 + w[x] * (1 << prec)) if w[x] < 0 else int(0.5 + w[x] * (1 << prec))
        bounds[xx] = (xmin, xmax)
    return kk, bounds, prec


def _pil_resample_axis(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    a = np.moveaxis(img, axis, 0).astype(np.int64)
    kk, b, prec = _pil_bilinear_coeffs(a.shape[0], out_size)
    out = np.zeros((out_size,) + a.shape[1:], np.int64)
    for xx in range(out_size):
        xmin, cnt = b[xx]
        ss = np.full(a.shape[1:], 1 << (prec - 1), np.int64)
        for x in range(cnt):
            ss += a[xmin + x] * kk[xx, x]
        out[xx] = np.clip(ss >> prec, 0, 255)
    return np.moveaxis(out.astype(np.uint8), 0, axis)


def pil_resize_bilinear_u8(images_u8: torch.Tensor, oh: int, ow: int) -> torch.Tensor:
    """transforms.Resize((oh, ow)) on PIL images, for a uint8 [N,H,W,C] batch: horizontal pass into an
    8-bit intermediate, then the vertical pass (ImagingResample)."""
    a = images_u8.numpy()
    if ow != a.shape[2]:
        a = _pil_resample_axis(a, ow, 2)
    if oh != a.shape[1]:
        a = _pil_resample_axis(a, oh, 1)
    return torch.from_numpy(np.ascontiguousarray(a))
