"""AdaIN operators -- drop-in for `style_transfer/AdaIN/function.py`.

Same names, argument order and assertion behaviour as the reference:

* ``calc_mean_std(feat, eps=1e-5)``                       function.py:4-13
* ``adaptive_instance_normalization(content, style)``     function.py:16-24
* ``adaIN_StyleStat_ContentFeat(content, style_stat)``    function.py:26-33
* ``calc_sum(feat)``   mean_std_computation_effcientMem.py:103-115,
                       CCST_SingleStyleTransfer.py:55-67

All of them run hand-written sm_100a kernels through the C ABI of
libccst_b200.so on the tensor's CUDA device and current stream.  Inputs are
NCHW float32 CUDA tensors; outputs are new tensors.  There is no CPU path.
"""
from __future__ import annotations

import torch

from . import _lib

EPS = 1e-5


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _prep(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} is on {t.device}: ccst_b200 runs only on B200 GPUs (no CPU fallback)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t.detach().contiguous()


def calc_mean_std(feat, eps=EPS):
    """Per-(n,c) mean and sqrt(unbiased var + eps), each [N,C,1,1]."""
    size = feat.size()
    assert (len(size) == 4)
    feat = _prep(feat, "feat")
    n, c = size[:2]
    hw = size[2] * size[3]
    mean = torch.empty((n, c, 1, 1), dtype=torch.float32, device=feat.device)
    std = torch.empty_like(mean)
    with _lib.on_device(feat.device):
        _lib.check(_lib.lib().ccst_stats_nchw_f32(
            feat.data_ptr(), n * c, hw, float(eps), 1, mean.data_ptr(), std.data_ptr(), _stream(feat)))
    return mean, std


def calc_mean_std_biased(feat, eps=EPS):
    """Per-(n,c) mean and sqrt(BIASED var + eps), each [N,C,1,1]: the single-style statistic of
    CCST_SingleStyleTransfer.py:199-203 (`calc_sum` + `var = E[x^2] - mean^2`) for every image of a
    batch of style features at once."""
    size = feat.size()
    assert (len(size) == 4)
    feat = _prep(feat, "feat")
    n, c = size[:2]
    mean = torch.empty((n, c, 1, 1), dtype=torch.float32, device=feat.device)
    std = torch.empty_like(mean)
    with _lib.on_device(feat.device):
        _lib.check(_lib.lib().ccst_stats_nchw_f32(
            feat.data_ptr(), n * c, size[2] * size[3], float(eps), 0, mean.data_ptr(), std.data_ptr(),
            _stream(feat)))
    return mean, std


def calc_mean_std_batch(feat, eps=EPS):
    """The second `calc_mean_std` of the reference (mean_std_computation_effcientMem.py:89-101, defined
    but never called): per CHANNEL over the whole batch N*H*W, unbiased variance -> each [1,C,1,1]."""
    size = feat.shape
    assert (len(size) == 4)
    state = WelfordState(size[1], feat.device)
    state.add_features(feat)
    mean = torch.empty((1, state.C, 1, 1), dtype=torch.float32, device=state.device)
    std = torch.empty_like(mean)
    with _lib.on_device(state.device):
        _lib.check(_lib.lib().ccst_welford_finalize_unbiased(
            state.buf.data_ptr(), state.C, float(eps), mean.data_ptr(), std.data_ptr(),
            torch.cuda.current_stream(state.device).cuda_stream))
    return mean, std


def calc_mean_std_vector(feat, eps=EPS):
    """`calc_mean_std` variant of reconstruct_img/test.py:36-46: cat(mean, std) -> [N, 2C]."""
    mean, std = calc_mean_std(feat, eps)
    return torch.cat([mean, std], dim=1).squeeze(-1).squeeze(-1)


def _style_stat_args(style_stat, n, c, device):
    style_mean, style_std = style_stat
    out = []
    stride = None
    for t, nm in ((style_mean, "style_mean"), (style_std, "style_std")):
        t = _prep(torch.as_tensor(t, device=device) if not isinstance(t, torch.Tensor) else t, nm)
        if t.device != device:
            raise RuntimeError(f"{nm} is on {t.device}, content on {device}")
        if t.numel() == c:
            s = 0
        elif t.numel() == n * c and t.dim() >= 2 and t.shape[0] == n:
            s = c
        else:
            raise RuntimeError(
                f"{nm} of shape {tuple(t.shape)} does not broadcast against [N={n}, C={c}, H, W]")
        if stride is None:
            stride = s
        elif stride != s:
            raise RuntimeError("style mean and std must have the same shape")
        out.append(t.reshape(-1))
    return out[0], out[1], stride


def adain_blend(content_feat, style_stat, alpha=1.0, eps=EPS):
    """alpha * AdaIN(content; style_stat) + (1 - alpha) * content in one pass over HBM
    (function.py:26-33 fused with CCST_OverallStyleTransfer.py:45)."""
    size = content_feat.size()
    assert (len(size) == 4)
    assert (0.0 <= alpha <= 1.0)
    x = _prep(content_feat, "content_feat")
    n, c = size[:2]
    hw = size[2] * size[3]
    mu, sg, stride = _style_stat_args(style_stat, n, c, x.device)
    out = torch.empty_like(x)
    with _lib.on_device(x.device):
        _lib.check(_lib.lib().ccst_adain_stat_nchw_f32(
            x.data_ptr(), n, c, hw, mu.data_ptr(), sg.data_ptr(), stride, float(alpha), float(eps),
            out.data_ptr(), _stream(x)))
    return out


def adaIN_StyleStat_ContentFeat(content_feat, style_stat):
    """AdaIN with precomputed style statistics ``style_stat = (mean, std)``, each
    [1,C,1,1] (broadcast over the batch, as the CCST drivers pass it) or [N,C,1,1]."""
    return adain_blend(content_feat, style_stat, 1.0)


def adaptive_instance_normalization(content_feat, style_feat, alpha=1.0):
    """AdaIN with style features; style H x W may differ from content."""
    assert (content_feat.size()[:2] == style_feat.size()[:2])
    assert (len(content_feat.size()) == 4)
    assert (0.0 <= alpha <= 1.0)
    x = _prep(content_feat, "content_feat")
    s = _prep(style_feat, "style_feat")
    if s.device != x.device:
        raise RuntimeError("content_feat and style_feat must be on the same device")
    n, c = x.shape[:2]
    out = torch.empty_like(x)
    scratch = torch.empty((2 * n * c,), dtype=torch.float32, device=x.device)
    with _lib.on_device(x.device):
        _lib.check(_lib.lib().ccst_adain_feat_nchw_f32(
            x.data_ptr(), s.data_ptr(), n, c, x.shape[2] * x.shape[3], s.shape[2] * s.shape[3],
            float(alpha), EPS, out.data_ptr(), scratch.data_ptr(), _stream(x)))
    return out


def mse_loss(a, b):
    """nn.MSELoss()(a, b) (net.py:104): mean of squared differences as a 0-dim tensor; squares in fp32,
    the sum in fp64 in a fixed order."""
    assert (a.size() == b.size())
    a, b = _prep(a, "a"), _prep(b, "b")
    out = torch.empty((1,), dtype=torch.float32, device=a.device)
    scratch = torch.empty((1024,), dtype=torch.float64, device=a.device)
    with _lib.on_device(a.device):
        _lib.check(_lib.lib().ccst_mse_f32(a.data_ptr(), b.data_ptr(), a.numel(), scratch.data_ptr(), out.data_ptr(),
                                           _stream(a)))
    return out[0]


def mixstyle_stats(x, eps=1e-6):
    """The statistics of MixStyle (nets/layers.py:46-49): mu = x.mean(dim=[2,3]), sig = sqrt(x.var(dim=[2,3])
    + eps) with torch's default unbiased variance and MixStyle's eps = 1e-6 -- calc_mean_std with another eps."""
    return calc_mean_std(x, eps)


def mixstyle(x, lmda, perm, eps=1e-6):
    """MixStyle's forward given its random draws (nets/layers.py:46-74): statistics of x, mixed with those
    of x[perm] by lmda [B,1,1,1], re-applied to the normalised x -- one fused statistics + affine pass."""
    mu, sig = mixstyle_stats(x, eps)
    lmda = lmda.to(x.device, torch.float32).view(-1, 1, 1, 1)
    perm = perm.to(x.device)
    mu_mix = mu * lmda + mu[perm] * (1 - lmda)      # [B,C,1,1]: a few KB of host-side glue
    sig_mix = sig * lmda + sig[perm] * (1 - lmda)
    return adain_blend(x, (mu_mix, sig_mix), 1.0, eps)


class WelfordState:
    """Device-resident running {count, mean[C], M2[C]} (fp64) of one client's features:
    the numerically safe replacement of `all_feat_sum, all_feat_square_sum, all_count`
    (mean_std_computation_effcientMem.py:117).  The buffer is 2 + 2C doubles: the last word is the
    ticket the accumulating kernels use to publish the new count (see ccst_b200.h)."""

    def __init__(self, channels: int, device):
        self.C = int(channels)
        self.device = torch.device(device)
        self.buf = torch.zeros((2 + 2 * self.C,), dtype=torch.float64, device=self.device)

    def reset(self):
        self.buf.zero_()

    def add_features(self, feat):
        """calc_sum(feat) + the three `+=` lines (:126-131) in one pass over `feat`."""
        size = feat.size()
        assert (len(size) == 4)
        feat = _prep(feat, "feat")
        n, c, h, w = size
        if c != self.C:
            raise RuntimeError(f"feature has {c} channels, state has {self.C}")
        scratch = torch.empty((2 * n * c,), dtype=torch.float32, device=feat.device)
        with _lib.on_device(feat.device):
            _lib.check(_lib.lib().ccst_welford_accumulate_nchw_f32(
                feat.data_ptr(), n, c, h * w, self.buf.data_ptr(), scratch.data_ptr(), _stream(feat)))
        return self

    @property
    def count(self) -> int:
        return int(round(self.buf[0].item()))

    def finalize(self, eps=EPS):
        """(mean, std) [1,C,1,1] fp32 with the biased variance of :135-137."""
        mean = torch.empty((1, self.C, 1, 1), dtype=torch.float32, device=self.device)
        std = torch.empty_like(mean)
        with _lib.on_device(self.device):
            _lib.check(_lib.lib().ccst_welford_finalize(
                self.buf.data_ptr(), self.C, float(eps), mean.data_ptr(), std.data_ptr(),
                torch.cuda.current_stream(self.device).cuda_stream))
        return mean, std

    def sums(self):
        """(sum, square_sum) [1,C,1,1] fp32 -- what calc_sum would have accumulated."""
        s1 = torch.empty((1, self.C, 1, 1), dtype=torch.float32, device=self.device)
        s2 = torch.empty_like(s1)
        with _lib.on_device(self.device):
            _lib.check(_lib.lib().ccst_welford_to_sums(
                self.buf.data_ptr(), self.C, s1.data_ptr(), s2.data_ptr(),
                torch.cuda.current_stream(self.device).cuda_stream))
        return s1, s2

    def moments(self):
        """Exactly-summable fp64 vector {n, n*mean, M2 + n*mean^2} (the all-reduce payload)."""
        m = torch.empty((1 + 2 * self.C,), dtype=torch.float64, device=self.device)
        with _lib.on_device(self.device):
            _lib.check(_lib.lib().ccst_welford_to_moments(
                self.buf.data_ptr(), self.C, m.data_ptr(),
                torch.cuda.current_stream(self.device).cuda_stream))
        return m

    def load_moments(self, moments):
        moments = moments.to(device=self.device, dtype=torch.float64).contiguous()
        with _lib.on_device(self.device):
            _lib.check(_lib.lib().ccst_welford_from_moments(
                moments.data_ptr(), self.C, self.buf.data_ptr(),
                torch.cuda.current_stream(self.device).cuda_stream))
        return self


def calc_sum(feat):
    """Per-channel (sum, square_sum, count) over N*H*W, shapes [1,C,1,1] + python int."""
    size = feat.shape
    assert (len(size) == 4)
    n, c, h, w = size
    st = WelfordState(c, feat.device).add_features(feat)
    s1, s2 = st.sums()
    return s1, s2, n * h * w
