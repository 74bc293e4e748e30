"""`style_transfer(vgg, decoder, content, style, alpha)` on B200.

Drop-in for the driver function of the CCST scripts
(`CCST_OverallStyleTransfer.py:32-46` == `CCST_SingleStyleTransfer.py:39-53`):
same positional arguments, same assertion, same result tensor (NCHW fp32 on the
content's device).  `vgg` / `decoder` are the caller's ``nn.Sequential`` objects
(the reference's `net.vgg[:31]` / `net.decoder` or `ccst_b200.net`'s); they are
used as weight containers only -- their conv weights are packed once per
(model, weight version) into a `ccst_handle` and all arithmetic runs in
libccst_b200.so.

`style` may be
  * ``[mean, std]`` (each [1,512,1,1] or [N,512,1,1]) -- the CCST form
    (`adaIN_StyleStat_ContentFeat`), or
  * an image batch [N,3,h,w] -- the upstream AdaIN form named in BASELINE.json
    (`adaptive_instance_normalization` on the encoded style, net.py:138-143).
"""
from __future__ import annotations

import ctypes as C
import weakref

import torch
import torch.nn as nn

from . import _lib
from . import function as F_

PRECISIONS = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16, "fp16": _lib.PREC_FP16,
              # x3 engines: split operands (hi + lo halves) on the tensor pipe.  "fp16x3" = 22 significand bits
              # (fp32-grade: statistics 1e-5, images 1e-4), "bf16x3" = 16 bits at the fp32 exponent range
              "fp16x3": _lib.PREC_FP16X3, "bf16x3": _lib.PREC_BF16X3}
# tensor-core path with f16 operands: meets the 1e-2 image tolerance of BASELINE.json; "bf16" runs
# the same kernels with bf16 operands (wider range, ~7x larger rounding error), "fp32" the FFMA
# validation mode (1e-4)
DEFAULT_PRECISION = "fp16"


class F16SaturationError(RuntimeError):
    """Activations left the f16 range on the precision="fp16" path (stores clamp at +-65504)."""


def _conv_params(seq: nn.Module, count: int, what: str):
    convs = [m for m in seq.modules() if isinstance(m, nn.Conv2d)]
    if len(convs) < count:
        raise RuntimeError(f"{what}: expected at least {count} Conv2d layers, found {len(convs)}")
    return convs[:count]


_ENC_SHAPES = [(3, 3, 1)] + [(o, i, 3) for i, o in
                             ((3, 64), (64, 64), (64, 128), (128, 128), (128, 256), (256, 256),
                              (256, 256), (256, 256), (256, 512))]
_DEC_SHAPES = [(o, i, 3) for i, o in
               ((512, 256), (256, 256), (256, 256), (256, 256), (256, 128), (128, 128), (128, 64),
                (64, 64), (64, 3))]


def _host_arrays(convs, shapes, what):
    ws, bs = [], []
    for k, (conv, (o, i, ks)) in enumerate(zip(convs, shapes)):
        w = conv.weight.detach()
        if tuple(w.shape) != (o, i, ks, ks):
            raise RuntimeError(f"{what}: conv {k} has weight {tuple(w.shape)}, expected {(o, i, ks, ks)}")
        b = conv.bias.detach() if conv.bias is not None else torch.zeros(o)
        ws.append(w.to("cpu", torch.float32).contiguous())
        bs.append(b.to("cpu", torch.float32).contiguous())
    return ws, bs


def _ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for k, t in enumerate(tensors):
        arr[k] = t.data_ptr()
    return arr


def _version(seq: nn.Module):
    return tuple((p.data_ptr(), p._version) for p in seq.parameters())


class Engine:
    """One `ccst_handle` (packed weights + activation arena) on one GPU."""

    def __init__(self, vgg: nn.Module = None, decoder: nn.Module = None, device=None):
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ccst_b200.Engine needs a CUDA (B200) device; there is no CPU fallback")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        with _lib.on_device(self.device):
            torch.cuda.init()
            self._h = _lib.lib().ccst_create(idx)
        if not self._h:
            raise _lib.CcstError(_lib.ECUDA, _lib.lib().ccst_last_error().decode())
        self._finalizer = weakref.finalize(self, _lib.lib().ccst_destroy, self._h)
        self._enc_version = None
        self._dec_version = None
        if vgg is not None:
            self.set_encoder(vgg)
        if decoder is not None:
            self.set_decoder(decoder)

    # -- weights -----------------------------------------------------------
    def set_encoder(self, vgg: nn.Module):
        ver = _version(vgg)
        if ver == self._enc_version:
            return
        ws, bs = _host_arrays(_conv_params(vgg, 10, "vgg"), _ENC_SHAPES, "vgg")
        with _lib.on_device(self.device):
            _lib.check(_lib.lib().ccst_set_encoder_weights(self._h, _ptr_array(ws), _ptr_array(bs)))
        self._enc_version = ver

    def set_decoder(self, decoder: nn.Module):
        ver = _version(decoder)
        if ver == self._dec_version:
            return
        ws, bs = _host_arrays(_conv_params(decoder, 9, "decoder"), _DEC_SHAPES, "decoder")
        with _lib.on_device(self.device):
            _lib.check(_lib.lib().ccst_set_decoder_weights(self._h, _ptr_array(ws), _ptr_array(bs)))
        self._dec_version = ver

    # -- helpers -----------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _img(self, t, name):
        t = F_._prep(t, name)
        if t.device != self.device:
            raise RuntimeError(f"{name} is on {t.device}, engine on {self.device}")
        if t.dim() != 4 or t.shape[1] != 3:
            raise RuntimeError(f"{name} must be [N,3,H,W], got {tuple(t.shape)}")
        return t

    # -- entry points --------------------------------------------------------
    def encode(self, images, precision=DEFAULT_PRECISION):
        """vgg(images) -> relu4_1 [N,512,h,w] fp32."""
        x = self._img(images, "images")
        n, _, h, w = x.shape
        fh, fw = _lib.feature_hw(h, w)
        out = torch.empty((n, 512, fh, fw), dtype=torch.float32, device=self.device)
        with _lib.on_device(self.device):
            _lib.check(_lib.lib().ccst_encoder_fwd(self._h, x.data_ptr(), n, h, w, out.data_ptr(),
                                                   PRECISIONS[precision], self._stream()))
        return out

    def encode_levels(self, images, precision=DEFAULT_PRECISION, eps=F_.EPS, want_feat=True):
        """Net.encode_with_intermediate + calc_mean_std per level (net.py:112-136), forward only:
        returns (relu4_1 [N,512,h,w] or None, [(mean, std)] for relu1_1, relu2_1, relu3_1, relu4_1, each
        [N,C_l,1,1]).  The intermediate maps never leave the arena; only their statistics do."""
        x = self._img(images, "images")
        n, _, h, w = x.shape
        fh, fw = _lib.feature_hw(h, w)
        feat = torch.empty((n, 512, fh, fw), dtype=torch.float32, device=self.device) if want_feat else None
        means = [torch.empty((n, c, 1, 1), dtype=torch.float32, device=self.device) for c in (64, 128, 256, 512)]
        stds = [torch.empty_like(m) for m in means]
        with _lib.on_device(self.device):
            _lib.check(_lib.lib().ccst_encoder_levels(
                self._h, x.data_ptr(), n, h, w, feat.data_ptr() if want_feat else None, _ptr_array(means),
                _ptr_array(stds), float(eps), PRECISIONS[precision], self._stream()))
        return feat, list(zip(means, stds))

    def decode(self, feat, precision=DEFAULT_PRECISION):
        """decoder(feat): [N,512,h,w] -> [N,3,8h,8w] fp32."""
        x = F_._prep(feat, "feat")
        if x.dim() != 4 or x.shape[1] != 512:
            raise RuntimeError(f"feat must be [N,512,h,w], got {tuple(x.shape)}")
        n, _, fh, fw = x.shape
        out = torch.empty((n, 3, 8 * fh, 8 * fw), dtype=torch.float32, device=self.device)
        with _lib.on_device(self.device):
            _lib.check(_lib.lib().ccst_decoder_fwd(self._h, x.data_ptr(), n, fh, fw, out.data_ptr(),
                                                   PRECISIONS[precision], self._stream()))
        return out

    def accumulate(self, images, state: "F_.WelfordState", precision=DEFAULT_PRECISION):
        """One iteration of the overall-statistics loop
        (mean_std_computation_effcientMem.py:121-131): encode + fold relu4_1 into `state`."""
        x = self._img(images, "images")
        n, _, h, w = x.shape
        if state.C != 512 or state.device != self.device:
            raise RuntimeError("state must be a 512-channel WelfordState on the engine's device")
        with _lib.on_device(self.device):
            _lib.check(_lib.lib().ccst_encoder_accumulate(self._h, x.data_ptr(), n, h, w,
                                                          state.buf.data_ptr(), PRECISIONS[precision],
                                                          self._stream()))
        return state

    def accumulate_u8(self, images_u8, state: "F_.WelfordState", precision=DEFAULT_PRECISION):
        """`accumulate` on the loader's uint8 HWC batch [N,H,W,3] (ToTensor on the GPU)."""
        x = images_u8
        if not isinstance(x, torch.Tensor) or not x.is_cuda or x.dtype != torch.uint8 or x.dim() != 4 \
                or x.shape[3] != 3:
            raise RuntimeError("images_u8 must be a CUDA uint8 [N,H,W,3] tensor: ccst_b200 has no CPU fallback")
        if x.device != self.device:
            raise RuntimeError(f"images_u8 is on {x.device}, engine on {self.device}")
        x = x.contiguous()
        n, h, w, _ = x.shape
        if state.C != 512 or state.device != self.device:
            raise RuntimeError("state must be a 512-channel WelfordState on the engine's device")
        with _lib.on_device(self.device):
            _lib.check(_lib.lib().ccst_encoder_accumulate_u8(self._h, x.data_ptr(), n, h, w,
                                                             state.buf.data_ptr(), PRECISIONS[precision],
                                                             self._stream()))
        return state

    def transfer(self, content, style_stat, alpha=1.0, precision=DEFAULT_PRECISION, out=None):
        """Fused encoder -> AdaIN(+alpha) -> decoder; activations never leave the arena."""
        assert (0.0 <= alpha <= 1.0)
        x = self._img(content, "content")
        n, _, h, w = x.shape
        mu, sg, stride = F_._style_stat_args(style_stat, n, 512, self.device)
        fh, fw = _lib.feature_hw(h, w)
        if out is None:
            out = torch.empty((n, 3, 8 * fh, 8 * fw), dtype=torch.float32, device=self.device)
        with _lib.on_device(self.device):
            _lib.check(_lib.lib().ccst_style_transfer(
                self._h, x.data_ptr(), n, h, w, mu.data_ptr(), sg.data_ptr(), stride, float(alpha),
                out.data_ptr(), PRECISIONS[precision], self._stream()))
        return out

    def transfer_u8(self, content_u8, style_stat, alpha=1.0, precision=DEFAULT_PRECISION, out=None):
        """`transfer` with the batch loop's image I/O fused around it: `content_u8` is the
        loader's uint8 HWC batch [N,H,W,3] before `ToTensor` (cjm_util/data_helper.py:45), the result
        the uint8 HWC batch [N,8h,8w,3] that `save_image` encodes
        (CCST_OverallStyleTransfer.py:167) -- a quarter of the fp32 bytes each way over PCIe."""
        assert (0.0 <= alpha <= 1.0)
        x = content_u8
        if not isinstance(x, torch.Tensor) or not x.is_cuda or x.dtype != torch.uint8:
            raise RuntimeError("content_u8 must be a CUDA uint8 tensor: ccst_b200 has no CPU fallback")
        if x.device != self.device:
            raise RuntimeError(f"content_u8 is on {x.device}, engine on {self.device}")
        if x.dim() != 4 or x.shape[3] != 3:
            raise RuntimeError(f"content_u8 must be [N,H,W,3], got {tuple(x.shape)}")
        x = x.contiguous()
        n, h, w, _ = x.shape
        mu, sg, stride = F_._style_stat_args(style_stat, n, 512, self.device)
        fh, fw = _lib.feature_hw(h, w)
        if out is None:
            out = torch.empty((n, 8 * fh, 8 * fw, 3), dtype=torch.uint8, device=self.device)
        with _lib.on_device(self.device):
            _lib.check(_lib.lib().ccst_style_transfer_u8(
                self._h, x.data_ptr(), n, h, w, mu.data_ptr(), sg.data_ptr(), stride, float(alpha),
                out.data_ptr(), PRECISIONS[precision], self._stream()))
        return out

    # -- f16 range guard -------------------------------------------------------
    def saturation_count(self, reset: bool = False) -> int:
        """Number of epilogue threads whose f16 stores hit the +-65504 clamp since the last reset
        (synchronises the current stream).  Non-zero means the weights drive activations outside the
        f16 range: use precision="bf16" (or "fp32") for them."""
        if not hasattr(self, "_sat_host"):
            self._sat_host = torch.zeros((1,), dtype=torch.int32).pin_memory()
        st = torch.cuda.current_stream(self.device)
        with _lib.on_device(self.device):
            _lib.check(_lib.lib().ccst_saturation_snapshot(self._h, self._sat_host.data_ptr(), st.cuda_stream))
            if reset:
                _lib.check(_lib.lib().ccst_saturation_reset(self._h, st.cuda_stream))
        st.synchronize()
        return int(self._sat_host.item()) & 0xFFFFFFFF

    def check_saturation(self):
        """Raise if any f16 store saturated since the last check (and reset the counter)."""
        n = self.saturation_count(reset=True)
        if n:
            raise F16SaturationError(
                f"{n} epilogue threads stored activations clamped to +-65504: these weights exceed the f16 "
                "range of precision='fp16'; re-run with precision='bf16' (same speed, wider range) or 'fp32'")

    def saturation_snapshot_async(self, host_i32: torch.Tensor, stream: "torch.cuda.Stream"):
        """Enqueue a copy of the counter into pinned `host_i32[0]` on `stream` (valid after it syncs)."""
        with _lib.on_device(self.device):
            _lib.check(_lib.lib().ccst_saturation_snapshot(self._h, host_i32.data_ptr(), stream.cuda_stream))

    def set_fusion(self, mask: int = _lib.FUSE_ALL):
        """Tests only: switch individual kernel fusions off (see CCST_FUSE_* in the header)."""
        _lib.check(_lib.lib().ccst_set_fusion(self._h, int(mask)))

    # -- profiling -----------------------------------------------------------
    def profile(self, on: bool):
        _lib.check(_lib.lib().ccst_profile_enable(self._h, 1 if on else 0))

    def profile_read(self, max_entries=48):
        ms = (C.c_float * max_entries)()
        fl = (C.c_double * max_entries)()
        by = (C.c_double * max_entries)()
        kd = (C.c_int * max_entries)()
        n = _lib.check(_lib.lib().ccst_profile_read(self._h, max_entries, ms, fl, by, kd))
        return [dict(ms=ms[i], flops=fl[i], bytes=by[i], kind=kd[i]) for i in range(n)]

    def debug_conv3x3(self, x_nhwc, weight, bias, relu=True, mode=0, precision=DEFAULT_PRECISION):
        """Single 3x3 reflect-pad conv through the selected engine (tests only)."""
        x = F_._prep(x_nhwc, "x_nhwc")
        n, h, w, cin = x.shape
        cout = weight.shape[0]
        wh = weight.detach().to("cpu", torch.float32).contiguous()
        bh = bias.detach().to("cpu", torch.float32).contiguous()
        if mode == 0:
            shape = (n, h, w, cout)
        elif mode in (1, 4):  # 1: upsample after the conv; 4: upsample before it (fused, tcgen05 only)
            shape = (n, 2 * h, 2 * w, cout)
        elif mode == 2:
            shape = (n, (h + 1) // 2, (w + 1) // 2, cout)
        else:
            shape = (n, cout, h, w)
        out = torch.empty(shape, dtype=torch.float32, device=self.device)
        with _lib.on_device(self.device):
            _lib.check(_lib.lib().ccst_debug_conv3x3(
                self._h, x.data_ptr(), n, h, w, cin, cout, wh.data_ptr(), bh.data_ptr(),
                1 if relu else 0, mode, out.data_ptr(), PRECISIONS[precision], self._stream()))
        return out


def to_tensor_u8(images_u8: torch.Tensor) -> torch.Tensor:
    """`transforms.ToTensor()` of a uint8 HWC batch on the device (cjm_util/data_helper.py:45):
    [N,H,W,C] uint8 -> [N,C,H,W] fp32 = float(u) / 255."""
    x = images_u8
    if not isinstance(x, torch.Tensor) or not x.is_cuda or x.dtype != torch.uint8 or x.dim() != 4:
        raise RuntimeError("images_u8 must be a CUDA uint8 [N,H,W,C] tensor: ccst_b200 has no CPU fallback")
    x = x.contiguous()
    n, h, w, c = x.shape
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    with _lib.on_device(x.device):
        _lib.check(_lib.lib().ccst_u8_to_tensor(x.data_ptr(), n, c, h, w, out.data_ptr(),
                                                torch.cuda.current_stream(x.device).cuda_stream))
    return out


def save_image_quantize(images: torch.Tensor) -> torch.Tensor:
    """The tensor -> uint8 step of `save_image` (CCST_OverallStyleTransfer.py:167; torchvision
    utils.save_image: mul(255).add_(0.5).clamp_(0,255).permute(1,2,0).to(uint8)) for a batch:
    [N,C,H,W] fp32 -> [N,H,W,C] uint8."""
    x = F_._prep(images, "images")
    if x.dim() != 4:
        raise RuntimeError(f"images must be [N,C,H,W], got {tuple(x.shape)}")
    n, c, h, w = x.shape
    out = torch.empty((n, h, w, c), dtype=torch.uint8, device=x.device)
    with _lib.on_device(x.device):
        _lib.check(_lib.lib().ccst_quantize_u8(x.data_ptr(), n, c, h, w, out.data_ptr(),
                                               torch.cuda.current_stream(x.device).cuda_stream))
    return out


def resize_input_u8(images_u8: torch.Tensor, size) -> torch.Tensor:
    """The loader's `transforms.Resize((S, S))` (cjm_util/data_helper.py:45-49) on a uint8 HWC batch
    [N,H,W,C] on the device: Pillow's 8-bit bilinear resample, bit-exact -> [N,S,S,C] uint8.  `size` is an
    int (square, as the reference's `(image_size, image_size)`) or (h, w)."""
    x = images_u8
    if not isinstance(x, torch.Tensor) or not x.is_cuda or x.dtype != torch.uint8 or x.dim() != 4:
        raise RuntimeError("images_u8 must be a CUDA uint8 [N,H,W,C] tensor: ccst_b200 has no CPU fallback")
    x = x.contiguous()
    n, h, w, c = x.shape
    oh, ow = (int(size), int(size)) if not isinstance(size, (tuple, list)) else (int(size[0]), int(size[1]))
    if (oh, ow) == (h, w):
        return x  # PIL returns a copy of the image; nothing to compute
    out = torch.empty((n, oh, ow, c), dtype=torch.uint8, device=x.device)
    nbytes = _lib.lib().ccst_resize_pil_scratch_bytes(n, h, w, c, oh, ow)
    scratch = torch.empty((nbytes,), dtype=torch.uint8, device=x.device)
    with _lib.on_device(x.device):
        _lib.check(_lib.lib().ccst_resize_pil_bilinear_u8(x.data_ptr(), n, h, w, c, oh, ow, out.data_ptr(),
                                                          scratch.data_ptr(),
                                                          torch.cuda.current_stream(x.device).cuda_stream))
    return out


def resized_output_size(h: int, w: int, size):
    """torchvision `_compute_resized_output_size` (no max_size): an int matches the SMALLER edge."""
    if isinstance(size, (tuple, list)):
        if len(size) == 2:
            return int(size[0]), int(size[1])
        size = size[0]
    size = int(size)
    if h <= w:
        return size, int(size * w / h)
    return int(size * h / w), size


def resize(images: torch.Tensor, size) -> torch.Tensor:
    """`transforms.Resize(size)(images)` for a float tensor [N,C,H,W]
    (CCST_OverallStyleTransfer.py:134-135,154-155): anti-aliased bilinear, on the device."""
    x = F_._prep(images, "images")
    if x.dim() != 4:
        raise RuntimeError(f"images must be [N,C,H,W], got {tuple(x.shape)}")
    n, c, h, w = x.shape
    oh, ow = resized_output_size(h, w, size)
    if (oh, ow) == (h, w):
        return x  # torchvision returns the image itself
    out = torch.empty((n, c, oh, ow), dtype=torch.float32, device=x.device)
    with _lib.on_device(x.device):
        _lib.check(_lib.lib().ccst_resize_bilinear_aa_f32(x.data_ptr(), n * c, h, w, oh, ow, out.data_ptr(),
                                                          torch.cuda.current_stream(x.device).cuda_stream))
    return out


_ENGINES = {}


def engine_for(vgg: nn.Module, decoder: nn.Module, device) -> Engine:
    """Engine cached per (vgg object, decoder object, device); weights are re-packed when the
    modules' parameters change (load_state_dict, .to(), in-place edits)."""
    device = torch.device(device)
    key = (id(vgg), id(decoder), device.index if device.index is not None else torch.cuda.current_device())
    eng = _ENGINES.get(key)
    if eng is None or eng._vgg_ref() is not vgg or eng._dec_ref() is not decoder:
        eng = Engine(device=device)
        eng._vgg_ref = weakref.ref(vgg)
        eng._dec_ref = weakref.ref(decoder)
        _ENGINES[key] = eng
    eng.set_encoder(vgg)
    eng.set_decoder(decoder)
    return eng


def style_transfer_u8(vgg, decoder, content_u8, style_stat, alpha=1.0, *, precision=None):
    """`style_transfer` on the loader's uint8 HWC batch, returning the uint8 HWC batch `save_image`
    would encode (see Engine.transfer_u8)."""
    assert (0.0 <= alpha <= 1.0)
    if not isinstance(content_u8, torch.Tensor) or not content_u8.is_cuda:
        raise RuntimeError("content_u8 must be a CUDA tensor: ccst_b200 has no CPU fallback")
    eng = engine_for(vgg, decoder, content_u8.device)
    return eng.transfer_u8(content_u8, style_stat, alpha, precision or DEFAULT_PRECISION)


def style_transfer(vgg, decoder, content, style, alpha=1.0, interpolation_weights=None, *,
                   precision=None):
    """Reference signature (CCST_OverallStyleTransfer.py:32).  `precision` is a keyword-only
    extension: "fp16" (tcgen05 convs, f16 operands, default), "bf16" (same kernels, bf16 operands) or
    "fp32" (FFMA validation mode)."""
    assert (0.0 <= alpha <= 1.0)
    precision = precision or DEFAULT_PRECISION
    if not isinstance(content, torch.Tensor) or not content.is_cuda:
        raise RuntimeError("content must be a CUDA tensor: ccst_b200 has no CPU fallback")
    eng = engine_for(vgg, decoder, content.device)
    is_image_style = isinstance(style, torch.Tensor) and style.dim() == 4 and style.shape[1] == 3
    if is_image_style:
        # upstream AdaIN form: per-sample statistics of the encoded style images
        assert (content.size()[0] == style.size()[0])
        style_f = eng.encode(style.to(content.device), precision)
        style = F_.calc_mean_std(style_f)
    if interpolation_weights:
        # CCST_OverallStyleTransfer.py:36-42 (never taken by the CCST drivers; kept for the signature)
        content_f = eng.encode(content, precision)
        base = F_.adaIN_StyleStat_ContentFeat(content_f, style)
        feat = torch.zeros_like(content_f[0:1])
        for i, w in enumerate(interpolation_weights):
            feat = feat + w * base[i:i + 1]
        feat = feat * alpha + content_f[0:1] * (1 - alpha)
        return eng.decode(feat, precision)
    return eng.transfer(content, style, alpha, precision)
