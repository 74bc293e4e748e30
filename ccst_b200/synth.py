"""Deterministic synthetic weights and images.

`decoder.pth`, `vgg_normalised.pth` and the datasets are external downloads in
the reference (`README.md:21`) and are not available offline, so every parity
test and benchmark uses random-init weights of the named architectures and
synthetic images.  PyTorch's *default* conv init collapses relu4_1 to
std = sqrt(eps), which would make every tolerance trivially
true; the init below keeps activations O(1) through all 19 convolutions:

* hidden convs: He-normal (std = sqrt(2 / fan_in)), bias ~ N(0, 0.05)
* the 1x1 colour conv: a fixed well-conditioned mixing matrix + offset
* last decoder conv: scaled so the output image is ~ 0.5 +- 0.2 (inside [0,1]
  for most pixels), so the "max-abs in [0,1]" tolerance of BASELINE.json means
  what it says.

Everything is generated on the CPU generator, which is bit-reproducible for a
given torch version, so the oracle (here) and the CUDA path (on the GPU box)
see identical tensors.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


def _convs(seq: nn.Sequential):
    return [m for m in seq if isinstance(m, nn.Conv2d)]


@torch.no_grad()
def init_vgg_(vgg: nn.Sequential, seed: int = 0) -> nn.Sequential:
    g = torch.Generator().manual_seed(1000 + seed)
    convs = _convs(vgg)
    first = convs[0]
    assert first.kernel_size == (1, 1)
    mix = torch.tensor([[0.9, 0.3, -0.2], [-0.25, 1.0, 0.35], [0.3, -0.2, 0.95]])
    first.weight.copy_((2.0 * mix).view(3, 3, 1, 1))
    first.bias.copy_(torch.tensor([-1.0, -1.1, -0.9]))
    for conv in convs[1:]:
        fan_in = conv.in_channels * 9
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * math.sqrt(2.0 / fan_in))
        conv.bias.copy_(torch.randn(conv.bias.shape, generator=g) * 0.05)
    return vgg


@torch.no_grad()
def init_decoder_(decoder: nn.Sequential, seed: int = 0) -> nn.Sequential:
    g = torch.Generator().manual_seed(2000 + seed)
    convs = _convs(decoder)
    for i, conv in enumerate(convs):
        fan_in = conv.in_channels * 9
        last = i == len(convs) - 1
        gain = 0.35 if last else math.sqrt(2.0)
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * gain / math.sqrt(fan_in))
        if last:
            conv.bias.fill_(0.5)
        else:
            conv.bias.copy_(torch.randn(conv.bias.shape, generator=g) * 0.05)
    return decoder


def make_models(seed: int = 0):
    """(vgg truncated to relu4_1, decoder), eval mode, synthetic weights."""
    from . import net

    vgg = init_vgg_(net.make_vgg(), seed)
    dec = init_decoder_(net.make_decoder(), seed)
    vgg = net.truncate_relu4_1(vgg).eval()
    dec = dec.eval()
    for p in list(vgg.parameters()) + list(dec.parameters()):
        p.requires_grad_(False)
    return vgg, dec


def images(n: int, h: int, w: int, seed: int) -> torch.Tensor:
    """Synthetic image batch in [0,1], NCHW fp32, smooth + noise so that the
    feature statistics differ between images and between channels."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand((n, 3, h, w), generator=g)
    # low-frequency component (bilinear up-sampling of a coarse random field)
    ch, cw = max(2, h // 16), max(2, w // 16)
    coarse = torch.rand((n, 3, ch, cw), generator=g)
    smooth = torch.nn.functional.interpolate(coarse, size=(h, w), mode="bilinear", align_corners=False)
    return (0.5 * base + 0.5 * smooth).clamp_(0.0, 1.0).contiguous()


def features(shape, seed: int, mean_spread: float = 1.0, relu: bool = True) -> torch.Tensor:
    """Synthetic relu4_1-like feature map: per-channel offset/scale + ReLU."""
    g = torch.Generator().manual_seed(seed)
    n, c = shape[:2]
    x = torch.randn(shape, generator=g)
    scale = torch.rand((1, c, 1, 1), generator=g) * 1.5 + 0.05
    shift = torch.randn((1, c, 1, 1), generator=g) * mean_spread
    x = x * scale + shift
    if relu:
        x = torch.relu(x)
    return x.contiguous()
