"""VGG-19 encoder and mirrored decoder definitions (architecture only).

Drop-in for the module-level ``vgg`` / ``decoder`` objects of the reference
(`style_transfer/AdaIN/net.py:6-36` decoder, `:38-92` vgg).  The objects are
ordinary ``nn.Sequential`` containers whose ``state_dict()`` keys are identical
to the reference's (integer positions ``0,2,5,9,...``), so
``load_state_dict(torch.load('vgg_normalised.pth'))`` works unchanged and the
callers' ``nn.Sequential(*list(vgg.children())[:31])`` truncation to relu4_1
(`CCST_OverallStyleTransfer.py:124`) keeps working.

These modules are *weight containers* for the B200 path: `ccst_b200.transfer`
reads their conv weights once, packs them for the CUDA kernels and never calls
``forward`` on them.  (Calling ``forward`` runs stock PyTorch and is what the
oracle does on CPU.)

The stacks are generated from a compact channel plan instead of being listed
layer by layer.
"""
from __future__ import annotations

import torch.nn as nn

# (kind, cin, cout): 'c' = reflect-pad + 3x3 conv + ReLU, 'p' = 2x2 ceil-mode
# max-pool, 'u' = nearest x2 upsample, 'l' = reflect-pad + 3x3 conv (linear).
_VGG_STAGES = (
    (64, 2),    # relu1_1, relu1_2
    (128, 2),   # relu2_1, relu2_2
    (256, 4),   # relu3_1 .. relu3_4
    (512, 4),   # relu4_1 .. relu4_4
    (512, 4),   # relu5_1 .. relu5_4
)
_DEC_PLAN = (
    ("c", 512, 256), ("u",),
    ("c", 256, 256), ("c", 256, 256), ("c", 256, 256), ("c", 256, 128), ("u",),
    ("c", 128, 128), ("c", 128, 64), ("u",),
    ("c", 64, 64), ("l", 64, 3),
)

#: number of leading vgg children that end at relu4_1 (reference slices [:31])
RELU4_1_CHILDREN = 31


def vgg_plan():
    """Layer plan of the full encoder as a list of tuples (see _DEC_PLAN)."""
    plan = [("1x1", 3, 3)]
    cin = 3
    for si, (width, reps) in enumerate(_VGG_STAGES):
        if si > 0:
            plan.append(("p",))
        for _ in range(reps):
            plan.append(("c", cin, width))
            cin = width
    return plan


def decoder_plan():
    return list(_DEC_PLAN)


def _materialise(plan):
    mods = []
    for item in plan:
        kind = item[0]
        if kind == "1x1":
            mods.append(nn.Conv2d(item[1], item[2], (1, 1)))
        elif kind in ("c", "l"):
            mods.append(nn.ReflectionPad2d((1, 1, 1, 1)))
            mods.append(nn.Conv2d(item[1], item[2], (3, 3)))
            if kind == "c":
                mods.append(nn.ReLU())
        elif kind == "p":
            mods.append(nn.MaxPool2d((2, 2), (2, 2), (0, 0), ceil_mode=True))
        elif kind == "u":
            mods.append(nn.Upsample(scale_factor=2, mode="nearest"))
        else:  # pragma: no cover
            raise ValueError(kind)
    return nn.Sequential(*mods)


def make_vgg() -> nn.Sequential:
    """Fresh full VGG-19 stack (53 children, through relu5_4)."""
    return _materialise(vgg_plan())


def make_decoder() -> nn.Sequential:
    """Fresh decoder stack (29 children)."""
    return _materialise(decoder_plan())


def truncate_relu4_1(vgg: nn.Sequential) -> nn.Sequential:
    """Same as the callers' ``nn.Sequential(*list(vgg.children())[:31])``."""
    return nn.Sequential(*list(vgg.children())[:RELU4_1_CHILDREN])


# module-level instances, like the reference (weights are whatever PyTorch's
# default init gives until load_state_dict / ccst_b200.synth.init_* is applied)
decoder = make_decoder()
vgg = make_vgg()


class Net(nn.Module):
    """Forward-only drop-in for the reference's training wrapper `Net` (net.py:95-152): `forward(content,
    style, alpha)` returns (loss_c, loss_s) -- the content loss MSE(relu4_1(g_t), t) and the style loss
    sum over relu1_1..relu4_1 of MSE(mean) + MSE(std) -- computed by the B200 engine: three encoder passes
    whose intermediate maps stay in the arena (only their calc_mean_std statistics leave it), the fused
    AdaIN + alpha blend, the decoder, and the deterministic MSE kernel.  No autograd (inference / evaluation
    of the losses only): backward is out of scope."""

    def __init__(self, encoder, decoder, precision="fp32"):
        super().__init__()
        enc_layers = list(encoder.children())
        self.encoder = nn.Sequential(*enc_layers[:RELU4_1_CHILDREN])  # input -> relu4_1 (net.py:98-102)
        self.decoder = decoder
        self.precision = precision

    def _engine(self, device):
        from .transfer import engine_for

        return engine_for(self.encoder, self.decoder, device)

    def encode_with_intermediate_stats(self, images):
        return self._engine(images.device).encode_levels(images, self.precision)

    def encode(self, images):
        return self._engine(images.device).encode(images, self.precision)

    @staticmethod
    def calc_content_loss(input, target):
        from . import function as F_

        assert (input.size() == target.size())
        return F_.mse_loss(input, target)

    @staticmethod
    def calc_style_loss_from_stats(input_stats, target_stats):
        from . import function as F_

        return F_.mse_loss(input_stats[0], target_stats[0]) + F_.mse_loss(input_stats[1], target_stats[1])

    def forward(self, content, style, alpha=1.0):
        from . import function as F_

        assert 0 <= alpha <= 1
        eng = self._engine(content.device)
        style_feat, style_stats = eng.encode_levels(style, self.precision, want_feat=False)
        content_feat = eng.encode(content, self.precision)
        assert (content_feat.size()[:2] == style_stats[3][0].size()[:2])  # adain's assert (function.py:17)
        t = F_.adain_blend(content_feat, style_stats[3], alpha)  # adain + alpha blend (net.py:141-142)
        g_t = eng.decode(t, self.precision)
        # calc_style_loss asserts input.size() == target.size() on every level (net.py:131): same image size
        assert (g_t.size() == style.size())
        g_feat, g_stats = eng.encode_levels(g_t, self.precision)
        loss_c = self.calc_content_loss(g_feat, t)
        loss_s = self.calc_style_loss_from_stats(g_stats[0], style_stats[0])
        for i in range(1, 4):
            loss_s = loss_s + self.calc_style_loss_from_stats(g_stats[i], style_stats[i])
        return loss_c, loss_s
