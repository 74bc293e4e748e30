"""Per-client overall-style statistics (mean_std_computation_effcientMem.py:117-146).

The reference streams one client's images through vgg[:31], keeps running
fp32 sums of x and x^2 per channel and finalises mean / biased std.  Here the
loop body is one library call per batch (encoder + Welford fold, features never
leave the GPU arena), the running state is a device-resident fp64
{count, mean, M2} triple, and -- when the client's images are sharded over
several GPUs -- the per-rank partials are merged with ONE all-reduce(sum) of
1+2C fp64 moments.

File format: `np.save(path, [mean, std])` -> float32 array (2,1,C,1,1), byte
compatible with what CCST_OverallStyleTransfer.py:140-144 loads.
"""
from __future__ import annotations

import numpy as np
import torch

from . import function as F_
from .transfer import Engine

# The statistics drivers default to the f16x3 engine: the tensor-core encoder with split f16 operands and
# promoted fp32 accumulation (csrc/conv_x3.cuh).  Its relu4_1 features -- and hence the .npy a user writes for
# a client -- meet the 1e-5 relative bar of BASELINE.json against the reference (measured 2e-6), like the fp32
# CUDA-core engine ("fp32", 1e-7, ~8x slower).  The plain 16-bit encoder ("fp16"/"bf16", ~4x faster again)
# lands at ~1e-3 relative (tests/test_gpu_net.py prints the measured figures) and must be asked for explicitly.
STATS_PRECISION = "fp16x3"


def shard_range(total: int, rank: int, world: int):
    """Contiguous [begin, end) slice of `total` items owned by `rank` (sizes differ by <= 1)."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allreduce_moments(moments: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the {n, n*mean, M2+n*mean^2} vectors of all ranks (NCCL on GPUs, gloo on CPU tests)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(moments, op=dist.ReduceOp.SUM, group=group)
    return moments


def finalize_moments(moments: torch.Tensor, eps: float = F_.EPS):
    """Host-side (device-agnostic, fp64) finalisation of a summed moments vector:
    mean = S1/n, var = S2/n - mean^2 (biased, :135-137), std = sqrt(var + eps)."""
    c = (moments.numel() - 1) // 2
    n = moments[0]
    mean = moments[1:1 + c] / n
    var = (moments[1 + c:] / n - mean * mean).clamp_min(0.0)
    std = torch.sqrt(var + eps)
    return (mean.to(torch.float32).view(1, c, 1, 1), std.to(torch.float32).view(1, c, 1, 1))


class OverallStyleAccumulator:
    """Running style statistics of one client on one GPU."""

    def __init__(self, engine: Engine, precision: str = STATS_PRECISION, channels: int = 512):
        self.engine = engine
        self.precision = precision
        self.state = F_.WelfordState(channels, engine.device)
        self.img_count = 0          # images folded in by THIS rank
        self.global_img_count = 0   # images of all ranks, set by finalize()

    def add_images(self, images: torch.Tensor):
        """`feat = vgg(data); calc_sum(feat); all_* += ...` (:121-131)."""
        if images.dtype == torch.uint8:  # the loader's HWC images before ToTensor
            self.engine.accumulate_u8(images, self.state, self.precision)
        else:
            self.engine.accumulate(images, self.state, self.precision)
        self.img_count += int(images.shape[0])
        return self

    def add_features(self, feat: torch.Tensor):
        self.state.add_features(feat)
        self.img_count += int(feat.shape[0])
        return self

    def finalize(self, group=None, eps: float = F_.EPS):
        """(mean, std) each [1,C,1,1] fp32 over all ranks of `group` (when torch.distributed is
        initialised): ONE all-reduce(sum) of the 1+2C fp64 moments {n, n*mean, M2+n*mean^2} plus the image
        counter.  The per-rank partial in `self.state` is left untouched, so finalize() may be called again
        (after more add_images(), or twice) without counting any rank's data twice; `self.img_count` stays
        this rank's count and `self.global_img_count` holds the merged one."""
        import torch.distributed as dist

        self.global_img_count = self.img_count
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            payload = torch.cat([self.state.moments(),
                                 torch.tensor([float(self.img_count)], dtype=torch.float64, device=self.state.device)])
            allreduce_moments(payload, group)
            self.global_img_count = int(round(payload[-1].item()))
            merged = F_.WelfordState(self.state.C, self.state.device).load_moments(payload[:-1])
            return merged.finalize(eps)
        return self.state.finalize(eps)


def save_style_stats(path: str, mean: torch.Tensor, std: torch.Tensor):
    """np.save(path, [mean, std]) as in mean_std_computation_effcientMem.py:146."""
    np.save(path, np.asarray([mean.detach().cpu().numpy(), std.detach().cpu().numpy()]))


def load_style_stats(path: str, device):
    """CCST_OverallStyleTransfer.py:140-144: np.load -> two [1,C,1,1] tensors on `device`."""
    arr = np.load(path)
    return [torch.Tensor(stat).to(device) for stat in arr]
