"""Batch loops of the three CCST scripts, restructured for a GPU that is faster than its PCIe link.

Reference loops (all in `style_transfer/AdaIN/`):

* `CCST_OverallStyleTransfer.py:138-167` -- for each style domain, for each content batch:
  `data.to(device)` -> `style_transfer` -> `output.cpu()` -> save.
* `CCST_SingleStyleTransfer.py:176-223`  -- same, but one random style image per *batch* whose
  relu4_1 statistics are computed on the fly (`:195-205`).
* `mean_std_computation_effcientMem.py:117-137` -- for each batch: `vgg(data)`, `calc_sum`, accumulate.

The reference does H2D, compute and D2H strictly one after the other on the default stream.  On a
B200 the encoder/decoder take ~8 ms per 32-image batch while moving 100 MB in and 100 MB out over
PCIe takes ~4 ms, so `TransferPipeline` double-buffers: the upload of batch i+1 and the download of
batch i-1 run on their own streams underneath the compute of batch i.  Results are identical to
calling `style_transfer` per batch.

Image decoding / resizing / `save_image` are out of scope; batches are tensors.
"""
from __future__ import annotations

import random
from typing import Callable, Iterable, Iterator, List, Optional, Sequence, Tuple

import torch

from . import function as F_
from .overall import STATS_PRECISION, OverallStyleAccumulator
from .transfer import DEFAULT_PRECISION, Engine, F16SaturationError


class TransferPipeline:
    """Pinned-host batches in, pinned-host stylised batches out, three streams, two device slots and
    `slots + 1` pinned result buffers (so a yielded result stays valid while the next one is produced).

    With precision "fp16" every result also carries a snapshot of the engine's f16-saturation counter
    (copied on the download stream right after the images); a non-zero count raises F16SaturationError
    when the batch is handed out instead of returning a silently clamped image."""

    def __init__(self, engine: Engine, precision: str = DEFAULT_PRECISION, slots: int = 2, u8: bool = False):
        """u8 = True: batches are the loader's uint8 HWC images [N,H,W,3] (before ToTensor) and the
        results the uint8 HWC images `save_image` encodes (Engine.transfer_u8): 4x fewer PCIe bytes."""
        self.engine = engine
        self.precision = precision
        self.slots = slots
        self.u8 = u8
        self.dtype = torch.uint8 if u8 else torch.float32
        dev = engine.device
        self.s_in = torch.cuda.Stream(dev)
        self.s_cmp = torch.cuda.Stream(dev)
        self.s_out = torch.cuda.Stream(dev)
        self._in: List[Optional[torch.Tensor]] = [None] * slots
        self._out: List[Optional[torch.Tensor]] = [None] * slots
        self._host: List[Optional[torch.Tensor]] = [None] * (slots + 1)
        self._sat = torch.zeros((slots + 1,), dtype=torch.int32).pin_memory()
        self._ev_in = [torch.cuda.Event() for _ in range(slots)]
        self._ev_cmp = [torch.cuda.Event() for _ in range(slots)]
        self._ev_out = [torch.cuda.Event() for _ in range(slots)]
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _buffers(self, slot: int, hslot: int, shape, out_shape):
        dev = self.engine.device
        if self._in[slot] is None or tuple(self._in[slot].shape) != tuple(shape):
            self._in[slot] = torch.empty(shape, dtype=self.dtype, device=dev)
        if self._out[slot] is None or tuple(self._out[slot].shape) != tuple(out_shape):
            self._out[slot] = torch.empty(out_shape, dtype=self.dtype, device=dev)
        if self._host[hslot] is None or tuple(self._host[hslot].shape) != tuple(out_shape):
            self._host[hslot] = torch.empty(out_shape, dtype=self.dtype).pin_memory()
        return self._in[slot], self._out[slot], self._host[hslot]

    def _hand_out(self, j: int, slot: int, hslot: int):
        self._ev_out[slot].synchronize()
        if self.precision == "fp16" and int(self._sat[hslot].item()) != 0:
            self.engine.saturation_count(reset=True)
            raise F16SaturationError(
                f"batch {j}: activations left the f16 range (stores clamped to +-65504); re-run with "
                "precision='bf16' (same speed, wider range) or 'fp32'")
        return j, self._host[hslot]

    @classmethod
    def for_engine(cls, engine: Engine, precision: str = DEFAULT_PRECISION, u8: bool = False) -> "TransferPipeline":
        """The engine's pipeline for (precision, u8), kept across calls: its device slots and pinned host buffers
        (3 x 100 MB for fp32 batches of 32 @512^2 -- cudaHostAlloc costs ~50 ms each) are allocated once per
        engine instead of once per loop.  A pipeline whose previous loop has not been exhausted is not reused."""
        cache = engine.__dict__.setdefault("_pipelines", {})
        pipe = cache.get((precision, u8))
        if pipe is None or pipe._busy:
            pipe = cls(engine, precision, u8=u8)
            cache[(precision, u8)] = pipe
        return pipe

    _busy = False

    def run(self, host_batches: Iterable[torch.Tensor],
            style_for_batch: Callable[[int, torch.Tensor], Sequence[torch.Tensor]],
            alpha: float = 1.0) -> Iterator[Tuple[int, torch.Tensor]]:
        self._busy = True
        try:
            yield from self._run(host_batches, style_for_batch, alpha)
        finally:
            self._busy = False

    def _run(self, host_batches: Iterable[torch.Tensor],
             style_for_batch: Callable[[int, torch.Tensor], Sequence[torch.Tensor]],
             alpha: float = 1.0) -> Iterator[Tuple[int, torch.Tensor]]:
        """Yields (batch index, stylised batch as a pinned host tensor).  The yielded tensor is a
        pipeline buffer that stays valid until the NEXT result has been requested and handed out
        (there is one more host buffer than batches in flight); copy it to keep it longer.  The drivers
        keep one pipeline per engine (`for_engine`), so the next loop on the same engine reuses -- and
        overwrites -- these buffers too.

        `style_for_batch(i, device_batch)` returns the `[mean, std]` to use for batch i (called on
        the compute stream, so it may itself run GPU work, e.g. encode a style image)."""
        from . import _lib

        pending: List[Tuple[int, int, int]] = []  # (batch index, slot, host slot) whose D2H has been issued
        cur = torch.cuda.current_stream(self.engine.device)
        for st in (self.s_in, self.s_cmp, self.s_out):
            st.wait_stream(cur)  # whatever prepared the inputs / statistics on the caller's stream
        for i, hb in enumerate(host_batches):
            slot = i % self.slots
            hslot = i % (self.slots + 1)
            if self.u8:
                n, h, w, _ = hb.shape
            else:
                n, _, h, w = hb.shape
            fh, fw = _lib.feature_hw(h, w)
            out_shape = (n, 8 * fh, 8 * fw, 3) if self.u8 else (n, 3, 8 * fh, 8 * fw)
            if len(pending) >= self.slots:  # the slot's previous result must have been handed out
                yield self._hand_out(*pending.pop(0))  # (before _buffers may re-shape it for a ragged batch)
            d_in, d_out, h_out = self._buffers(slot, hslot, hb.shape, out_shape)
            with torch.cuda.stream(self.s_in):
                self.s_in.wait_event(self._ev_cmp[slot])  # compute of batch i-slots has read d_in
                d_in.copy_(hb, non_blocking=True)
                self._ev_in[slot].record(self.s_in)
            self.h2d_bytes += hb.numel() * hb.element_size()
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(self._ev_in[slot])
                self.s_cmp.wait_event(self._ev_out[slot])  # download of batch i-slots has read d_out
                stat = style_for_batch(i, d_in)
                if self.u8:
                    self.engine.transfer_u8(d_in, stat, alpha, self.precision, out=d_out)
                else:
                    self.engine.transfer(d_in, stat, alpha, self.precision, out=d_out)
                self._ev_cmp[slot].record(self.s_cmp)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self._ev_cmp[slot])
                h_out.copy_(d_out, non_blocking=True)
                if self.precision == "fp16":
                    self.engine.saturation_snapshot_async(self._sat[hslot:hslot + 1], self.s_out)
                self._ev_out[slot].record(self.s_out)
            self.d2h_bytes += d_out.numel() * d_out.element_size()
            pending.append((i, slot, hslot))
        for item in pending:
            yield self._hand_out(*item)


def overall_transfer(engine: Engine, host_batches: Iterable[torch.Tensor], style_stat, alpha: float = 1.0,
                     precision: str = DEFAULT_PRECISION, u8: bool = False):
    """Inner loop of CCST_OverallStyleTransfer.py:149-167 for one (content domain, style) pair.
    u8 = True: uint8 HWC batches in and out (ToTensor / save_image's quantisation on the GPU)."""
    pipe = TransferPipeline.for_engine(engine, precision, u8=u8)
    stat = [t.to(engine.device) for t in style_stat]
    return pipe.run(host_batches, lambda i, x: stat, alpha)


def single_style_stat(engine: Engine, style_image: torch.Tensor, precision: str = STATS_PRECISION):
    """CCST_SingleStyleTransfer.py:196-205: relu4_1 statistics (biased variance) of ONE style image
    [1,3,h,w] -> [mean, std] each [1,512,1,1].  Stand-alone it defaults to the fp32 engine (statistics
    within 1e-5 of the reference); the transfer loops below pass their own precision -- there the
    statistic is an intermediate of the image path and the image tolerance is what counts."""
    acc = OverallStyleAccumulator(engine, precision)
    acc.add_images(style_image)
    return list(acc.state.finalize())


class _StyleSet:
    """The loop's style images on the device (uploaded once: the reference re-opens the drawn image from
    disk every batch, CCST_SingleStyleTransfer.py:196-198; here the set is fixed for the loop).  Sets larger
    than `limit_bytes` stay on the host and are uploaded when drawn."""

    def __init__(self, engine: Engine, style_images: Sequence[torch.Tensor], limit_bytes: int = 4 << 30):
        self.engine = engine
        self.host = list(style_images)
        total = sum(im.numel() * im.element_size() for im in self.host)
        self.dev = [im.to(engine.device) for im in self.host] if total <= limit_bytes else None
        shapes = {tuple(im.shape[1:]) for im in self.host}
        # same-size sets are stacked so that a per-image draw is one gather
        self.stack = torch.cat(self.dev) if (self.dev is not None and len(shapes) == 1 and
                                             all(im.shape[0] == 1 for im in self.host)) else None

    def get(self, k: int) -> torch.Tensor:
        return self.dev[k] if self.dev is not None else self.host[k].to(self.engine.device, non_blocking=True)

    def gather(self, ks: Sequence[int]) -> Optional[torch.Tensor]:
        if self.stack is None:
            return None
        idx = torch.tensor(list(ks), dtype=torch.int64).to(self.engine.device, non_blocking=True)
        return self.stack.index_select(0, idx)


def single_transfer(engine: Engine, host_batches: Iterable[torch.Tensor], style_images: Sequence[torch.Tensor],
                    alpha: float = 1.0, precision: str = DEFAULT_PRECISION, seed: int = 1):
    """Inner loop of CCST_SingleStyleTransfer.py:176-223: per BATCH one style image drawn with
    python's `random.choice` (seeded like the reference, `:22-26`), its statistics computed on the
    GPU, then the transfer."""
    rng = random.Random(seed)
    pipe = TransferPipeline.for_engine(engine, precision)
    styles = _StyleSet(engine, style_images)

    def stat_for(i, x):
        k = rng.choice(range(len(style_images)))  # == rng.choice(style_images): one _randbelow(len) draw
        return single_style_stat(engine, styles.get(k), precision)

    return pipe.run(host_batches, stat_for, alpha)


def single_transfer_per_image(engine: Engine, host_batches: Iterable[torch.Tensor],
                              style_images: Sequence[torch.Tensor], alpha: float = 1.0,
                              precision: str = DEFAULT_PRECISION, seed: int = 1):
    """BASELINE.json's config 4 as worded ("per-image style sampling"): every IMAGE of a batch gets
    its own randomly drawn style image (the reference draws one per batch,
    CCST_SingleStyleTransfer.py:195).  The drawn style images of a batch (same size) are encoded as
    one batch, their per-image statistics (:199-203, biased variance) come from one pass of the
    statistics kernel, and AdaIN takes them as [N,512,1,1] (`stat_batch_stride = C`)."""
    rng = random.Random(seed)
    pipe = TransferPipeline.for_engine(engine, precision)
    styles = _StyleSet(engine, style_images)

    def stat_for(i, x):
        n = x.shape[0]
        ks = [rng.choice(range(len(style_images))) for _ in range(n)]
        batch = styles.gather(ks)
        if batch is None and len({tuple(style_images[k].shape[1:]) for k in ks}) == 1:
            batch = torch.cat([styles.get(k) for k in ks])
        if batch is not None:
            return list(F_.calc_mean_std_biased(engine.encode(batch, precision)))
        stats = [single_style_stat(engine, styles.get(k), precision) for k in ks]
        return [torch.cat([s[0] for s in stats]), torch.cat([s[1] for s in stats])]

    return pipe.run(host_batches, stat_for, alpha)


def overall_statistics(engine: Engine, host_batches: Iterable[torch.Tensor], precision: str = STATS_PRECISION,
                       group=None):
    """Loop of mean_std_computation_effcientMem.py:117-137 over this rank's share of one client;
    uploads are double-buffered under the encoder.  Returns (mean, std, images seen by all ranks).
    Default engine: "fp16x3" (tensor cores with split f16 operands: statistics within 1e-5 of the reference);
    "fp32" is the CUDA-core engine (same bar, ~8x slower), "fp16"/"bf16" the plain 16-bit tensor-core encoder
    (~4x faster, ~1e-3 relative)."""
    dev = engine.device
    acc = OverallStyleAccumulator(engine, precision)
    s_in, s_cmp = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    bufs: List[Optional[torch.Tensor]] = [None, None]
    ev_in = [torch.cuda.Event(), torch.cuda.Event()]
    ev_cmp = [torch.cuda.Event(), torch.cuda.Event()]
    s_cmp.wait_stream(torch.cuda.current_stream(dev))  # the zeroed Welford state
    for i, hb in enumerate(host_batches):
        s = i & 1
        if bufs[s] is None or bufs[s].shape != hb.shape or bufs[s].dtype != hb.dtype:
            bufs[s] = torch.empty(hb.shape, dtype=hb.dtype, device=dev)  # fp32 NCHW or uint8 NHWC batches
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_cmp[s])
            bufs[s].copy_(hb, non_blocking=True)
            ev_in[s].record(s_in)
        with torch.cuda.stream(s_cmp):
            s_cmp.wait_event(ev_in[s])
            acc.add_images(bufs[s])
            ev_cmp[s].record(s_cmp)
    torch.cuda.current_stream(dev).wait_stream(s_cmp)
    s_cmp.synchronize()
    mean, std = acc.finalize(group)
    if precision in ("fp16", "fp16x3"):
        engine.check_saturation()
    return mean, std, acc.global_img_count
