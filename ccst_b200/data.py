"""Input side of the batch loops: the txt image lists and the loader transform
of the reference, restructured so that only the bytes of the ORIGINAL image cross PCIe.

Reference (style_transfer/AdaIN/cjm_util/):
  * `ImageLoader._dataset_info(txt)` (ImageLoader.py:31-42): one image per line, "<path> <label>\\n";
  * `ImageTestDataset.__getitem__` (:74-85): `Image.open(name).convert('RGB')` then the transform
    `Resize((S, S)) + ToTensor()` (data_helper.py:45-49) on the host, single-threaded.

Here the host only decodes (PIL, out of scope for the GPU); the decoded uint8 HWC image is uploaded at its own
size and `Resize((S, S))` runs on the device (`ccst_resize_pil_bilinear_u8`, bit-exact with Pillow); `ToTensor`
is fused further down (`style_transfer_u8` / `accumulate_u8`).  PACS images are 227x227: 5x fewer bytes than
the 512x512 batch the reference uploads.
"""
from __future__ import annotations

from typing import Iterator, List, Sequence, Tuple

import numpy as np
import torch

from .transfer import resize_input_u8


def dataset_info(txt_labels: str) -> Tuple[List[str], List[int]]:
    """`_dataset_info` (cjm_util/ImageLoader.py:31-42): file names and integer labels of a txt list."""
    with open(txt_labels, "r") as f:
        images_list = f.readlines()
    file_names, labels = [], []
    for row in images_list:
        row = row.split(" ")
        file_names.append(row[0])
        labels.append(int(row[1]))
    return file_names, labels


def load_rgb_u8(path: str) -> np.ndarray:
    """`Image.open(framename).convert('RGB')` (ImageLoader.py:82) as a uint8 HWC array (host decode)."""
    from PIL import Image

    with Image.open(path) as im:
        return np.asarray(im.convert("RGB"))


def resized_batches(images: Sequence[np.ndarray], image_size: int, batch: int, device) -> Iterator[torch.Tensor]:
    """Batches of the loader transform's `Resize((S, S))` output as uint8 [n,S,S,3] DEVICE tensors (what
    `style_transfer_u8` / `OverallStyleAccumulator.add_images` take): every image is uploaded at its
    original size from pinned memory and resized on the GPU."""
    device = torch.device(device)
    for b0 in range(0, len(images), batch):
        chunk = images[b0:b0 + batch]
        out = torch.empty((len(chunk), image_size, image_size, 3), dtype=torch.uint8, device=device)
        for i, im in enumerate(chunk):
            h = torch.from_numpy(np.ascontiguousarray(im)).pin_memory()
            out[i:i + 1] = resize_input_u8(h.to(device, non_blocking=True)[None], image_size)
        yield out


def list_batches(txt_labels: str, data_path: str, image_size: int, batch: int, device):
    """The test loader of data_helper.py:38-43 (unshuffled): yields (uint8 [n,S,S,3] device batch, names)."""
    names, _ = dataset_info(txt_labels)
    for b0 in range(0, len(names), batch):
        frames = [data_path + "/" + nm for nm in names[b0:b0 + batch]]
        imgs = [load_rgb_u8(f) for f in frames]
        yield next(resized_batches(imgs, image_size, len(imgs), device)), frames
