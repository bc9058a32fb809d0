"""`dataset` as the reference's train.py imports it (train.py:18).  The reference's LMDB / folder
datasets need `lmdb` and `imutils` (dataset.py:3,7), which are outside the hot path; this module adds a
`synthetic` type yielding LSUN-shaped tensors and serves folders of images with PIL only."""
import os

import torch
from torch.utils import data


class SyntheticDataset(data.Dataset):
    """(3, R, R) tensors uniform in [-1, 1): the value range the reference's ToTensor + Normalize(0.5, 0.5)
    pipeline produces (train.py:445-451)."""

    def __init__(self, path, transform, resolution, length=1 << 16):
        self.resolution, self.length = resolution, length

    def __len__(self):
        return self.length

    def __getitem__(self, index):
        g = torch.Generator().manual_seed(index)
        return torch.rand(3, self.resolution, self.resolution, generator=g) * 2 - 1


class NormalDataset(data.Dataset):
    EXT = (".jpg", ".jpeg", ".png", ".ppm", ".bmp", ".pgm", ".tif", ".tiff", ".webp")

    def __init__(self, path, transform, resolution):
        self.files = sorted(os.path.join(r, f) for r, _, fs in os.walk(path) for f in fs if f.lower().endswith(self.EXT))
        self.transform, self.resolution = transform, resolution

    def __len__(self):
        return len(self.files)

    def __getitem__(self, index):
        from PIL import Image
        img = Image.open(self.files[index]).convert("RGB").resize((self.resolution, self.resolution))
        return self.transform(img)


def set_dataset(type, path, transform, resolution):
    if type == "synthetic":
        return SyntheticDataset(path, transform, resolution)
    if type == "normal":
        return NormalDataset(path, transform, resolution)
    raise NotImplementedError(f"dataset type {type!r}: only 'synthetic' and 'normal' are served here (lmdb needs the lmdb package)")
