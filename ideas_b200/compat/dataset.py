"""`dataset` as the reference's train.py imports it (train.py:18): ``set_dataset(type, path, transform, resolution)``.

Serves the reference's two dataset types (dataset.py:10-85) with their semantics -- same file order, same
``max_num`` cap, same ``resize((R, R))`` before the transform -- without its hard dependency on ``imutils``:

  * ``normal``: every image file under ``path`` (recursive, sorted), opened with PIL;
  * ``lmdb``:   an LSUN-style LMDB of encoded images, when the ``lmdb`` package is importable (it is not part
                of this image; the loader raises a clear error otherwise).

Plus one addition that the unmodified ``train.py`` argparse can reach (its ``--dataset_type`` only admits
``lmdb`` / ``normal``): a ``--dataset_path`` of the form ``synthetic`` or ``synthetic:<length>`` yields
LSUN-shaped random tensors in [-1, 1) -- the value range of ToTensor + Normalize(0.5, 0.5) (train.py:445-451)."""
import os
from io import BytesIO

import torch
from torch.utils import data

IMG_EXTENSIONS = ("webp", ".png", ".jpg", ".jpeg", ".ppm", ".bmp", ".pgm", ".tif", ".tiff")   # dataset.py:52


class SyntheticDataset(data.Dataset):
    def __init__(self, path, transform, resolution, length=1 << 16):
        self.resolution, self.length = resolution, length

    def __len__(self):
        return self.length

    def __getitem__(self, index):
        g = torch.Generator().manual_seed(index)
        return torch.rand(3, self.resolution, self.resolution, generator=g) * 2 - 1


class NormalDataset(data.Dataset):
    def __init__(self, path, transform, resolution=256, max_num=70000):
        listed = sorted(os.path.join(r, f) for r, _, fs in os.walk(path) for f in fs)
        self.files = [f for f in listed[:max_num] if f.lower().endswith(IMG_EXTENSIONS)]
        self.transform, self.resolution = transform, resolution

    def __len__(self):
        return len(self.files)

    def __getitem__(self, index):
        from PIL import Image
        img = Image.open(self.files[index]).resize((self.resolution, self.resolution))
        return self.transform(img)


class LMDBDataset(data.Dataset):
    def __init__(self, path, transform, resolution=256, max_num=70000):
        try:
            import lmdb
        except ImportError as e:
            raise RuntimeError("dataset type 'lmdb' needs the `lmdb` package, which is not installed") from e
        self.env = lmdb.open(path, max_readers=32, readonly=True, lock=False, readahead=False, meminit=False)
        if not self.env:
            raise IOError("Cannot open lmdb dataset", path)
        self.keys = []
        with self.env.begin(write=False) as txn:
            for idx, (key, _) in enumerate(txn.cursor()):
                self.keys.append(key)
                if idx > max_num:
                    break
        self.transform, self.resolution = transform, resolution

    def __len__(self):
        return len(self.keys)

    def __getitem__(self, index):
        from PIL import Image
        with self.env.begin(write=False) as txn:
            img_bytes = txn.get(self.keys[index])
        img = Image.open(BytesIO(img_bytes)).resize((self.resolution, self.resolution))
        return self.transform(img)


def set_dataset(type, path, transform, resolution):
    if type == "synthetic" or (isinstance(path, str) and path.split(":")[0] == "synthetic"):
        parts = str(path).split(":")
        length = int(parts[1]) if len(parts) > 1 and parts[1] else 1 << 16
        return SyntheticDataset(path, transform, resolution, length)
    if type == "normal":
        return NormalDataset(path, transform, resolution)
    if type == "lmdb":
        return LMDBDataset(path, transform, resolution)
    raise NotImplementedError(type)
