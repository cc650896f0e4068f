"""Data loaders (reference: include/data_loading/*.hpp, src/data_loading/data_loading.cpp:17-95).

Image datasets are memory-mapped record files handled by the C++ data path
(csrc/runtime/data_loader.hpp): a batch is assembled by worker threads straight into a PINNED
host buffer (uint8 CHW record -> float NHWC /255 + augmentation), then copied to the GPU by the
trainer's prefetcher.  Labels are int32 class ids (what CrossEntropyLoss consumes); the
OpenWebText loader yields int32 token ids ``x[B, L]`` and next-token ids ``y[B, L]`` (the
reference emits float ids and dense one-hot targets).

Every loader exposes the reference's BaseDataLoader surface: ``load_data, get_batch(B) ->
(x, y) | None, reset, shuffle, size, get_data_shape, set_seed, set_augmentation``.
"""
from __future__ import annotations

import glob
import os

import numpy as np
import torch

from ..ops import native
from .augmentation import Augmentation


class BaseDataLoader:
    def __init__(self):
        self.seed = 0
        self.epoch = 0
        self.cursor = 0
        self.order: np.ndarray | None = None
        self.augmentation: Augmentation | None = None
        self.num_threads = min(8, os.cpu_count() or 1)
        self.pin = torch.cuda.is_available()

    def size(self) -> int:
        raise NotImplementedError

    def get_data_shape(self) -> tuple:
        raise NotImplementedError

    def set_seed(self, seed: int):
        self.seed = int(seed)

    def set_augmentation(self, aug: Augmentation | None):
        self.augmentation = aug

    def reset(self):
        self.cursor = 0
        if self.order is None:
            self.order = np.arange(self.size(), dtype=np.int64)

    def shuffle(self):
        rng = np.random.default_rng(self.seed + self.epoch)
        self.order = rng.permutation(self.size()).astype(np.int64)
        self.epoch += 1
        self.cursor = 0

    def _next_indices(self, batch_size: int):
        if self.order is None:
            self.reset()
        if self.cursor >= self.size():
            return None
        idx = self.order[self.cursor:self.cursor + batch_size]
        self.cursor += len(idx)
        return np.ascontiguousarray(idx)

    def get_batch(self, batch_size: int):
        raise NotImplementedError

    def __iter__(self):
        return self

    def __len__(self):
        return self.size()


class RecordImageLoader(BaseDataLoader):
    """Fixed-size ``{label bytes, C*H*W uint8 CHW}`` records through the native dataset."""

    def __init__(self, C, H, W, label_bytes):
        super().__init__()
        self.C, self.H, self.W = C, H, W
        self.ds = native.require().RecordImageDataset(C, H, W, label_bytes)
        self._bufs = {}

    def add_file(self, path: str):
        self.ds.add_file(path)

    def size(self):
        return self.ds.size()

    def get_data_shape(self):
        return (self.H, self.W, self.C)

    def _buffers(self, B):
        key = B
        if key not in self._bufs:
            ring = []
            for _ in range(3):      # small ring so a batch can still be in flight to the GPU
                x = torch.empty((B, self.H, self.W, self.C), dtype=torch.float32, pin_memory=self.pin)
                y = torch.empty((B,), dtype=torch.int32, pin_memory=self.pin)
                ring.append((x, y))
            self._bufs[key] = [ring, 0]
        ring, i = self._bufs[key]
        self._bufs[key][1] = (i + 1) % len(ring)
        return ring[i]

    def get_batch(self, batch_size: int):
        idx = self._next_indices(batch_size)
        if idx is None:
            return None
        B = len(idx)
        x, y = self._buffers(batch_size)
        aug = self.augmentation.native_params() if self.augmentation is not None else {}
        self.ds.assemble(idx, x.data_ptr(), y.data_ptr(), aug, self.seed * 7919 + self.epoch, self.num_threads)
        x, y = x[:B], y[:B]
        if self.augmentation is not None:
            x = self.augmentation.apply_python(x)
        return x, y


class CIFAR10DataLoader(RecordImageLoader):
    """data_batch_{1..5}.bin / test_batch.bin: 1 label byte + 3072 pixel bytes per record."""

    def __init__(self):
        super().__init__(3, 32, 32, 1)

    def load_data(self, path: str, train: bool = True):
        sub = os.path.join(path, "cifar-10-batches-bin")       # layout of the official archive
        if os.path.isdir(sub):
            path = sub
        files = sorted(glob.glob(os.path.join(path, "data_batch_*.bin"))) if train else [os.path.join(path, "test_batch.bin")]
        if os.path.isfile(path):
            files = [path]
        if not files or not all(os.path.exists(f) for f in files):
            raise FileNotFoundError(f"CIFAR-10 binaries not found under {path}")
        for f in files:
            self.add_file(f)
        self.reset()
        return self


class CIFAR100DataLoader(RecordImageLoader):
    """train.bin / test.bin: coarse label, fine label, 3072 pixel bytes (fine label is used)."""

    def __init__(self):
        super().__init__(3, 32, 32, 2)

    def load_data(self, path: str, train: bool = True):
        sub = os.path.join(path, "cifar-100-binary")
        if os.path.isdir(sub):
            path = sub
        f = path if os.path.isfile(path) else os.path.join(path, "train.bin" if train else "test.bin")
        if not os.path.exists(f):
            raise FileNotFoundError(f)
        self.add_file(f)
        self.reset()
        return self


class MNISTDataLoader(RecordImageLoader):
    """CSV ``label,p0..p783`` -> [N, 28, 28, 1] (reference: mnist CSV loader)."""

    def __init__(self):
        super().__init__(1, 28, 28, 1)

    def load_data(self, path: str, train: bool = True):
        f = path if os.path.isfile(path) else os.path.join(path, "mnist_train.csv" if train else "mnist_test.csv")
        if not os.path.exists(f):
            raise FileNotFoundError(f)
        self.ds.add_mnist_csv(f)
        self.reset()
        return self


class TinyImageNetDataLoader(BaseDataLoader):
    """TinyImageNet / ImageNet-100 directory trees of JPEGs.  JPEG decoding needs an image codec;
    this image ships none usable offline (no PIL/stb), so the loader consumes the ``.npy`` / raw
    record caches produced by ``python -m tnn_b200.data.prepare`` and raises a clear error otherwise."""

    def __init__(self, size: int = 64, classes: int = 200):
        super().__init__()
        self.img, self.classes = size, classes
        self.x = self.y = None

    def load_data(self, path: str, train: bool = True):
        split = "train" if train else "val"
        xf, yf = os.path.join(path, f"{split}_x.npy"), os.path.join(path, f"{split}_y.npy")
        if not (os.path.exists(xf) and os.path.exists(yf)):
            raise FileNotFoundError(f"expected decoded caches {xf} / {yf} (uint8 NHWC + int32 labels)")
        self.x = np.load(xf, mmap_mode="r")
        self.y = np.load(yf, mmap_mode="r")
        self.reset()
        return self

    def size(self):
        return 0 if self.x is None else int(self.x.shape[0])

    def get_data_shape(self):
        return (self.img, self.img, 3)

    def get_batch(self, batch_size: int):
        idx = self._next_indices(batch_size)
        if idx is None:
            return None
        order = np.sort(idx)
        x = torch.from_numpy(np.asarray(self.x[order], dtype=np.float32) / 255.0)
        y = torch.from_numpy(np.asarray(self.y[order], dtype=np.int32))
        if self.augmentation is not None:
            x = self.augmentation.apply_python(x, full=True)
        return x, y


class OpenWebTextDataLoader(BaseDataLoader):
    """uint16 token stream (.bin) -> random windows x[B, L], y = next tokens (int32)."""

    def __init__(self, seq_len: int = 1024):
        super().__init__()
        self.seq_len = seq_len
        self.ds = None
        self._step = 0

    def load_data(self, path: str, train: bool = True):
        f = path if os.path.isfile(path) else os.path.join(path, "train.bin" if train else "val.bin")
        if not os.path.exists(f):
            raise FileNotFoundError(f)
        self.ds = native.require().TokenDataset(f)
        self.reset()
        return self

    def size(self):
        return 0 if self.ds is None else max(0, self.ds.num_tokens() // self.seq_len)

    def get_data_shape(self):
        return (self.seq_len,)

    def get_batch(self, batch_size: int):
        if self.cursor >= self.size():
            return None
        B = min(batch_size, self.size() - self.cursor)
        self.cursor += B
        x = torch.empty((B, self.seq_len), dtype=torch.int32, pin_memory=self.pin)
        y = torch.empty((B, self.seq_len), dtype=torch.int32, pin_memory=self.pin)
        self._step += 1
        self.ds.sample(B, self.seq_len, self.seed * 1000003 + self.epoch * 7919 + self._step, x.data_ptr(), y.data_ptr())
        return x, y

    def window(self, start: int, length: int) -> torch.Tensor:
        x = torch.empty((length,), dtype=torch.int32)
        self.ds.window(start, length, x.data_ptr())
        return x


class RegressionCSVDataLoader(BaseDataLoader):
    """Numeric CSV (WiFi / UJI style): last ``n_targets`` columns are regression targets."""

    def __init__(self, n_targets: int = 1):
        super().__init__()
        self.n_targets = n_targets
        self.x = self.y = None

    def load_data(self, path: str, train: bool = True):
        arr = np.loadtxt(path, delimiter=",", skiprows=1, dtype=np.float32)
        self.x = torch.from_numpy(arr[:, :-self.n_targets].copy())
        self.y = torch.from_numpy(arr[:, -self.n_targets:].copy())
        self.reset()
        return self

    def size(self):
        return 0 if self.x is None else self.x.shape[0]

    def get_data_shape(self):
        return (self.x.shape[1],)

    def get_batch(self, batch_size: int):
        idx = self._next_indices(batch_size)
        if idx is None:
            return None
        i = torch.from_numpy(idx)
        return self.x[i], self.y[i]


class SyntheticDataLoader(BaseDataLoader):
    """Random images / tokens of a given shape (benchmarks, tests, no dataset on disk)."""

    def __init__(self, shape, num_classes, n=1024, token_vocab: int | None = None, seed=0):
        super().__init__()
        self.shape, self.num_classes, self.n, self.vocab = tuple(shape), num_classes, n, token_vocab
        self.seed = seed

    def load_data(self, path: str = "", train: bool = True):
        self.reset()
        return self

    def size(self):
        return self.n

    def get_data_shape(self):
        return self.shape

    def get_batch(self, batch_size: int):
        idx = self._next_indices(batch_size)
        if idx is None:
            return None
        g = torch.Generator().manual_seed(self.seed * 100003 + int(idx[0]))
        B = len(idx)
        if self.vocab:
            x = torch.randint(0, self.vocab, (B, *self.shape), generator=g, dtype=torch.int32)
            y = torch.randint(0, self.vocab, (B, *self.shape), generator=g, dtype=torch.int32)
        else:
            x = torch.rand((B, *self.shape), generator=g)
            y = torch.randint(0, self.num_classes, (B,), generator=g, dtype=torch.int32)
        if self.pin:
            x, y = x.pin_memory(), y.pin_memory()
        return x, y


class DataLoaderFactory:
    """``create(name, path) -> (train_loader, val_loader)`` (data_loading.cpp:17-95)."""

    @staticmethod
    def create(name: str, path: str = "", **kw):
        n = name.lower()
        if n == "mnist":
            return MNISTDataLoader().load_data(path, True), MNISTDataLoader().load_data(path, False)
        if n == "cifar10":
            return CIFAR10DataLoader().load_data(path, True), CIFAR10DataLoader().load_data(path, False)
        if n == "cifar100":
            return CIFAR100DataLoader().load_data(path, True), CIFAR100DataLoader().load_data(path, False)
        if n in ("tiny_imagenet", "tinyimagenet"):
            return TinyImageNetDataLoader(64, 200).load_data(path, True), TinyImageNetDataLoader(64, 200).load_data(path, False)
        if n in ("imagenet100", "imagenet_100"):
            return TinyImageNetDataLoader(224, 100).load_data(path, True), TinyImageNetDataLoader(224, 100).load_data(path, False)
        if n in ("open_webtext", "openwebtext"):
            L = kw.get("seq_len", 1024)
            return OpenWebTextDataLoader(L).load_data(path, True), OpenWebTextDataLoader(L).load_data(path, False)
        if n == "synthetic":
            shape, classes = kw.get("shape", (32, 32, 3)), kw.get("num_classes", 10)
            return (SyntheticDataLoader(shape, classes, kw.get("n", 1024), kw.get("vocab")).load_data(),
                    SyntheticDataLoader(shape, classes, kw.get("n_val", 256), kw.get("vocab"), seed=1).load_data())
        raise ValueError(f"unknown dataset {name}")
