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


class ImageFolderDataLoader(BaseDataLoader):
    """Directory trees of JPEG / PNG images (reference: image_data_loader.hpp + stb_image decoding,
    tiny_imagenet_data_loader.hpp, imagenet100_data_loader.hpp).

    Two sources, tried in this order:
      * decoded caches ``<split>_x.npy`` (uint8 NHWC) + ``<split>_y.npy`` (int32) -- memory mapped;
      * the image tree itself, decoded lazily per batch by a thread pool (PIL releases the GIL while
        decoding), resized to ``size`` x ``size``.  ``cache()`` writes the caches once.
    Layouts: TinyImageNet (``train/<wnid>/images/*.JPEG``, ``val/images`` + ``val_annotations.txt``,
    ``wnids.txt``) and class-per-directory (``train/<class>/*``, ``val/<class>/*`` -- ImageNet-100)."""

    EXTS = (".jpeg", ".jpg", ".png", ".bmp")

    def __init__(self, size: int = 64, classes: int = 200):
        super().__init__()
        self.img, self.classes = size, classes
        self.x = self.y = None                 # caches (memmap) ...
        self.files: list[str] = []             # ... or the image tree
        self.labels: np.ndarray | None = None
        self.class_names: list[str] = []
        self._pool = None

    # ------------------------------------------------------------------ discovery
    def _scan(self, path: str, train: bool):
        split = "train" if train else "val"
        root = os.path.join(path, split)
        if not os.path.isdir(root):
            raise FileNotFoundError(f"expected {root}/ (image tree) or {split}_x.npy / {split}_y.npy in {path}")
        wn = os.path.join(path, "wnids.txt")
        ann = os.path.join(root, "val_annotations.txt")
        if os.path.exists(wn):
            names = [l.strip() for l in open(wn) if l.strip()]
        elif os.path.isdir(os.path.join(path, "train")):
            names = sorted(d for d in os.listdir(os.path.join(path, "train")) if os.path.isdir(os.path.join(path, "train", d)))
        else:
            names = sorted(d for d in os.listdir(root) if os.path.isdir(os.path.join(root, d)))
        index = {n: i for i, n in enumerate(names)}
        files, labels = [], []
        if not train and os.path.exists(ann):                       # TinyImageNet validation: flat dir + annotation file
            for line in open(ann):
                parts = line.split()
                if len(parts) >= 2 and parts[1] in index:
                    files.append(os.path.join(root, "images", parts[0]))
                    labels.append(index[parts[1]])
        else:
            for n in names:
                d = os.path.join(root, n)
                if os.path.isdir(os.path.join(d, "images")):
                    d = os.path.join(d, "images")
                if not os.path.isdir(d):
                    continue
                for f in sorted(os.listdir(d)):
                    if f.lower().endswith(self.EXTS):
                        files.append(os.path.join(d, f))
                        labels.append(index[n])
        if not files:
            raise FileNotFoundError(f"no images under {root}")
        self.files, self.labels, self.class_names = files, np.asarray(labels, dtype=np.int32), names

    def load_data(self, path: str, train: bool = True):
        split = "train" if train else "val"
        xf, yf = os.path.join(path, f"{split}_x.npy"), os.path.join(path, f"{split}_y.npy")
        if os.path.exists(xf) and os.path.exists(yf):
            self.x = np.load(xf, mmap_mode="r")
            self.y = np.load(yf, mmap_mode="r")
        else:
            self._scan(path, train)
        self.order = None
        self.reset()
        return self

    # ------------------------------------------------------------------ decoding
    def _decode(self, f: str) -> np.ndarray:
        from PIL import Image
        with Image.open(f) as im:
            im = im.convert("RGB")
            if im.size != (self.img, self.img):
                im = im.resize((self.img, self.img), Image.BILINEAR)
            return np.asarray(im, dtype=np.uint8)

    def _decode_many(self, files) -> np.ndarray:
        if self._pool is None:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=max(1, self.num_threads))
        return np.stack(list(self._pool.map(self._decode, files)))

    def cache(self, path: str, train: bool = True, chunk: int = 1024):
        """Decode the whole split once into ``<split>_x.npy`` / ``<split>_y.npy`` (uint8 NHWC / int32)."""
        if not self.files:
            self._scan(path, train)
        split = "train" if train else "val"
        out = np.lib.format.open_memmap(os.path.join(path, f"{split}_x.npy"), mode="w+", dtype=np.uint8,
                                        shape=(len(self.files), self.img, self.img, 3))
        for i in range(0, len(self.files), chunk):
            out[i:i + chunk] = self._decode_many(self.files[i:i + chunk])
        out.flush()
        np.save(os.path.join(path, f"{split}_y.npy"), self.labels)
        return self.load_data(path, train)

    # ------------------------------------------------------------------ BaseDataLoader
    def size(self):
        if self.x is not None:
            return int(self.x.shape[0])
        return len(self.files)

    def get_data_shape(self):
        return (self.img, self.img, 3)

    def get_batch(self, batch_size: int):
        idx = self._next_indices(batch_size)
        if idx is None:
            return None
        order = np.sort(idx)
        if self.x is not None:
            xb, yb = np.asarray(self.x[order]), np.asarray(self.y[order], dtype=np.int32)
        else:
            xb, yb = self._decode_many([self.files[i] for i in order]), self.labels[order]
        x = torch.from_numpy(xb.astype(np.float32) / 255.0)
        y = torch.from_numpy(np.ascontiguousarray(yb, dtype=np.int32))
        if self.augmentation is not None:
            x = self.augmentation.apply_python(x, full=True)
        return x, y


class TinyImageNetDataLoader(ImageFolderDataLoader):
    """TinyImageNet-200, 64 x 64 (tiny_imagenet_data_loader.hpp)."""

    def __init__(self, size: int = 64, classes: int = 200):
        super().__init__(size, classes)


class ImageNet100DataLoader(ImageFolderDataLoader):
    """ImageNet-100 subset, class-per-directory, resized to 224 x 224 (imagenet100_data_loader.hpp)."""

    def __init__(self, size: int = 224, classes: int = 100):
        super().__init__(size, classes)


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


class WiFiDataLoader(BaseDataLoader):
    """UJIIndoorLoc-style CSV (wifi_data_loader.hpp): a block of feature columns (RSSI readings) and a block
    of target columns, regression (float targets, optionally z-scored) or classification (integer class per
    row), with the reference's ``normalize_data`` (feature / target mean and std kept for de-normalisation)."""

    def __init__(self, is_regression: bool = True):
        super().__init__()
        self.is_regression = bool(is_regression)
        self.x = self.y = None
        self.feature_means = self.feature_stds = self.target_means = self.target_stds = None
        self.is_normalized = False

    def load_data(self, path: str, feature_start_col: int = 0, feature_end_col: int = 0, target_start_col: int = 0,
                  target_end_col: int = 0, has_header: bool = True):
        arr = np.loadtxt(path, delimiter=",", skiprows=1 if has_header else 0, dtype=np.float32, ndmin=2)
        ncol = arr.shape[1]
        if feature_end_col <= feature_start_col:             # reference default: all but the last column(s)
            feature_start_col, feature_end_col = 0, ncol - 1
        if target_end_col <= target_start_col:
            target_start_col, target_end_col = feature_end_col, ncol
        self.x = torch.from_numpy(arr[:, feature_start_col:min(feature_end_col, ncol)].copy())
        t = arr[:, target_start_col:min(target_end_col, ncol)]
        if self.is_regression:
            self.y = torch.from_numpy(t.copy())
        else:
            self.y = torch.from_numpy(t[:, 0].astype(np.int32))
        self.is_normalized = False
        self.order = None
        self.reset()
        return self

    def normalize_data(self):
        if self.x is None or self.x.shape[0] == 0:
            raise RuntimeError("no data to normalize")
        self.feature_means = self.x.mean(0)
        self.feature_stds = self.x.std(0, unbiased=False).clamp_min(1e-8)
        self.x = (self.x - self.feature_means) / self.feature_stds
        if self.is_regression:
            self.target_means = self.y.mean(0)
            self.target_stds = self.y.std(0, unbiased=False).clamp_min(1e-8)
            self.y = (self.y - self.target_means) / self.target_stds
        self.is_normalized = True
        return self

    def denormalize_targets(self, pred: torch.Tensor) -> torch.Tensor:
        if not (self.is_normalized and self.is_regression):
            return pred
        return pred * self.target_stds.to(pred.device) + self.target_means.to(pred.device)

    def num_features(self) -> int:
        return 0 if self.x is None else int(self.x.shape[1])

    def num_outputs(self) -> int:
        if self.y is None:
            return 0
        return int(self.y.shape[1]) if self.is_regression else int(self.y.max()) + 1

    def size(self):
        return 0 if self.x is None else int(self.x.shape[0])

    def get_data_shape(self):
        return (self.num_features(),)

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
            return ImageNet100DataLoader().load_data(path, True), ImageNet100DataLoader().load_data(path, False)
        if n in ("wifi", "uji", "uji_indoor_loc"):
            reg = kw.get("is_regression", True)
            tr = WiFiDataLoader(reg).load_data(os.path.join(path, kw.get("train_file", "trainingData.csv")), **kw.get("columns", {}))
            va = WiFiDataLoader(reg).load_data(os.path.join(path, kw.get("val_file", "validationData.csv")), **kw.get("columns", {}))
            return tr, va
        if n in ("regression", "regression_csv"):
            nt = kw.get("n_targets", 1)
            return (RegressionCSVDataLoader(nt).load_data(os.path.join(path, kw.get("train_file", "train.csv"))),
                    RegressionCSVDataLoader(nt).load_data(os.path.join(path, kw.get("val_file", "val.csv"))))
        if n in ("open_webtext", "openwebtext"):
            L = kw.get("seq_len", 1024)
            return OpenWebTextDataLoader(L).load_data(path, True), OpenWebTextDataLoader(L).load_data(path, False)
        if n == "synthetic":
            shape, classes = kw.get("shape", (32, 32, 3)), kw.get("num_classes", 10)
            return (SyntheticDataLoader(shape, classes, kw.get("n", 1024), kw.get("vocab")).load_data(),
                    SyntheticDataLoader(shape, classes, kw.get("n_val", 256), kw.get("vocab"), seed=1).load_data())
        raise ValueError(f"unknown dataset {name}")
