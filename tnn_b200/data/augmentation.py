"""Augmentation strategy chain + builder (reference: include/data_augmentation/augmentation.hpp:17-160
and the nine strategy headers: horizontal/vertical flip, rotation, brightness, contrast, gaussian
noise, random crop (pad 4), cutout, normalize).

Flip / crop / brightness / contrast / noise / cutout / normalize are executed by the native batch
assembler (``native_params``) while the image is decoded; rotation (rare) and the path for
loaders that are not record based run here on the assembled float batch.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


class AugmentationStrategy:
    name = "strategy"

    def native(self) -> dict:
        return {}

    def apply(self, x: torch.Tensor, gen: torch.Generator) -> torch.Tensor:
        return x


class HorizontalFlip(AugmentationStrategy):
    name = "horizontal_flip"

    def __init__(self, p=0.5):
        self.p = p

    def native(self):
        return {"hflip_p": float(self.p)}

    def apply(self, x, gen):
        m = torch.rand(x.shape[0], generator=gen) < self.p
        x = x.clone()
        x[m] = x[m].flip(2)
        return x


class VerticalFlip(AugmentationStrategy):
    name = "vertical_flip"

    def __init__(self, p=0.5):
        self.p = p

    def native(self):
        return {"vflip_p": float(self.p)}

    def apply(self, x, gen):
        m = torch.rand(x.shape[0], generator=gen) < self.p
        x = x.clone()
        x[m] = x[m].flip(1)
        return x


class Rotation(AugmentationStrategy):
    name = "rotation"

    def __init__(self, p=0.5, max_degrees=15.0):
        self.p, self.max_degrees = p, max_degrees

    def apply(self, x, gen):
        B = x.shape[0]
        ang = (torch.rand(B, generator=gen) * 2 - 1) * math.radians(self.max_degrees)
        ang = torch.where(torch.rand(B, generator=gen) < self.p, ang, torch.zeros_like(ang))
        cos, sin = torch.cos(ang), torch.sin(ang)
        theta = torch.stack([torch.stack([cos, -sin, torch.zeros_like(cos)], 1),
                             torch.stack([sin, cos, torch.zeros_like(cos)], 1)], 1)
        xn = x.permute(0, 3, 1, 2)
        grid = F.affine_grid(theta, list(xn.shape), align_corners=False)
        return F.grid_sample(xn, grid, padding_mode="zeros", align_corners=False).permute(0, 2, 3, 1).contiguous()


class Brightness(AugmentationStrategy):
    name = "brightness"

    def __init__(self, p=0.5, delta=0.2):
        self.p, self.delta = p, delta

    def native(self):
        return {"brightness": float(self.delta)}

    def apply(self, x, gen):
        d = (torch.rand(x.shape[0], 1, 1, 1, generator=gen) * 2 - 1) * self.delta
        return (x + d).clamp(0, 1)


class Contrast(AugmentationStrategy):
    name = "contrast"

    def __init__(self, p=0.5, factor=0.2):
        self.p, self.factor = p, factor

    def native(self):
        return {"contrast": float(self.factor)}

    def apply(self, x, gen):
        f = 1 + (torch.rand(x.shape[0], 1, 1, 1, generator=gen) * 2 - 1) * self.factor
        return ((x - 0.5) * f + 0.5).clamp(0, 1)


class GaussianNoise(AugmentationStrategy):
    name = "gaussian_noise"

    def __init__(self, p=0.5, std=0.05):
        self.p, self.std = p, std

    def native(self):
        return {"noise_std": float(self.std)}

    def apply(self, x, gen):
        return (x + torch.randn(x.shape, generator=gen) * self.std).clamp(0, 1)


class RandomCrop(AugmentationStrategy):
    name = "random_crop"

    def __init__(self, p=1.0, padding=4):
        self.p, self.padding = p, padding

    def native(self):
        return {"random_crop_pad": int(self.padding)}

    def apply(self, x, gen):
        B, H, W, C = x.shape
        p = self.padding
        xp = F.pad(x.permute(0, 3, 1, 2), (p, p, p, p))
        out = torch.empty_like(x)
        for b in range(B):
            dy, dx = int(torch.randint(0, 2 * p + 1, (1,), generator=gen)), int(torch.randint(0, 2 * p + 1, (1,), generator=gen))
            out[b] = xp[b, :, dy:dy + H, dx:dx + W].permute(1, 2, 0)
        return out


class Cutout(AugmentationStrategy):
    name = "cutout"

    def __init__(self, p=0.5, size=8):
        self.p, self.size = p, size

    def native(self):
        return {"cutout": int(self.size)}

    def apply(self, x, gen):
        B, H, W, C = x.shape
        x = x.clone()
        for b in range(B):
            if float(torch.rand(1, generator=gen)) < self.p:
                cy, cx = int(torch.randint(0, H, (1,), generator=gen)), int(torch.randint(0, W, (1,), generator=gen))
                x[b, max(0, cy - self.size // 2):cy + self.size // 2 + 1, max(0, cx - self.size // 2):cx + self.size // 2 + 1] = 0
        return x


class Normalize(AugmentationStrategy):
    name = "normalize"

    def __init__(self, mean, std):
        self.mean, self.std = list(mean), list(std)

    def native(self):
        return {"mean": self.mean, "std": self.std}

    def apply(self, x, gen):
        return (x - torch.tensor(self.mean)) / torch.tensor(self.std)


class Augmentation:
    def __init__(self, strategies=None, seed=0):
        self.strategies: list[AugmentationStrategy] = list(strategies or [])
        self.gen = torch.Generator().manual_seed(seed)

    def add(self, s: AugmentationStrategy):
        self.strategies.append(s)
        return self

    # strategy management of the reference (include/data_augmentation/augmentation.hpp:67-97)
    add_augmentation = add

    def remove_augmentation(self, which) -> int:
        """Remove by index (int) or by strategy name (str: every strategy of that name); returns how many were removed."""
        if isinstance(which, int):
            if not 0 <= which < len(self.strategies):
                raise IndexError(f"no strategy at index {which}")
            del self.strategies[which]
            return 1
        before = len(self.strategies)
        self.strategies = [s for s in self.strategies if s.name != which]
        return before - len(self.strategies)

    def clear_augmentations(self):
        self.strategies.clear()

    def size(self) -> int:
        return len(self.strategies)

    def names(self) -> list[str]:
        return [s.name for s in self.strategies]

    def native_params(self) -> dict:
        d = {}
        for s in self.strategies:
            d.update(s.native())
        return d

    def apply_python(self, x: torch.Tensor, full: bool = False) -> torch.Tensor:
        """Strategies the native assembler does not implement (or all of them when ``full``)."""
        for s in self.strategies:
            if full or not s.native():
                x = s.apply(x, self.gen)
        return x

    def apply(self, x: torch.Tensor) -> torch.Tensor:
        return self.apply_python(x, full=True)


class AugmentationBuilder:
    def __init__(self, seed=0):
        self.aug = Augmentation(seed=seed)

    def horizontal_flip(self, p=0.5):
        self.aug.add(HorizontalFlip(p)); return self

    def vertical_flip(self, p=0.5):
        self.aug.add(VerticalFlip(p)); return self

    def rotation(self, p=0.5, max_degrees=15.0):
        self.aug.add(Rotation(p, max_degrees)); return self

    def brightness(self, p=0.5, delta=0.2):
        self.aug.add(Brightness(p, delta)); return self

    def contrast(self, p=0.5, factor=0.2):
        self.aug.add(Contrast(p, factor)); return self

    def gaussian_noise(self, p=0.5, std=0.05):
        self.aug.add(GaussianNoise(p, std)); return self

    def random_crop(self, p=1.0, padding=4):
        self.aug.add(RandomCrop(p, padding)); return self

    def cutout(self, p=0.5, size=8):
        self.aug.add(Cutout(p, size)); return self

    def normalize(self, mean, std):
        self.aug.add(Normalize(mean, std)); return self

    def custom_augmentation(self, strategy: AugmentationStrategy):
        """Any user strategy (a subclass of AugmentationStrategy with ``apply(x, gen)``), applied in sequence with the others."""
        self.aug.add(strategy); return self

    def build(self) -> Augmentation:
        return self.aug
