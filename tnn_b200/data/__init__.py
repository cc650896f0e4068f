from .augmentation import Augmentation, AugmentationBuilder  # noqa: F401
from .loaders import (BaseDataLoader, CIFAR10DataLoader, CIFAR100DataLoader, DataLoaderFactory, MNISTDataLoader,  # noqa: F401
                      OpenWebTextDataLoader, SyntheticDataLoader, TinyImageNetDataLoader, ImageFolderDataLoader,
                      ImageNet100DataLoader, RegressionCSVDataLoader, WiFiDataLoader)
from .tokenizer import Tokenizer  # noqa: F401
