"""tnn_b200: a Blackwell-native (sm_100a) deep-learning training / inference framework with the
capabilities of tungphambasement/TNN.  See DESIGN.md and SURVEY.md."""
__version__ = "0.1.0"

from .core.config import TConfig  # noqa: F401
from .core.dtype import DType  # noqa: F401
