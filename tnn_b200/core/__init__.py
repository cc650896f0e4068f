from .config import TConfig  # noqa: F401
from .dtype import DType, dtype_eps, string_to_dtype, dtype_to_string  # noqa: F401
from .matrix import Matrix  # noqa: F401,E402
