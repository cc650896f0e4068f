from .env import Env  # noqa: F401
