from .model import Model, Config  # noqa: F401
