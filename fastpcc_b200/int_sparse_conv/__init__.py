from . import cuda_ops, ext  # noqa: F401
