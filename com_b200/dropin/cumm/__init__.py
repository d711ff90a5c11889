from . import tensorview  # noqa: F401
