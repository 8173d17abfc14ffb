from .bfm import HeadModel  # noqa: F401
