from .affine2d import Affine2d  # noqa: F401
