from . import batch  # noqa: F401
from .fused import FusedPoseAugmentation  # noqa: F401
