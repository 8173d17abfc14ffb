"""Tensor-level entry points, the names of trackertraincode/datatransformation/tensors/__init__.py:1-11."""
from .image_geometric_cv2 import affine_transform_image_cv2, croprescale_image_cv2, UpFilters, DownFilters  # noqa: F401
from .affinetrafo import position_normalization, position_unnormalization, apply_affine2d  # noqa: F401
from .normalization import unwhiten_image, whiten_image  # noqa: F401
from .representation import ensure_image_nchw, ensure_image_nhwc  # noqa: F401
