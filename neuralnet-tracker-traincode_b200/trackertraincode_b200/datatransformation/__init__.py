from . import batch  # noqa: F401
from . import tensors  # noqa: F401
from .fused import FusedPoseAugmentation  # noqa: F401
from .loader import PostprocessingLoader, SampleBySampleLoader, SegmentedCollationDataLoader, TransformedDataset  # noqa: F401
from . import sharding  # noqa: F401
