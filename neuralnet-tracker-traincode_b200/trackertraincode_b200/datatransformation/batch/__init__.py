from ...datasets.batch import FieldCategory, imagelike_categories  # noqa: F401
from .geometric import (  # noqa: F401
    FocusRoi, GeneralFocusRoi, MakeRoiRandomizationParameters, NoRoiRandomization, RandomFocusRoi,
    RoiFocusRandomizationParameters, horizontal_flip_and_rot_90)
from .normalization import normalize_batch, offset_points_by_half_pixel, unnormalize_batch, whiten_batch  # noqa: F401
from .intensity import (  # noqa: F401
    KorniaImageDistortions, OnlyClip, RandomBrightness, RandomContrast, RandomEqualize, RandomGamma, RandomGaussianBlur,
    RandomGaussianNoise, RandomGaussianNoiseWithClipping, RandomPosterize, photometric_f32)
from .misc import PutRoiFromLandmarks  # noqa: F401
from .representation import to_numpy, to_tensor  # noqa: F401
