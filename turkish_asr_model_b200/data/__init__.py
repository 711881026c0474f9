from .preprocessing import AudioPreprocessor, SpecAugment, SpeedPerturbation  # noqa: F401
from .dataset import BucketingSampler, collate_fn  # noqa: F401
