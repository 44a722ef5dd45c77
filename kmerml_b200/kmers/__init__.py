from .generate import KmerExtractor  # noqa: F401
from .statistics import KmerFeatureExtractor  # noqa: F401
