from .features import KmerFeatureBuilder  # noqa: F401
