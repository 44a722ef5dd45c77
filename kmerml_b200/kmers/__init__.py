"""kmerml.kmers -- KmerExtractor (GPU) and KmerFeatureExtractor (host pandas / GPU text), imported lazily so that
the feature extractor can be used without torch, like the reference's (kmerml/kmers/__init__.py)."""
import importlib

_LAZY = {"KmerExtractor": "generate", "KmerFeatureExtractor": "statistics", "generate": "generate", "statistics": "statistics"}


def __getattr__(name):
    if name not in _LAZY:
        raise AttributeError(name)
    mod = importlib.import_module(f"{__name__}.{_LAZY[name]}")
    return mod if name == _LAZY[name] else getattr(mod, name)
