"""kmerml_b200 -- B200 (sm_100a) implementation of kmer-ml's k-mer extraction ->
counts -> feature-matrix hot path, behind the reference's own Python surface.

    from kmerml_b200.kmers.generate import KmerExtractor          # kmerml.kmers.generate
    from kmerml_b200.kmers.statistics import KmerFeatureExtractor  # kmerml.kmers.statistics
    from kmerml_b200.ml.features import KmerFeatureBuilder         # kmerml.ml.features
    python -m kmerml_b200.scripts.extract_kmers ...                # scripts/extract_kmers.py

`install_as_kmerml()` makes `import kmerml...` resolve to these modules.
The compute path is libkmerml_b200.so (include/kmerml_b200.h); there is no CPU fallback.
"""
import sys

__version__ = "0.1.0"


def install_as_kmerml():
    """Alias this package's drop-in modules under the reference's names in sys.modules."""
    import importlib
    import types

    from . import kmers, ml, utils
    from .kmers import generate, statistics
    from .ml import clustering, features
    from .utils import genome_metadata, kmer_metadata, path_utils
    root = types.ModuleType("kmerml")
    root.__version__ = __version__
    root.__path__ = []
    mapping = {
        "kmerml": root, "kmerml.kmers": kmers, "kmerml.kmers.generate": generate,
        "kmerml.kmers.statistics": statistics, "kmerml.ml": ml, "kmerml.ml.features": features,
        "kmerml.ml.clustering": clustering,
        "kmerml.utils": utils, "kmerml.utils.path_utils": path_utils,
        "kmerml.utils.genome_metadata": genome_metadata, "kmerml.utils.kmer_metadata": kmer_metadata,
    }
    for name, mod in mapping.items():
        sys.modules[name] = mod
    root.kmers, root.ml, root.utils = kmers, ml, utils
    scripts = importlib.import_module(".scripts", __name__)
    sys.modules["scripts"] = scripts
    for sub in ("extract_kmers", "generate_kmers_features"):
        sys.modules[f"scripts.{sub}"] = importlib.import_module(f".scripts.{sub}", __name__)
    return root
