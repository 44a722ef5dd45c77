"""`python -m kmerml_b200.scripts.generate_kmers_features`: the CLI surface (flags, defaults, messages, exit codes)
of the reference's scripts/generate_kmers_features.py:7-75 over the B200 drop-in classes."""
import argparse
import sys

from ..kmers.statistics import KmerFeatureExtractor
from ..utils.path_utils import find_files

# (long flag, short flag, default, help) -- the reference's options, scripts/generate_kmers_features.py:11-28
_OPTIONS = (
    ("--input", "-i", "data/processed/kmers/", "Input directory with k-mer files"),
    ("--output-dir", "-o", "data/processed/features", "Output directory for feature files"),
    ("--feature-set", "-f", "all", "Feature set to generate (all, basic, advanced)"),
    ("--metadata", "-m", "data/metadata/genome_metadata.json", "Path to genome metadata file"),
    ("--k-values", "-k", "all", "Comma-separated list of k values to process (or 'all')"),
)
_BASIC = ["gc_content", "base_counts"]
_FEATURE_SETS = {"all": None, "basic": _BASIC, "advanced": _BASIC + ["entropy", "cpg_sites", "repeats"]}


class _Stop(Exception):
    """Carries the message printed before the script exits with status 1."""


def _file_patterns(spec):
    if spec.lower() == "all":
        return ["k*.txt"]
    try:
        return ["k%d.txt" % int(token) for token in spec.split(",")]
    except ValueError:
        raise _Stop("Error: k values must be integers") from None


def _run(opts):
    patterns = _file_patterns(opts.k_values)
    files = find_files(opts.input, patterns=patterns, recursive=True)
    if not files:
        raise _Stop(f"No k-mer files found in {opts.input} matching patterns: {patterns}")
    print(f"Found {len(files)} k-mer files")
    try:
        wanted = _FEATURE_SETS[opts.feature_set.lower()]
    except KeyError:
        raise _Stop(f"Unknown feature set: {opts.feature_set}") from None
    written = KmerFeatureExtractor(input_paths=files, output_dir=opts.output_dir,
                                   metadata_file=opts.metadata).extract_features(required_features=wanted)
    print(f"Generated {len(written)} feature files")


def main(argv=None):
    parser = argparse.ArgumentParser(description="Generate ML features from k-mer files")
    for long_flag, short_flag, default, text in _OPTIONS:
        parser.add_argument(long_flag, short_flag, default=default, help=text)
    try:
        _run(parser.parse_args(argv))
    except _Stop as stop:
        print(stop)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
