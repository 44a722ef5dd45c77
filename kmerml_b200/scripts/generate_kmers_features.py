"""`python -m kmerml_b200.scripts.generate_kmers_features` -- same flags, defaults, messages
and exit codes as the reference's scripts/generate_kmers_features.py:7-75."""
import argparse
import sys

from ..kmers.statistics import KmerFeatureExtractor
from ..utils.path_utils import find_files

_FEATURE_SETS = {
    "all": None,
    "basic": ["gc_content", "base_counts"],
    "advanced": ["gc_content", "base_counts", "entropy", "cpg_sites", "repeats"],
}


def main(argv=None):
    parser = argparse.ArgumentParser(description="Generate ML features from k-mer files")
    parser.add_argument("--input", "-i", default="data/processed/kmers/", help="Input directory with k-mer files")
    parser.add_argument("--output-dir", "-o", default="data/processed/features",
                        help="Output directory for feature files")
    parser.add_argument("--feature-set", "-f", default="all", help="Feature set to generate (all, basic, advanced)")
    parser.add_argument("--metadata", "-m", default="data/metadata/genome_metadata.json",
                        help="Path to genome metadata file")
    parser.add_argument("--k-values", "-k", default="all",
                        help="Comma-separated list of k values to process (or 'all')")
    args = parser.parse_args(argv)
    if args.k_values.lower() == "all":
        patterns = ["k*.txt"]
    else:
        try:
            patterns = [f"k{int(tok)}.txt" for tok in args.k_values.split(",")]
        except ValueError:
            print("Error: k values must be integers")
            return 1
    kmer_files = find_files(args.input, patterns=patterns, recursive=True)
    if not kmer_files:
        print(f"No k-mer files found in {args.input} matching patterns: {patterns}")
        return 1
    print(f"Found {len(kmer_files)} k-mer files")
    choice = args.feature_set.lower()
    if choice not in _FEATURE_SETS:
        print(f"Unknown feature set: {args.feature_set}")
        return 1
    extractor = KmerFeatureExtractor(input_paths=kmer_files, output_dir=args.output_dir, metadata_file=args.metadata)
    outputs = extractor.extract_features(required_features=_FEATURE_SETS[choice])
    print(f"Generated {len(outputs)} feature files")
    return 0


if __name__ == "__main__":
    sys.exit(main())
