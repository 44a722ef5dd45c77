"""`python -m kmerml_b200.scripts.extract_kmers` -- same flags, defaults, messages and exit
codes as the reference's scripts/extract_kmers.py:7-66; the counting runs on the GPU.
Added (opt-in): --device, --canonical."""
import argparse
import sys
from pathlib import Path

from ..kmers.generate import KmerExtractor
from ..utils.path_utils import find_files

_OPTIONS = [
    (("--input", "-i"), dict(required=True, help="Input directory or file(s) with genome sequences")),
    (("--pattern", "-p"), dict(default="*.fa,*.fasta", help="Comma-separated patterns to match genome files")),
    (("--output-dir", "-o"), dict(default="data/processed/kmers", help="Output directory for k-mer files")),
    (("--compress", "-c"), dict(action="store_true", help="Compress output files")),
    (("--k-values", "-k"), dict(default="8,9,10,11,12", help="Comma-separated list of k values to extract")),
    (("--recursive", "-r"), dict(action="store_true", help="Search input directory recursively")),
    (("--device",), dict(default=None, help="CUDA device (default: current)")),
    (("--canonical",), dict(action="store_true", help="Count min(k-mer, reverse complement) [extension]")),
]


def main(argv=None):
    parser = argparse.ArgumentParser(description="Extract k-mers from genome files")
    for flags, kw in _OPTIONS:
        parser.add_argument(*flags, **kw)
    args = parser.parse_args(argv)
    try:
        k_values = [int(tok) for tok in args.k_values.split(",")]
    except ValueError:
        print("Error: k values must be integers")
        return 1
    source = Path(args.input)
    if source.is_file():
        genomes = [source]
    else:
        genomes = find_files(args.input, patterns=args.pattern.split(","), recursive=args.recursive)
    if not genomes:
        print(f"No genome files found matching patterns: {args.pattern}")
        return 1
    print(f"Found {len(genomes)} genome files")
    extractor = KmerExtractor(output_dir=args.output_dir, compress=args.compress, device=args.device,
                              canonical=args.canonical)
    for position, genome in enumerate(genomes, start=1):
        print(f"Processing {genome.stem} ({position}/{len(genomes)})")
        extractor.extract_kmers_from_fasta(genome, k_values, organism_id=genome.stem)
    print("K-mer extraction completed successfully")
    return 0


if __name__ == "__main__":
    sys.exit(main())
