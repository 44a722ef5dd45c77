"""GenomeMetadataManager -- the JSON store of kmerml/utils/genome_metadata.py:8-94 (same constructor, methods and
record layout), with the per-genome tallies (contigs, total_size, gc_content, n_count; :55-85) computed on the GPU
by kmerml_genome_stats instead of a second Bio.SeqIO pass."""
import json
from datetime import datetime
from pathlib import Path

from .path_utils import ensure_directory_exists, find_files


class GenomeMetadataManager:
    """Collect, store, and retrieve genome metadata across pipeline runs."""

    def __init__(self, metadata_file="data/metadata/genome_metadata.json", *, device=None):
        self.metadata_file = Path(metadata_file)
        ensure_directory_exists(self.metadata_file.parent)
        self.device = device
        self.metadata = self._load_metadata()

    def _load_metadata(self):
        if self.metadata_file.exists():
            with open(self.metadata_file) as fh:
                try:
                    return json.load(fh)
                except json.JSONDecodeError:
                    return {}
        return {}

    def _save_metadata(self):
        with open(self.metadata_file, "w") as fh:
            json.dump(self.metadata, fh, indent=2)

    def collect_metadata(self, genome_dir, refresh=False, patterns=None):
        """Metadata of every genome file under `genome_dir` (existing entries are kept unless `refresh`)."""
        genome_files = find_files(genome_dir, patterns=patterns or ["*.fa", "*.fasta", "*.fna"], recursive=True)
        if not genome_files:
            print("No genome files found. Please check your data/raw/ directory.")
            return None
        for genome_file in genome_files:
            if refresh or genome_file.stem not in self.metadata:
                self.metadata[genome_file.stem] = self._extract_genome_metadata(genome_file)
        self._save_metadata()
        return self.metadata

    def _extract_genome_metadata(self, genome_file):
        import numpy as np
        import torch

        from .. import engine
        dev = torch.device(self.device if self.device is not None else "cuda")
        host = np.fromfile(str(genome_file), dtype=np.uint8)
        data = torch.from_numpy(host).to(dev) if host.size else torch.zeros(0, dtype=torch.uint8, device=dev)
        stats = engine.genome_stats_device(data)
        return {"file_path": str(genome_file), "last_updated": datetime.now().strftime("%Y-%m-%d %H:%M:%S"),
                "contigs": stats["contigs"], "total_size": stats["total_size"], "gc_content": stats["gc_content"],
                "n_count": stats["n_count"]}

    def get_genome_size(self, genome_id):
        entry = self.metadata.get(genome_id)
        return entry["total_size"] if entry is not None else None

    def list_available_genomes(self):
        return list(self.metadata.keys())
