"""KmerMetadataManager -- kmerml/utils/kmer_metadata.py:4-117 (same constructor, methods and record layout): the
count summaries of every k{k}.txt (:59-78) from the GPU: the file's lines are parsed by kmerml_parse_kmer_lines and
reduced by kmerml_count_stats (total / unique / max / min in one pass, the median by an exact radix select)."""
import gzip
import json
import re
from collections import defaultdict
from pathlib import Path

from .genome_metadata import GenomeMetadataManager


class KmerMetadataManager:
    """Manage basic k-mer metadata and integrate with genome metadata."""

    def __init__(self, metadata_file="data/metadata/genome_metadata.json", *, device=None):
        self.metadata_file = Path(metadata_file)
        self.device = device
        self.genome_manager = GenomeMetadataManager(metadata_file, device=device)
        self.metadata = self.genome_manager.metadata

    def add_kmer_metadata(self, kmer_files, recalculate=False):
        """Summaries of the given k-mer files, stored under metadata[organism]["kmers"][str(k)]."""
        for organism, files in self._group_files_by_organism(kmer_files).items():
            entry = self.metadata.setdefault(organism, {})
            if "kmers" not in entry or recalculate:
                entry["kmers"] = {}
            for kmer_file in files:
                k_val = self._extract_k_from_filename(kmer_file.name)
                if k_val is None or (str(k_val) in entry["kmers"] and not recalculate):
                    continue
                entry["kmers"][str(k_val)] = self._calculate_basic_stats(kmer_file, k_val)
        self._save_metadata()
        return self.metadata

    def _calculate_basic_stats(self, kmer_file, k_val):
        import ctypes

        import torch

        from .. import _lib, engine
        raw = gzip.open(kmer_file, "rb").read() if str(kmer_file).endswith(".gz") else Path(kmer_file).read_bytes()
        dev = torch.device(self.device if self.device is not None else "cuda")
        L = _lib.load()
        ctx = _lib.context(dev.index if dev.index is not None else torch.cuda.current_device())
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        text = torch.frombuffer(bytearray(raw or b"\n"), dtype=torch.uint8).to(dev)
        ends = torch.nonzero(text == 10).flatten()
        if raw and raw[-1] != 10:
            ends = torch.cat([ends, torch.tensor([len(raw)], dtype=torch.int64, device=dev)])
        n = int(ends.numel()) if raw else 0
        value = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
        count = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
        bad = ctypes.c_uint32(0)
        _lib.check(L.kmerml_parse_kmer_lines(ctx.handle, text.data_ptr(), ends.data_ptr(), n, value.data_ptr(), count.data_ptr(),
                                             ctypes.byref(bad), stream))
        if bad.value or n == 0 or int(count[:n].max().item()) >= 1 << 32:
            raise ValueError(f"{kmer_file}: not a k-mer count file of '<digits>\\t<count>' lines")
        stats = engine.kmer_count_stats_device(count[:n].to(torch.int32), k_val)       # (uint32 storage)
        return {"file_path": str(kmer_file), **stats}

    @staticmethod
    def _group_files_by_organism(kmer_files):
        groups = defaultdict(list)
        for file_path in kmer_files:
            path = Path(file_path)
            groups[path.parent.name].append(path)
        return groups

    @staticmethod
    def _extract_k_from_filename(filename):
        m = re.search(r"k(\d+)", str(filename))
        return int(m.group(1)) if m else None

    def _save_metadata(self):
        with open(self.metadata_file, "w") as fh:
            json.dump(self.metadata, fh, indent=2)
