"""KmerFeatureBuilder -- drop-in for the reference's kmerml/ml/features.py:13-117, plus the
GPU-resident matrix operations the reference only gestures at (tests/test_ml.py:8-12):
normalize(), distance_matrix(), filter_features(), get_top_features(), from_counts().
"""
from pathlib import Path
from typing import Union

import numpy as np
import pandas as pd

from ..utils.path_utils import find_files


class KmerFeatureBuilder:
    """Convert k-mer statistics (CSV files, or GPU count rows) into ML-ready matrices."""

    def __init__(self, stats_dir: Union[str, Path] = None):
        self.stats_dir = Path(stats_dir) if stats_dir else None
        self.feature_matrix = None
        self.organisms = []
        self.kmers = []
        self._device_counts = None        # (int32 CUDA tensor of uint32 counts, int64 totals) when built on GPU

    # ---------------------------------------------------------- reference surface
    def build_from_statistics_files(self, metric: str = "count",
                                    file_pattern: str = "*kmer_features.csv") -> pd.DataFrame:
        """organisms x k-mers DataFrame from <organism>_kmer_features.csv files
        (features.py:28-74: per file dict(zip(kmer, metric)) -- later rows win -- then the
        sorted union of k-mer strings as columns, 0 where absent)."""
        if not self.stats_dir:
            raise ValueError("Statistics directory not set")
        files = find_files(self.stats_dir, patterns=[file_pattern], recursive=True)
        if not files:
            raise ValueError(f"No statistics files found matching pattern: {file_pattern}")
        per_organism = {}
        for path in files:
            organism = self._extract_organism_id(path)
            try:
                table = pd.read_csv(path)
                if "kmer" not in table.columns or metric not in table.columns:
                    raise ValueError(f"Required columns not found in {path}. Available: {', '.join(table.columns)}")
                per_organism[organism] = table.drop_duplicates("kmer", keep="last").set_index("kmer")[metric]
            except Exception as exc:
                print(f"Error processing {path}: {exc}")
        return self._build_matrix(per_organism)

    @staticmethod
    def _extract_organism_id(file_path: Path) -> str:
        parts = Path(file_path).stem.split("_")
        return f"{parts[0]}_{parts[1]}" if len(parts) >= 2 else Path(file_path).stem

    def _build_matrix(self, organism_data) -> pd.DataFrame:
        """organism_data: {organism: mapping k-mer -> value} (dict or Series)."""
        series = {org: (v if isinstance(v, pd.Series) else pd.Series(v, dtype=object if not len(v) else None))
                  for org, v in organism_data.items()}
        columns = sorted(set().union(*[s.index for s in series.values()])) if series else []
        self.organisms = list(series.keys())
        if series:
            frame = pd.DataFrame({org: s.reindex(columns) for org, s in series.items()}).T
            frame = frame.reindex(index=self.organisms, columns=columns)
            ints = all(pd.api.types.is_integer_dtype(s.dtype) for s in series.values())
            frame = frame.fillna(0)
            frame = frame.astype(np.int64) if ints else frame.astype(np.float64)
        else:
            frame = pd.DataFrame([], index=[], columns=[])
        self.feature_matrix = frame
        self.kmers = columns
        self._device_counts = None
        return self.feature_matrix

    # ------------------------------------------------------------- GPU extensions
    def from_counts(self, result, organisms, k, observed_only=True):
        """Feature matrix straight from GPU count rows (engine.DenseResult): columns are the
        k-mer strings in lexicographic order -- the order _build_matrix produces."""
        import torch
        lay_off = 0
        for kk in result.k_list:
            if kk == k:
                break
            lay_off += 4 ** kk
        counts = result.counts[:, lay_off:lay_off + 4 ** k]
        ki = result.k_list.index(k)
        host = counts.cpu().numpy().view(np.uint32).astype(np.int64)
        keep = np.nonzero(host.any(axis=0))[0] if observed_only else np.arange(4 ** k)
        letters = np.frombuffer(b"ACGT", dtype=np.uint8)
        names = np.empty((keep.size, k), dtype=np.uint8)
        for i in range(k):
            names[:, i] = letters[(keep >> (2 * (k - 1 - i))) & 3]
        self.kmers = names.view(f"S{k}").ravel().astype(f"U{k}").tolist() if keep.size else []
        self.organisms = list(organisms)
        self.feature_matrix = pd.DataFrame(host[:, keep], index=self.organisms, columns=self.kmers)
        if counts.is_cuda:
            idx = torch.as_tensor(keep, device=counts.device)
            self._device_counts = (counts.index_select(1, idx).contiguous(), result.totals[:, ki].clone())
        return self.feature_matrix

    def normalize(self, method="frequency"):
        """Row-normalised copy of the feature matrix (frequency = row / row sum)."""
        if self.feature_matrix is None:
            raise ValueError("No feature matrix built")
        if method != "frequency":
            raise ValueError(f"Unknown normalisation method: {method}")
        if self._device_counts is not None:
            from .. import engine
            counts, _ = self._device_counts
            sums = (counts.to(__import__("torch").int64) & 0xFFFFFFFF).sum(dim=1)
            freq = engine.normalize_rows_device(counts, sums)
            return pd.DataFrame(freq.cpu().numpy(), index=self.organisms, columns=self.kmers)
        m = self.feature_matrix.astype(np.float64)
        sums = m.sum(axis=1).replace(0, 1.0)
        return m.div(sums, axis=0)

    def distance_matrix(self, metric="cosine", normalized=True, device=None):
        """organisms x organisms distance DataFrame (cosine or euclidean), computed on the GPU
        with float64 accumulation."""
        if self.feature_matrix is None:
            raise ValueError("No feature matrix built")
        import torch
        from .. import engine
        if self._device_counts is not None and (metric == "cosine" or not normalized):
            x = self._device_counts[0]                       # cosine is scale-free: exact integer Gram
        else:
            base = self.normalize() if normalized else self.feature_matrix
            dev = torch.device(device if device is not None else "cuda")
            x = torch.as_tensor(np.ascontiguousarray(base.to_numpy(dtype=np.float64)), device=dev)
        d = engine.pairwise_distance_device(x, metric, out_dtype=torch.float64)
        return pd.DataFrame(d.cpu().numpy(), index=self.organisms, columns=self.organisms)

    def _column_stats(self):
        """(prevalence, population variance) per column: on the GPU when the matrix came from GPU count rows
        (kmerml_column_stats), else pandas."""
        m = self.feature_matrix
        if self._device_counts is not None:
            from .. import engine
            nnz, _, var = engine.column_stats_device(self._device_counts[0])
            n = max(len(self.organisms), 1)
            return nnz.cpu().numpy().astype(np.float64) / n, var.cpu().numpy()
        return (m != 0).mean(axis=0).to_numpy(), m.var(axis=0, ddof=0).to_numpy()

    def filter_features(self, min_prevalence=0.0, min_variance=0.0):
        """Columns present in at least min_prevalence of the organisms with variance >= min_variance."""
        if self.feature_matrix is None:
            raise ValueError("No feature matrix built")
        prev, var = self._column_stats()
        keep = (prev >= min_prevalence) & (var >= min_variance)
        return self.feature_matrix.loc[:, keep]

    def get_top_features(self, n_features=500, method="variance"):
        """The n_features columns of largest variance, ties in column order."""
        if self.feature_matrix is None:
            raise ValueError("No feature matrix built")
        if method != "variance":
            raise ValueError(f"Unknown method: {method}")
        _, var = self._column_stats()
        top = np.argsort(-var, kind="stable")[:n_features]
        return self.feature_matrix.iloc[:, top]
