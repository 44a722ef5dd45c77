"""KmerFeatureExtractor -- drop-in for the reference's kmerml/kmers/statistics.py:9-273.

Reads k{k}.txt[.gz] files, computes the per-k-mer feature columns and writes one
<organism>_kmer_features.csv per organism, with the reference's column order and its
"bug-for-bug" behaviours: the digit column is parsed by pandas as an integer
(statistics.py:261-271), so leading zeros (leading 'A's) are lost before decoding
(:158,:248-251) and every feature is computed on that shortened string.

The reference spends ~70 us per k-mer in df.iterrows(); here every column is computed
for all rows at once.  (The features are functions of the k-mer string only; the same
integers per k-mer index are also available on the GPU as
engine.static_features_device / kmerml_static_features for in-memory pipelines.)
"""
import json
import math
import re
from collections import defaultdict
from pathlib import Path

import numpy as np
import pandas as pd

ALL_FEATURES = ["base_counts", "gc_content", "cpg_sites", "entropy", "repeats", "presence"]
_LETTER = np.frombuffer(b"ATCGNNNNNN", dtype=np.uint8)       # digit -> letter (statistics.py:250)
_K_IN_NAME = re.compile(r"k(\d+)")


class _GenomeSizes:
    """The slice of GenomeMetadataManager the feature extractor uses
    (kmerml/utils/genome_metadata.py:11-28,87-91): a JSON map organism -> {"total_size"}."""

    def __init__(self, metadata_file):
        self.path = Path(metadata_file)
        self.path.parent.mkdir(parents=True, exist_ok=True)
        self.metadata = {}
        if self.path.exists():
            try:
                with open(self.path) as fh:
                    self.metadata = json.load(fh)
            except json.JSONDecodeError:
                self.metadata = {}

    def get_genome_size(self, genome_id):
        entry = self.metadata.get(genome_id)
        return entry["total_size"] if entry is not None else None


def _entropy_of(counts, n):
    """-sum p log2 p over the letters present (statistics.py:220-226), Python floats."""
    h = 0
    for c in counts:
        if c > 0:
            p = c / n
            h -= p * math.log2(p)
    return h


def features_from_digits(values, wanted):
    """Feature columns for an int64 array of digit-encoded k-mers (leading zeros already
    lost, exactly what the reference's DataFrame holds).  Returns an ordered dict of
    column name -> numpy array, starting with the decoded 'kmer' strings."""
    v = np.asarray(values, dtype=np.int64)
    rows = v.size
    if rows and v.min() < 0:
        raise ValueError("negative k-mer codes")
    nd = np.ones(rows, dtype=np.int64)
    t = v // 10
    while t.any():
        nd += t > 0
        t //= 10
    width = int(nd.max()) if rows else 1
    # digit matrix, most significant first, right-aligned rows padded with 255 on the left
    digits = np.full((rows, width), 255, dtype=np.uint8)
    t = v.copy()
    for col in range(width - 1, -1, -1):
        live = (width - 1 - col) < nd
        digits[live, col] = (t[live] % 10).astype(np.uint8)
        t //= 10
    letters = np.where(digits == 255, 0, _LETTER[np.minimum(digits, 9)]).astype(np.uint8)
    kmer = np.empty(rows, dtype=object)
    for n in np.unique(nd):
        sel = nd == n
        block = np.ascontiguousarray(letters[sel][:, width - n:])
        kmer[sel] = block.view(f"S{int(n)}").ravel().astype(f"U{int(n)}")
    cols = {"kmer": kmer}
    n_a = (digits == 0).sum(1)
    n_t = (digits == 1).sum(1)
    n_c = (digits == 2).sum(1)
    n_g = (digits == 3).sum(1)
    n_n = ((digits >= 4) & (digits != 255)).sum(1)
    cpg = ((digits[:, :-1] == 2) & (digits[:, 1:] == 3)).sum(1) if width > 1 else np.zeros(rows, np.int64)
    rep = np.zeros(rows, dtype=bool)
    sym = np.where(digits == 255, 250, np.minimum(digits, 4))           # every non-ACGT digit decodes to 'N'
    for i in range(width - 3):
        both = (digits[:, i] != 255)
        rep |= both & (sym[:, i] == sym[:, i + 2]) & (sym[:, i + 1] == sym[:, i + 3])
    cols.update(columns_from_composition(nd, n_a, n_c, n_g, n_t, n_n, cpg, rep.astype(np.int64), wanted))
    return cols


def columns_from_composition(nd, n_a, n_c, n_g, n_t, n_n, cpg, rep, wanted):
    """The feature columns of statistics.py:149-240 from a k-mer's composition (length, letter counts, CpG count,
    repeat flag): all of them are functions of it.  Shared by the pandas path (per row) and the GPU text path (per
    composition class, kmerml_feature_keys)."""
    nd = np.asarray(nd, dtype=np.int64)
    n_a, n_c, n_g, n_t, n_n = (np.asarray(x, dtype=np.int64) for x in (n_a, n_c, n_g, n_t, n_n))
    cpg = np.asarray(cpg, dtype=np.int64)
    nf = nd.astype(np.float64)
    cols = {}
    if "gc_content" in wanted:
        cols["gc_percent"] = ((n_g + n_c) / nf) * 100
    if "base_counts" in wanted:
        cols["A_count"], cols["C_count"], cols["G_count"], cols["T_count"] = n_a, n_c, n_g, n_t
    if "presence" in wanted:
        for name, cnt in (("A", n_a), ("C", n_c), ("G", n_g), ("T", n_t)):
            cols[f"{name}_present"] = (cnt > 0).astype(np.int64)
    if "cpg_sites" in wanted:
        cols["cpg_count"] = cpg
        prod = (n_c / nf) * (n_g / nf)
        expected = np.where(prod > 0, prod * (nd - 1), 0.001)
        with np.errstate(divide="ignore", invalid="ignore"):
            cols["cpg_obs_exp"] = np.where(expected > 0, cpg / np.where(expected > 0, expected, 1.0), 0.0)
    if "entropy" in wanted:
        # (counts are < 64: one packed integer per composition instead of a row-wise unique)
        key = ((((nd * 64 + n_a) * 64 + n_c) * 64 + n_g) * 64 + n_t) * 64 + n_n
        uniq, inverse = np.unique(key, return_inverse=True)
        table = np.empty(uniq.size, dtype=np.float64)
        for i, u in enumerate(uniq.tolist()):
            parts = []
            for _ in range(5):
                parts.append(u % 64)
                u //= 64
            nn_, nt_, ng_, nc_, na_ = parts
            table[i] = _entropy_of([na_, nc_, ng_, nt_, nn_], int(u))
        ent = table[inverse.ravel()] if key.size else np.zeros(0)
        cols["shannon_entropy"] = ent
        cols["normalized_entropy"] = ent / 2.0
    if "repeats" in wanted:
        cols["has_repeat"] = np.asarray(rep, dtype=np.int64)
    return cols


def _features_from_strings(kmers, wanted):
    """Slow path for files that already hold letter strings (not produced by KmerExtractor)."""
    out = defaultdict(list)
    for s in kmers:
        s = str(s)
        n = len(s)
        out["kmer"].append(s)
        if "gc_content" in wanted:
            out["gc_percent"].append(((s.count("G") + s.count("C")) / n) * 100 if n > 0 else 0)
        if "base_counts" in wanted:
            for b in "ACGT":
                out[f"{b}_count"].append(s.count(b))
        if "presence" in wanted:
            for b in "ACGT":
                out[f"{b}_present"].append(1 if b in s else 0)
        if "cpg_sites" in wanted:
            cpg = sum(1 for i in range(n - 1) if s[i:i + 2] == "CG")
            out["cpg_count"].append(cpg)
            prod = (s.count("C") / n) * (s.count("G") / n) if n > 0 else 0
            expected = prod * (n - 1) if prod > 0 else 0.001
            out["cpg_obs_exp"].append(cpg / expected if expected > 0 else 0)
        if "entropy" in wanted:
            h = _entropy_of([s.count(b) for b in sorted(set(s))], n)
            out["shannon_entropy"].append(h)
            out["normalized_entropy"].append(h / 2.0)
        if "repeats" in wanted:
            out["has_repeat"].append(1 if any(s[i:i + 2] == s[i + 2:i + 4] for i in range(n - 3)) else 0)
    return {k: np.asarray(v, dtype=object if k == "kmer" else None) for k, v in out.items()}


def _fast_csv(frames, genome_size):
    """The text DataFrame.to_csv(index=False) writes for the concatenated frames, built without formatting every
    cell: all columns after 'count' are functions of the (shortened) k-mer, and only a few thousand distinct
    combinations of them occur, so pandas formats one row per combination and the rest is string joining.
    Returns None when a frame does not have the expected shape (the caller then lets pandas write it)."""
    if not frames:
        return None
    columns = list(frames[0].columns)
    if columns[:3] != ["kmer", "count", "k"] or any(list(f.columns) != columns for f in frames):
        return None
    header = ",".join(columns + (["genome_size"] if genome_size else []))
    tail = f",{genome_size}" if genome_size else ""
    out = [header]
    for f in frames:
        n = len(f)
        if n == 0:
            continue
        rest = f.iloc[:, 2:]
        if rest.shape[1] > 1:
            # one small integer per row and distinct combination: factorise column by column
            key = np.zeros(n, dtype=np.int64)
            for name in rest.columns:
                codes, uniques = pd.factorize(rest[name].to_numpy(), sort=False)
                key, combos = pd.factorize(key * len(uniques) + codes, sort=False)
                if len(combos) > (1 << 22):
                    return None
            uniq, first, inverse = np.unique(key, return_index=True, return_inverse=True)
        else:
            uniq, first, inverse = np.zeros(1, np.int64), np.zeros(1, np.int64), np.zeros(n, np.int64)
        lines = rest.iloc[first].to_csv(index=False, header=False, lineterminator="\n").split("\n")[:len(first)]
        suffix = [ln + tail for ln in lines]
        kmers = f["kmer"].tolist()
        if any(("," in str(x)) or ('"' in str(x)) or ("\n" in str(x)) for x in kmers[:1]):
            return None
        counts = f["count"].tolist()
        inv = inverse.ravel().tolist()
        out.extend(f"{a},{b},{suffix[i]}" for a, b, i in zip(kmers, counts, inv))
    return "\n".join(out) + "\n"


def _gpu_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def gpu_feature_text(kmer_file, k_val, wanted, tail, device=None):
    """The CSV lines (no header) the reference writes for one k{k}.txt, produced on the GPU: the file's bytes are
    parsed, classified by composition, sized and written by the kernels of csrc/featcsv.cu; the host formats ONE
    suffix string per composition class that occurs (pandas' own formatting).  Returns bytes, or None when the file
    is not of the plain "<digits>\t<count>" form (the caller then takes the pandas path)."""
    import ctypes
    import gzip
    import torch
    from .. import _lib
    raw = gzip.open(kmer_file, "rb").read() if str(kmer_file).endswith(".gz") else Path(kmer_file).read_bytes()
    if not raw:
        return None
    dev = torch.device(device if device is not None else "cuda")
    L = _lib.load()
    ctx = _lib.context(dev.index if dev.index is not None else torch.cuda.current_device())
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    text = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
    ends = torch.nonzero(text == 10).flatten()
    if raw[-1] != 10:
        ends = torch.cat([ends, torch.tensor([len(raw)], dtype=torch.int64, device=dev)])
    n = int(ends.numel())
    value = torch.empty(n, dtype=torch.int64, device=dev)
    count = torch.empty(n, dtype=torch.int64, device=dev)
    bad = ctypes.c_uint32(0)
    _lib.check(L.kmerml_parse_kmer_lines(ctx.handle, text.data_ptr(), ends.data_ptr(), n, value.data_ptr(), count.data_ptr(),
                                         ctypes.byref(bad), stream))
    if bad.value or n == 0:
        return None
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    _lib.check(L.kmerml_feature_keys(ctx.handle, value.data_ptr(), n, keys.data_ptr(), stream))
    uniq, cls = torch.unique(keys, return_inverse=True)
    u = uniq.cpu().numpy()
    field = lambda shift, bits: (u >> shift) & ((1 << bits) - 1)                     # noqa: E731
    cols = {"k": np.full(u.size, k_val)}
    cols.update(columns_from_composition(field(0, 5), field(5, 5), field(10, 5), field(15, 5), field(20, 5), field(25, 5),
                                         field(30, 5), field(35, 1), wanted))
    lines = pd.DataFrame(cols).to_csv(index=False, header=False, lineterminator="\n").split("\n")[:u.size]
    suffix = [(ln + tail).encode() for ln in lines]
    lens = np.fromiter((len(x) for x in suffix), dtype=np.int32, count=len(suffix))
    offs = np.concatenate(([0], np.cumsum(lens)[:-1])).astype(np.int64)
    d_suffix = torch.frombuffer(bytearray(b"".join(suffix) or b"\0"), dtype=torch.uint8).to(dev)
    d_lens, d_offs = torch.from_numpy(lens).to(dev), torch.from_numpy(offs).to(dev)
    cls = cls.contiguous()
    line_len = torch.empty(n, dtype=torch.int64, device=dev)
    _lib.check(L.kmerml_feature_line_lengths(ctx.handle, value.data_ptr(), count.data_ptr(), cls.data_ptr(), d_lens.data_ptr(),
                                             n, line_len.data_ptr(), stream))
    line_off = torch.cumsum(line_len, 0) - line_len
    total = int((line_off[-1] + line_len[-1]).item())
    out = torch.empty(total, dtype=torch.uint8, device=dev)
    _lib.check(L.kmerml_feature_write_lines(ctx.handle, value.data_ptr(), count.data_ptr(), cls.data_ptr(), d_offs.data_ptr(),
                                            d_lens.data_ptr(), d_suffix.data_ptr(), line_off.data_ptr(), n, out.data_ptr(),
                                            stream))
    return out.cpu().numpy().tobytes()


class KmerFeatureExtractor:
    """Extract machine-learning features from k-mer count files."""

    def __init__(self, input_paths=None, output_dir=None, metadata_file=None, *, device="auto"):
        # device: "auto" = the GPU text path when a CUDA device is there (byte-identical output), None / "cpu" = the
        # vectorised pandas path only
        self.device = device
        self.input_paths = [Path(p) for p in input_paths] if input_paths else []
        self.output_dir = Path(output_dir) if output_dir else Path("kmer_features")
        self.output_dir.mkdir(exist_ok=True, parents=True)
        self.metadata = None
        if metadata_file:
            self.metadata_manager = _GenomeSizes(metadata_file)

    def add_paths(self, paths):
        self.input_paths.extend(Path(p) for p in paths)

    def extract_features(self, required_features=None):
        """Write one CSV per organism; returns {organism: csv path or None}."""
        wanted = list(ALL_FEATURES) if required_features is None else required_features
        return {organism: self._process_organism_kmers(organism, files, wanted)
                for organism, files in self._group_files_by_organism().items()}

    def _group_files_by_organism(self):
        groups = defaultdict(list)
        for path in self.input_paths:
            if path.is_file():
                groups[path.parent.name].append(path)
            elif path.is_dir():
                for sub in path.iterdir():
                    if sub.is_dir():
                        groups[sub.name].extend(sub.glob("k*.txt*"))
        return groups

    def _process_organism_kmers(self, organism, kmer_files, required_features):
        genome_size = None
        if hasattr(self, "metadata_manager"):
            genome_size = self.metadata_manager.get_genome_size(organism)
        frames = []
        use_gpu = self.device not in (None, "cpu") and _gpu_available()
        tail = f",{genome_size}" if genome_size else ""
        for kmer_file in kmer_files:
            k_val = self._extract_k_from_filename(kmer_file.name)
            if k_val is None:
                print(f"Warning: Could not extract k value from {kmer_file}")
                continue
            if use_gpu:
                body = gpu_feature_text(kmer_file, k_val, required_features, tail,
                                        None if self.device == "auto" else self.device)
                if body is not None:
                    frames.append(body)                  # (bytes: this file's CSV lines, already formatted)
                    continue
            table = self._load_kmer_file(kmer_file)
            if len(table):
                frames.append(self._extract_kmer_features(table, k_val, organism, required_features))
        if not frames:
            print(f"No features extracted for {organism}")
            return None
        output_file = self.output_dir / f"{organism}_kmer_features.csv"
        if any(isinstance(f, bytes) for f in frames):
            # header as DataFrame.to_csv writes it: the columns of any frame (all files share them)
            probe = self._extract_kmer_features(pd.DataFrame({"kmer": np.array([0], np.int64), "count": np.array([1], np.int64)}),
                                                0, organism, required_features)
            header = ",".join(list(probe.columns) + (["genome_size"] if genome_size else [])) + "\n"
            with open(output_file, "wb") as fh:
                fh.write(header.encode())
                for f in frames:
                    if isinstance(f, bytes):
                        fh.write(f)
                        continue
                    t = _fast_csv([f], genome_size)
                    if t is None:
                        g = f.copy()
                        if genome_size:
                            g["genome_size"] = genome_size
                        t = "\n" + g.to_csv(index=False, header=False)
                    fh.write(t.split("\n", 1)[1].encode())
            print(f"Created feature CSV for {organism}: {output_file}")
            return output_file
        text = _fast_csv(frames, genome_size)
        if text is not None:
            with open(output_file, "w", newline="") as fh:
                fh.write(text)
        else:
            result = pd.concat(frames, ignore_index=True) if len(frames) > 1 else frames[0]
            if genome_size:
                result["genome_size"] = genome_size
            result.to_csv(output_file, index=False)
        print(f"Created feature CSV for {organism}: {output_file}")
        return output_file

    def _extract_kmer_features(self, df, k_val, organism, required_features):
        """DataFrame of the feature columns for one k-mer file (all rows at once)."""
        col = df["kmer"]
        # (a digit column whose values exceed int64 but fit uint64 -- k = 20 without a leading C / G -- goes through
        # the string path like the reference's str(int))
        if pd.api.types.is_integer_dtype(col.dtype) and col.dtype != np.uint64:
            cols = features_from_digits(col.to_numpy(), required_features)
        else:
            decoded = [self._decode_kmer(x) if str(x).isdigit() else x for x in col.tolist()]
            cols = _features_from_strings(decoded, required_features)
        ordered = {"kmer": cols.pop("kmer"), "count": df["count"].to_numpy(), "k": np.full(len(df), k_val)}
        ordered.update(cols)
        return pd.DataFrame(ordered)

    @staticmethod
    def _extract_k_from_filename(filename):
        m = _K_IN_NAME.search(filename)
        return int(m.group(1)) if m else None

    @staticmethod
    def _decode_kmer(encoded_kmer):
        table = {"0": "A", "1": "T", "2": "C", "3": "G"}
        return "".join(table.get(ch, "N") for ch in str(encoded_kmer))

    @staticmethod
    def _load_kmer_file(filepath):
        """TSV -> DataFrame(kmer, count), read the way the reference reads it: the digit column
        is left to pandas' type inference, which is what drops the leading zeros."""
        compression = "gzip" if str(filepath).endswith(".gz") else None
        names = ["kmer", "count"]
        try:
            df = pd.read_csv(filepath, sep="\t", compression=compression)
            if "kmer" not in df.columns and "count" not in df.columns:
                df = pd.read_csv(filepath, sep="\t", header=None, names=names, compression=compression)
        except Exception:
            try:
                df = pd.read_csv(filepath, sep="\t", header=None, names=names, compression=compression)
            except pd.errors.EmptyDataError:
                df = pd.DataFrame({"kmer": pd.Series(dtype="int64"), "count": pd.Series(dtype="int64")})
        return df
