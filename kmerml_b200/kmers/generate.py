"""KmerExtractor -- drop-in for the reference's kmerml/kmers/generate.py:7-129.

Same constructor, methods, prints, on-disk layout and file bytes; the window loop
(generate.py:39-58) runs on the B200 through libkmerml_b200.so.  Extensions are
keyword-only and default to the reference's behaviour (device=None -> current CUDA
device, canonical=False).
"""
import gzip
from pathlib import Path

import numpy as np
import torch

from .. import _lib, engine
from ..utils.path_utils import ensure_directory_exists

# lexicographic code (A0 C1 G2 T3) -> on-disk digit (A0 T1 C2 G3, generate.py:71)
_FILE_DIGIT = np.frombuffer(b"0231", dtype=np.uint8)
_POW10 = 10 ** np.arange(1, 20, dtype=np.uint64)


def format_kmer_lines(bins, counts, k):
    """bytes of the k{k}.txt lines "<digits>\\t<count>\\n" for the given bins, in order."""
    bins = np.asarray(bins, dtype=np.uint64)
    counts = np.asarray(counts, dtype=np.uint64)
    n = bins.size
    if n == 0:
        return b""
    ndig = np.searchsorted(_POW10, counts, side="right").astype(np.int64) + 1
    line_len = k + 2 + ndig
    starts = np.concatenate(([0], np.cumsum(line_len)[:-1]))
    out = np.empty(int(line_len.sum()), dtype=np.uint8)
    for i in range(k):
        code = ((bins >> np.uint64(2 * (k - 1 - i))) & np.uint64(3)).astype(np.intp)
        out[starts + i] = _FILE_DIGIT[code]
    out[starts + k] = 9
    rest = counts.copy()
    for d in range(int(ndig.max())):
        live = ndig > d
        pos = starts[live] + k + 1 + (ndig[live] - 1 - d)
        out[pos] = (48 + rest[live] % np.uint64(10)).astype(np.uint8)
        rest //= np.uint64(10)
    out[starts + k + 1 + ndig] = 10
    return out.tobytes()


class KmerExtractor:
    """Extract k-mers from genomic sequences for multiple k values (GPU)."""

    def __init__(self, output_dir="kmer_data", compress=True, *, device=None, canonical=False):
        self.output_dir = ensure_directory_exists(Path(output_dir))
        self.compress = compress
        self.device = device
        self.canonical = canonical

    # ------------------------------------------------------------------ core
    def _device(self):
        if not torch.cuda.is_available():
            raise _lib.KmermlError("KmerExtractor needs a CUDA device (kmerml_b200 has no CPU fallback)")
        return torch.device(self.device if self.device is not None else "cuda")

    def _records(self, dev_bytes, host_bytes, max_k):
        """[(record.id, is_too_short)] in file order."""
        ctx = _lib.context(dev_bytes.device.index)
        L = _lib.load()
        import ctypes
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev_bytes.device).cuda_stream)
        cap = 1024
        while True:
            offs = torch.empty(cap, dtype=torch.int64, device=dev_bytes.device)
            n = ctypes.c_uint32(0)
            _lib.check(L.kmerml_find_records(ctx.handle, dev_bytes.data_ptr(), dev_bytes.numel(), offs.data_ptr(),
                                             cap, ctypes.byref(n), stream))
            if n.value <= cap:
                break
            cap = int(n.value)
        offs = torch.sort(offs[:n.value]).values
        short = torch.zeros(max(n.value, 1), dtype=torch.uint8, device=dev_bytes.device)
        _lib.check(L.kmerml_records_short(ctx.handle, dev_bytes.data_ptr(), dev_bytes.numel(), offs.data_ptr(),
                                          n.value, int(max_k), short.data_ptr(), stream))
        offs_h = offs.cpu().numpy()
        short_h = short.cpu().numpy()
        out = []
        raw = host_bytes
        for i, o in enumerate(offs_h):
            o = int(o)
            end = o
            limit = len(raw)
            while end < limit and raw[end] not in (10, 13):
                end += 1
            title = bytes(raw[o + 1:end]).decode("utf-8", "replace").rstrip()
            words = title.split(None, 1)
            out.append((words[0] if words else "", bool(short_h[i])))
        return out

    def extract_kmers_from_fasta(self, fasta_file, k_values, organism_id=None):
        """Count the k-mers of every k in `k_values` over all records of one FASTA file
        and write <output_dir>/<organism_id>/k{k}.txt[.gz]; returns the organism id."""
        if organism_id is None:
            organism_id = Path(fasta_file).stem
        # one dict per DISTINCT k (generate.py:36), but `for k in k_values` (:49) runs once per entry: a k listed
        # m times is counted m times.  Counted once here, multiplied when the file is written.
        ks = list(dict.fromkeys(int(k) for k in k_values))
        mult = {k: sum(1 for v in k_values if int(v) == k) for k in ks}
        max_k = max(ks)
        if max_k > _lib.MAX_K or min(ks) < 1:
            raise _lib.KmermlError(f"k must be in 1..{_lib.MAX_K}")
        device = self._device()
        host = np.fromfile(str(fasta_file), dtype=np.uint8)
        dev = torch.from_numpy(host).to(device) if host.size else torch.zeros(0, dtype=torch.uint8, device=device)

        for rid, too_short in self._records(dev, host, max_k):
            if too_short:
                print(f"Skipping {rid}: too short for k-mer extraction")
            else:
                print(f"Processed chromosome/contig: {rid}")

        dense = [k for k in ks if k <= _lib.MAX_DENSE_K]
        res = engine.count_dense_device(dev, [0, int(dev.numel())], dense, min_record_len=max_k,
                                        canonical=self.canonical, want_freq=False) if dense else None
        for k in ks:
            if k > _lib.MAX_DENSE_K:                     # sparse path: distinct k-mers, sorted by first occurrence
                keys, cnts, first, _ = engine.count_sparse_device(dev, k, min_record_len=max_k, canonical=self.canonical)
                order = torch.argsort(first.to(torch.int64) & 0xFFFFFFFF, stable=True)
                text = engine.format_kmer_lines_device(keys[order].contiguous(), self._times(cnts[order], mult[k]), k)
            else:                                        # dense row + first-occurrence offsets -> text, on the GPU
                row = self._times(res.counts_of(0, k), mult[k])
                first = engine.first_occurrence_device(dev, k, min_record_len=max_k)
                text = engine.format_kmer_file_device(row, first, k, canonical=self.canonical)
            self._write_lines(organism_id, k, text.cpu().numpy())
        return organism_id

    @staticmethod
    def _times(counts, m):
        """uint32 counts (int32 storage) times the multiplicity of k in k_values."""
        if m == 1:
            return counts.contiguous()
        wide = (counts.to(torch.int64) & 0xFFFFFFFF) * m
        if wide.numel() and int(wide.max().item()) >= 1 << 32:
            raise _lib.KmermlError("a k-mer count times the multiplicity of k in k_values exceeds 32 bits")
        return wide.to(torch.int32).contiguous()           # (wraps: int32 storage of uint32 values)

    # ------------------------------------------------------------- file output
    def _target(self, organism_id, k):
        folder = ensure_directory_exists(self.output_dir / organism_id)
        return folder / (f"k{k}.txt.gz" if self.compress else f"k{k}.txt")

    def _write_lines(self, organism_id, k, payload):
        path = self._target(organism_id, k)
        if self.compress:
            with gzip.open(path, "wb") as fh:
                fh.write(payload)
        else:
            with open(path, "wb") as fh:
                fh.write(payload)

    def _save_kmers_to_file(self, kmers, organism_id, k):
        """Write a {kmer string: count} mapping in its iteration order (the reference's
        private writer, generate.py:68-91); unknown letters become 'X'."""
        digit = {"A": "0", "T": "1", "C": "2", "G": "3"}
        text = "".join("".join(digit.get(b, "X") for b in kmer) + f"\t{count}\n" for kmer, count in kmers.items())
        self._write_lines(organism_id, k, text.encode())

    # ------------------------------------------------------------------ batch
    def extract_from_genome_list(self, genome_paths, k_values, organism_ids=None):
        """Run extract_kmers_from_fasta over several genomes; a failing genome is reported
        and skipped.  Returns the ids that were processed."""
        if organism_ids is None:
            organism_ids = [Path(p).stem for p in genome_paths]
        if len(organism_ids) != len(genome_paths):
            raise ValueError("Number of organism IDs must match number of genome paths")
        done = []
        total = len(genome_paths)
        for index, (path, org) in enumerate(zip(genome_paths, organism_ids), start=1):
            print(f"Processing genome {org} ({index}/{total})")
            try:
                self.extract_kmers_from_fasta(path, k_values, org)
            except Exception as exc:
                print(f"Error processing {org}: {str(exc)}")
                continue
            done.append(org)
            print(f"Completed {org}")
        print(f"Completed processing {len(done)} out of {total} genomes")
        return done
