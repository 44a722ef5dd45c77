"""Minimal `Bio.SeqIO.parse(handle, "fasta")` restatement (biopython==1.85,
pinned by /root/reference/requirements.txt:1; source not vendored there).

Restates the published behaviour of biopython's plain-FASTA text iterator
(Bio/SeqIO/FastaIO.py, SimpleFastaParser/FastaIterator) that the reference
relies on at kmerml/kmers/generate.py:39-41 and
kmerml/utils/genome_metadata.py:68-70:

* the file is opened in TEXT mode (universal newlines: "\n", "\r\n" and a lone
  "\r" all end a line);
* every line before the first line whose first character is ">" is ignored;
* a line starting with ">" starts a record; record.id is the first
  whitespace-delimited word of the rest of that line ("" if none);
* every following line up to the next ">" line is `rstrip()`ed (trailing
  whitespace removed), the pieces are concatenated, then every " " and "\r"
  is removed; case is preserved (upper-casing is the reference's job).

TEST INFRASTRUCTURE ONLY — parity for FASTA corner cases is "unpinned" (no
reference test touches them); see DESIGN.md.
"""
from pathlib import Path


class _Seq:
    __slots__ = ("_data",)

    def __init__(self, data):
        self._data = data

    def __str__(self):
        return self._data

    def __len__(self):
        return len(self._data)

    def upper(self):
        return _Seq(self._data.upper())

    def count(self, sub):
        return self._data.count(sub)


class _Record:
    __slots__ = ("id", "name", "description", "seq")

    def __init__(self, title, sequence):
        words = title.split(None, 1)
        self.id = words[0] if words else ""
        self.name = self.id
        self.description = title
        self.seq = _Seq(sequence)


def _records(handle):
    title = None
    pieces = []
    for line in handle:
        if line[:1] == ">":
            if title is not None:
                yield _Record(title, "".join(pieces).replace(" ", "").replace("\r", ""))
            title = line[1:].rstrip()
            pieces = []
        elif title is not None:
            pieces.append(line.rstrip())
    if title is not None:
        yield _Record(title, "".join(pieces).replace(" ", "").replace("\r", ""))


def parse(source, fmt="fasta"):
    if fmt != "fasta":
        raise ValueError("shim only restates the 'fasta' format")
    if isinstance(source, (str, Path)):
        with open(source, "r") as handle:
            yield from _records(handle)
    else:
        yield from _records(source)
