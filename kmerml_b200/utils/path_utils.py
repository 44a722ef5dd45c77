"""File discovery helpers with the behaviour of the reference's
kmerml/utils/path_utils.py:4-60 (glob per pattern, optional recursion, sorted result;
mkdir -p; readable-file test).  Host-side only."""
import os
from pathlib import Path


def find_files(directory, patterns=None, recursive=False):
    """Sorted list of Paths under `directory` matching any glob in `patterns`
    (default ["*"]); with `recursive` every sub-directory is searched too."""
    root = Path(directory)
    prefix = "**/" if recursive else ""
    found = []
    for pattern in (patterns if patterns is not None else ["*"]):
        found += root.glob(prefix + pattern)
    return sorted(found)


def ensure_directory_exists(directory_path):
    """mkdir -p; returns the directory as a Path."""
    target = Path(directory_path)
    target.mkdir(parents=True, exist_ok=True)
    return target


def is_valid_file(file_path):
    """True for an existing, readable regular file."""
    target = Path(file_path)
    return target.is_file() and os.access(target, os.R_OK)
