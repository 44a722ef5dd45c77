"""CPU oracle for the k-mer extraction / feature path.  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product (kmerml_b200) never does.
"""
from .kmer_oracle import *  # noqa: F401,F403
