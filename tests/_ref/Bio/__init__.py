"""Throw-away stand-in for biopython (absent from this image; no network).

TEST INFRASTRUCTURE ONLY.  It exists so that the UNMODIFIED reference at
/root/reference can be imported in the build container to generate golden
vectors (tests/golden/make_golden.py).  It is never imported by the product.
"""
