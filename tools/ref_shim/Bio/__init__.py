"""Minimal stand-in for Biopython so the UNMODIFIED reference can run in this
container (Biopython is not installed, no network). Only used by
tools/make_golden.py to produce golden vectors; never on the product path."""
